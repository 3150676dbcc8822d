// CUDA-core direct convolution with the same semantics, operands and packed-weight layout as conv_tc.cu.
// It exists to cross-check the tcgen05 kernels on the GPU (tests, yond_net_set_conv_impl(net, 1)); it is
// not the product path.
#include "conv_tc.cuh"

namespace {

struct RefParams {
  int mode, B, Hin, Win, Hout, Wout, Cin0, Cin1, Cout, N, CB, taps;
  const bf16* src0;
  const bf16* src1;
  const bf16* w;
  const float* bias;
  const float* scale;
  const float* shift;
  int act;
  float slope;
  const bf16* res;
  bf16* out0;
  bf16* out1;
};

__device__ __forceinline__ float in_at(const RefParams& p, int b, int h, int w, int ci) {
  if (h < 0 || h >= p.Hin || w < 0 || w >= p.Win) return 0.f;
  size_t pix = ((size_t)b * p.Hin + h) * p.Win + w;
  return ci < p.Cin0 ? __bfloat162float(p.src0[pix * p.Cin0 + ci]) : __bfloat162float(p.src1[pix * p.Cin1 + (ci - p.Cin0)]);
}

__global__ void conv_ref_kernel(const RefParams p) {
  const size_t total = (size_t)p.B * p.Hout * p.Wout * p.Cout;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % p.Cout);
    size_t pix = idx / p.Cout;
    const int wo = (int)(pix % p.Wout);
    const int ho = (int)((pix / p.Wout) % p.Hout);
    const int b = (int)(pix / ((size_t)p.Wout * p.Hout));
    const int Cin = p.Cin0 + p.Cin1;
    float acc = 0.f;
    if (p.mode == CONVT_2X2) {
      const int h = ho >> 1, w = wo >> 1, n = ((ho & 1) * 2 + (wo & 1)) * p.Cout + co;
      for (int ci = 0; ci < Cin; ++ci) {
        const int cbg = ci / p.CB, j = ci % p.CB;
        acc += in_at(p, b, h, w, ci) * __bfloat162float(p.w[((size_t)cbg * p.N + n) * p.CB + j]);
      }
    } else {
      const int st = (p.mode == CONV_3X3_S2) ? 2 : 1;
      const int kk = (p.mode == CONV_1X1) ? 1 : 3;
      const int pad = (kk == 3) ? 1 : 0;
      for (int r = 0; r < kk; ++r)
        for (int s = 0; s < kk; ++s) {
          const int h = ho * st + r - pad, w = wo * st + s - pad;
          if (h < 0 || h >= p.Hin || w < 0 || w >= p.Win) continue;
          const int tap = r * kk + s;
          for (int ci = 0; ci < Cin; ++ci) {
            const int cbg = ci / p.CB, j = ci % p.CB;
            acc += in_at(p, b, h, w, ci) * __bfloat162float(p.w[(((size_t)cbg * p.taps + tap) * p.N + co) * p.CB + j]);
          }
        }
    }
    float v = acc + p.bias[co];
    if (p.scale) v *= p.scale[(size_t)b * p.Cout + co];
    if (p.shift) v += p.shift[(size_t)b * p.Cout + co];
    if (p.act == ACT_LRELU) v = v > 0.f ? v : v * p.slope;
    else if (p.act == ACT_SILU) v = silu_f(v);
    if (p.res) v += __bfloat162float(p.res[idx]);
    p.out0[idx] = __float2bfloat16_rn(v);
    if (p.out1) p.out1[idx] = __float2bfloat16_rn(silu_f(v));
  }
}

}  // namespace

int conv_ref_launch(const ConvLayer& L, cudaStream_t stream) {
  RefParams p{};
  p.mode = L.mode;
  p.B = L.B;
  p.Hin = L.Hin;
  p.Win = L.Win;
  p.Hout = L.mode == CONV_3X3_S2 ? L.Hin / 2 : (L.mode == CONVT_2X2 ? L.Hin * 2 : L.Hin);
  p.Wout = L.mode == CONV_3X3_S2 ? L.Win / 2 : (L.mode == CONVT_2X2 ? L.Win * 2 : L.Win);
  p.Cin0 = L.Cin0;
  p.Cin1 = L.Cin1;
  p.Cout = L.Cout;
  p.N = L.mode == CONVT_2X2 ? 4 * L.Cout : L.Cout;
  p.CB = conv_tc_channel_block(L.Cin0, L.Cin1);
  p.taps = (L.mode == CONV_3X3_S1 || L.mode == CONV_3X3_S2) ? 9 : 1;
  p.src0 = L.src0;
  p.src1 = L.src1;
  p.w = L.wpacked;
  p.bias = L.bias;
  p.scale = L.scale;
  p.shift = L.shift;
  p.act = L.act;
  p.slope = L.slope;
  p.res = L.res;
  p.out0 = L.out0;
  p.out1 = L.out1;
  const size_t total = (size_t)p.B * p.Hout * p.Wout * p.Cout;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 32) grid = 148 * 32;
  conv_ref_kernel<<<grid, 256, 0, stream>>>(p);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}
