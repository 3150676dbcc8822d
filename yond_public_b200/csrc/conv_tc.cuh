// Implicit-GEMM convolution on tcgen05 / TMEM / TMA (sm_100a) — host-side interface.
#pragma once
#include "common.cuh"

enum ConvMode {
  CONV_3X3_S1 = 0,  // 3x3, stride 1, pad 1 (halo slab in smem shared by the 3 vertical taps)
  CONV_1X1 = 1,     // 1x1 (plain GEMM over pixels)
  CONV_3X3_S2 = 2,  // 3x3, stride 2, pad 1 (GuidedResUnet / SNRnet "pool"; parity-split tensor maps)
  CONVT_2X2 = 3,    // ConvTranspose2d 2x2 stride 2 (GEMM with N = 4*Cout, scatter epilogue)
  // Fused decoder hand-over of GuidedResUnet / SNRnet: ConvTranspose2d(2x2, s2)(src0) followed by the 1x1 shortcut conv on
  // cat[up, src1].  The up-sampling folds into the shortcut's weights (W'[a,b] = Wsc[:, :C] * Wt[:, :, a, b]^T), so the layer is
  // one GEMM per output parity (a,b): [src0 | src1 at parity (a,b)] x [W'[a,b] ; Wsc[:, C:]] and the up-sampled tensor is
  // never written.  src0: (B, Hin, Win, Cin0) low resolution, src1: (B, 2 Hin, 2 Win, Cin1 = Cout) skip.
  CONV_UPSC = 4,
};
enum ConvAct { ACT_NONE = 0, ACT_LRELU = 1, ACT_SILU = 2 };

struct ConvLayer {
  int mode;
  int B, Hin, Win;          // input spatial size per image
  int Cin0, Cin1;           // channels of the (up to two, concatenated) NHWC bf16 sources
  const bf16* src0;
  const bf16* src1;
  int Cout;                 // real output channels
  const bf16* wpacked;      // [cbg][tap][N][CB] bf16, N = Cout (x4 for CONVT_2X2: n = (a*2+b)*Cout + co);
                            // CONV_UPSC: [cb0][4*Cout][CB] (folded up-sampling part) then [cb1][Cout][CB] (skip part)
  const bf16* wpaired;      // optional (3x3 s1, 32 -> 32 channels): the same weights packed for the pixel-pair formulation, see
                            // conv_tc_pack_paired; conv_tc_launch uses it when the width is even
  const float* bias;        // [Cout]
  const float* scale;       // [B][Cout] or null: v = v*scale + shift   (FiLM / SNR gates)
  const float* shift;       // [B][Cout] or null
  int act;                  // applied after scale/shift
  float slope;
  const bf16* res;          // [B,Hout,Wout,Cout] or null, added after the activation
  bf16* out0;               // [B,Hout,Wout,Cout]
  bf16* out1;               // optional second output = SiLU(out0 value)
  // Fused last layer of the network (Cout == 32 only): instead of storing out0, the epilogue applies the 1x1 conv nf -> 4 of
  // archs/Unet.py:96-103 / :462-469 to the finished pixel, adds the network input and multiplies by the per-sample maximum
  // (data_inv_normalize): y = (sum_c out[c] * tail_w[c][:] + tail_b + (tail_res ? tail_z / ub : 0)) * ub, float32 NHWC4.
  // The 64 B/px activation tensor is neither written nor re-read.  tail_w == nullptr: off.
  const float* tail_w;      // [Cout][4]
  const float* tail_b;      // [4]
  const float* tail_z;      // [B,H,W,4] network input (not normalised)
  const float* tail_ub;     // [B] or null
  float* tail_y;            // [B,H,W,4]
  int tail_res;
};

// Channel block (K slice per pipeline stage) used for a layer: 64 when every source allows it, else 32.
int conv_tc_channel_block(int Cin0, int Cin1);
// Packed weight size in elements for a layer.
size_t conv_tc_packed_elems(int mode, int Cin_total, int Cout);
// 3x3 stride-1 conv with 32 input and 32 output channels as a conv over PIXEL PAIRS: the NHWC tensor (B,H,W,32) is the
// tensor (B,H,W/2,64) of pixel pairs, and out[2X+o] = sum_{r,P,i} W[r][2P+i-o] in[2(X+P)+i] is a 3x3 conv with 64 -> 64
// "channels" whose weight is zero where |2P+i-o| > 1.  The MMA then runs with N = 64 (48 cycles per M=128,K=16 step instead of
// 40 for N = 32, for twice the outputs) and the all-zero K steps (left pair: first pixel, right pair: second pixel) are skipped:
// 24 MMAs per 256 pixels instead of 36 per 256.  Layout [tap = r*3 + (P+1)][n = o*32 + co][k = i*32 + ci], bf16.
// w: torch layout (32, 32, 3, 3).
void conv_tc_pack_paired(const float* w, bf16* out);
constexpr size_t kConvPairedElems = 9 * 64 * 64;
int conv_tc_launch(const ConvLayer& L, cudaStream_t stream);
// CUDA-core direct convolution with identical semantics (debug cross-check; reads the same packed weights).
int conv_ref_launch(const ConvLayer& L, cudaStream_t stream);
