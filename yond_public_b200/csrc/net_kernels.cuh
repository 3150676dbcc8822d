// CUDA-core kernels around the tensor-core conv stack — host-side launchers.
#pragma once
#include "common.cuh"

struct FilmWeights {
  const float* w0; const float* b0;  // (C,1) , (C)      first 1x1 on the scalar t
  const float* w2; const float* b2;  // (C,C) , (C)
  const float* w3; const float* b3;  // guided: beta.1 (C,C),(C)   | snr: sfm2.0 (C,1),(C)
  const float* w4; const float* b4;  // snr only: sfm2.2 (C,C),(C)
};

int head_conv_launch(const float* z, const float* ub, const float* w, const float* bias, int B, int H, int W, int nf,
                     float slope, bf16* out0, bf16* out1, cudaStream_t s);
int tail_conv_launch(const bf16* act, const float* w, const float* bias, const float* z, const float* ub, int res, int B,
                     int H, int W, int nf, float* y, cudaStream_t s);
int maxpool2_launch(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s);
// SelfResUNet / GuidedSelfUnet up path (archs/comp.py:815-826): out (B, 2 Hlo, 2 Wlo, C0 + C1) = cat[nearest-neighbour x2 of lo (C0
// channels), skip (C1 channels at the high resolution; may be null with C1 = 0)], NHWC bf16, C0 and C1 multiples of 8.
int upcat_launch(const bf16* lo, const bf16* skip, bf16* out, int B, int Hlo, int Wlo, int C0, int C1, cudaStream_t s);
// act (B,H,W,C) bf16 += w_in[C][4] . (z / ub): the 4 input channels of a 1x1 conv on cat[features, network input], in float32.
int add_in4_launch(bf16* act, const float* w_in, const float* z, const float* ub, int B, int H, int W, int C, cudaStream_t s);
struct FilmAll {  // every conditioned block of a network, evaluated by one launch
  FilmWeights fw[12];
  int C[12];
  float* out_a[12];
  float* out_b[12];
  int n;
};
int film_launch(const FilmAll& all, const float* t, const float* ub, int B, int guided, cudaStream_t s);
int nchw_to_nhwc4_launch(const float* x, float* z, float* ub, int B, int H, int W, cudaStream_t s);
int nhwc4_to_nchw_launch(const float* y, float* out, int B, int H, int W, cudaStream_t s);
