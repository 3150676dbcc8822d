/*
 * yond_b200.h — C-ABI of libyond_b200.so: B200 (sm_100a) kernels for YOND's per-image blind raw
 * denoising path (fenghansen/YOND_public).  This is the drop-in boundary: plain pointers and sizes,
 * no torch / C++ types.  Every entry point names the reference interface it replaces (file:line in
 * the reference tree).  The reference is pure Python; a maintainer binds this library with ctypes
 * (see INTEGRATION.md — the stub is yond_public_b200/_lib.py).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless the name says `host`.  The caller (PyTorch) owns
 *     every buffer; the library never frees caller memory.  Opaque handles own packed weights.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All functions are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - Return value: 0 = ok, non-zero = error; yond_last_error() gives the message (thread-local).
 *   - Layouts: Bayer frames (B,H,W) f32; packed frames NHWC (B,h,w,4) f32 with channel = 2*(row&1)+(col&1)
 *     (the reference's `bayer2rggb` order, CFA-agnostic); activations NHWC bf16.
 */
#ifndef YOND_B200_H
#define YOND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOND_OK 0
#define YOND_ERR_INVALID 1
#define YOND_ERR_CUDA 2
#define YOND_ERR_UNSUPPORTED 3

const char* yond_last_error(void);
int yond_version(void);
/* Number of kernels launched by this library since process start (bench.py's `gpu_launches`). */
uint64_t yond_launch_count(void);

/* ---- A1/A2: Bayer pack / unpack — utils/isp_ops.py:57-63 (bayer2rggb, rggb2bayer), batched :65-71 ---- */
int yond_pack(const float* bayer, float* rggb, int B, int H, int W, void* stream);
int yond_unpack(const float* rggb, float* bayer, int B, int h, int w, void* stream);

/* ---- RAW ingest — data_process/process.py:40-64 (pack_raw_bayer) ----
 * uint16 sensor mosaic -> four float32 planes in the order R, G1, B, G2 given by the 2x2 `raw_pattern`
 * (pos4[c] = 2*row + col of colour c inside the CFA cell), out = (v - black[c]) / (white - black[c]) in float32 with
 * IEEE subtraction / division like NumPy, optionally clipped to [0,1].  `pos4` and `black4` are HOST arrays.
 * layout 0: (B,4,H/2,W/2) planes like the reference; layout 1: (B,H/2,W/2,4) interleaved (what yond_vst_fwd-style
 * kernels and the estimator consume).  Reads 2 B/px instead of the 4 B/px of a float32 mosaic. */
int yond_pack_raw(const uint16_t* raw, float* out, int B, int H, int W, const int* pos4, const float* black4, float white,
                  int clip, int layout, void* stream);

/* CFA canonicalisation — utils/sidd_utils.py:198-213 (rot_bayer = np.rot90 by k quarter turns, counter-clockwise, over the
 * last two axes; the driver rotates every SIDD frame to the RGGB phase before denoising and back afterwards,
 * YOND_SIDD.py:403,463).  in: (B,H,W) float32, out: (B,W,H) for odd k, (B,H,W) for even k.  Pure data movement. */
int yond_rot90(const float* in, float* out, int B, int H, int W, int k, void* stream);

/* ---- A3/A4 elementwise, for the function-level surface — utils/isp_algos.py:5-14, :17-33 ---- */
int yond_vst(const float* x, float* z, size_t n, double sigma, double gain, void* stream);
int yond_inverse_vst(const float* z, float* x, size_t n, double sigma, double gain, int exact, void* stream);

/* ---- A5: BiasLUT — utils/isp_algos.py:162-231.
 * yond_lut_row: sigma-lerp of the (nx=1921, nsg=1101) [x,sigma] table into one nx-entry row (data_merge over
 *   sigma, :225); `sg_pos` is the fractional sigma index computed on the host (pos_interp, :199).
 * yond_lut_apply: per-element lookup bias(max(x,0)/K) with the reference's piecewise-linear node inversion
 *   (pos_interp :179-186 + data_merge :188-194); x in DN units.  `xnodes` = the nx node positions (electrons). */
int yond_lut_row(const float* lut2d, int nx, int nsg, double sg_pos, float* row, void* stream);
int yond_lut_apply(const float* x, float* bias, size_t n, const float* row, const float* xnodes, int nx,
                   double gain, double sigma, void* stream);

/* Per-frame parameters of the fused VST stages (one entry per frame of a batch). */
typedef struct {
  float gain;      /* K   (DN)                                   YOND_SIDD.py:356 */
  float sigma;     /* sigma_read (DN)                                               */
  float scale;     /* wp - bl (/ratio)                            YOND_SIDD.py:251,504 */
  float lower;     /* VST(0)                                      YOND_SIDD.py:264 */
  float upper;     /* VST(scale)                                  YOND_SIDD.py:265 */
  int32_t lut_row; /* row index into `rows` (−1: no bias correction, bias_corr=None) */
  int32_t table_n; /* 0: `rows[lut_row]` is a sigma-interpolated BiasLUT row (1921 nodes, electrons);
                      >0: a fallback get_bias table (isp_algos.py:98-140) with `table_n` nodes in DN */
  int32_t exact_inverse; /* 1: closed-form exact unbiased inverse (isp_algos.py:20-27) */
} yond_vst_params;

/* ---- A18 front half (YOND_SIDD.py:251-269,275,281-282,286): pack*scale -> bias -> VST-bias -> normalise ->
 * clamp(0,1) -> reflect-pad to (hp,wp) -> z (B,hp,wp,4) f32; also ub[b] = max(z[b]) (A14, modules.py:15-21).
 * `rows`: (nrows, row_stride) f32 bias tables; `xnodes`: (nrows, row_stride) node positions of each row (the BiasLUT
 * x-grid in electrons for LUT rows, the get_bias nodes in DN for fallback tables).  p2d = (left, right, top, bottom) in packed pixels (utils/utils.py:246-252). */
int yond_vst_fwd(const float* bayer, float* z, float* ub, int B, int H, int W, int pad_l, int pad_r, int pad_t,
                 int pad_b, const yond_vst_params* params_dev, const float* rows, const float* xnodes,
                 int row_stride, void* stream);
/* ---- A18 back half (YOND_SIDD.py:286,289-298, caller's clip :389/:406): y (B,hp,wp,4) f32 -> clamp(0,1) -> crop ->
 * de-normalise -> inverse VST -> unpack -> /scale -> [clip 0..1] -> Bayer (B,H,W) f32. */
int yond_vst_inv(const float* y, float* bayer, int B, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                 const yond_vst_params* params_dev, int clip01, void* stream);
/* Simple_Denoiser's front/back (YOND_SIDD.py:238-248): pack -> reflect pad -> clamp, and clamp -> crop -> unpack. */
int yond_pack_pad(const float* bayer, float* z, float* ub, int B, int H, int W, int pad_l, int pad_r, int pad_t,
                  int pad_b, void* stream);
int yond_crop_unpack(const float* y, float* bayer, int B, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                     void* stream);

/* ---- A7/A8/A11: box statistics — utils/isp_algos.py:234-242 (stdfilt = cv2.blur pair), YOND_SIDD.py:62-71, :89-98.
 * Input: packed frames (B,h,w,C) f32 interleaved (C = 4, or 128 for the SIDD_256 channel stack).
 * yond_box_blur: normalised k x k box, BORDER_REFLECT_101, float64 sums -> f32 (what cv2.blur computes).
 * yond_nlf_maps: mode 0 (self)  : var = std_k(x)^2, mean = blur_k(x), lap = std_k(blur_k2(x)), k2 = k/3*2+1
 *                mode 1 (collab): var = std_k(x)^2 - std_k(y)^2, mean = blur_k(y), lap = std_k(y)
 * `work`: scratch of yond_nlf_work_bytes(). */
int yond_box_blur(const float* x, float* out, int B, int h, int w, int C, int k, int square_input, void* work,
                  void* stream);
size_t yond_nlf_work_bytes(int B, int h, int w, int C);
int yond_nlf_maps(const float* x, const float* y, float* var, float* mean, float* lap, int B, int h, int w, int C,
                  int k, int mode, void* work, void* stream);

/* ---- A9: get_threshold(mode='score3') — YOND_SIDD.py:22-49.  All three entry points are batched over `nseg`
 * independent segments (one per image) of `seg_len` contiguous floats each.
 * yond_order_stats: exact k-th smallest values (0-based ranks, shared by all segments, nranks <= 64) by radix select;
 *   out (nseg, nranks).  The host applies np.percentile's linear interpolation in float64.
 *   `work`: yond_select_work_bytes(nseg).
 * yond_score3_bins: npeaks[s][i] = number of occupied bins of int(clip(mean,0,1)*1000) over {lap <= ths[s][i]} (:37-43),
 *   ths ascending per segment (float64, (nseg, nth), nth <= 32).  `work`: >= nseg*1001*4 bytes. */
size_t yond_select_work_bytes(int nseg);
int yond_order_stats(const float* data, size_t seg_len, int nseg, const uint64_t* ranks_dev, int nranks, float* out_dev,
                     void* work, void* stream);
int yond_score3_bins(const float* lap, const float* mean, size_t seg_len, int nseg, const double* ths_dev, int nth,
                     int32_t* npeaks_dev, void* work, void* stream);
/* ---- A10: masked line fit — YOND_SIDD.py:77-78 (strict lap<th), utils/isp_algos.py:345-365.  Per segment s:
 * sums_dev[s][0..5]  = {N, Sx, Sy, Sxx, Sxy, Syy} over {lap < ths_dev[s]};
 * sums_dev[s][6..11] = same over {lap < th, 1e-4 < mean < 0.8} (polyfit's non-saturated subset), float64. */
int yond_masked_sums(const float* lap, const float* mean, const float* var, size_t seg_len, int nseg,
                     const double* ths_dev, double* sums_dev, void* stream);

/* ---- A14-A17, A20: denoiser networks — archs/Unet.py:4-104 (UNetSeeInDark), :380-470 (GuidedResUnet),
 * :288-378 (SNRnet); blocks archs/modules.py:117-125,163-233.  Plugin descriptor = the yml `arch:` block. */
typedef struct yond_net yond_net_t;
#define YOND_ARCH_UNET 0     /* UNetSeeInDark  */
#define YOND_ARCH_GUIDED 1   /* GuidedResUnet  */
#define YOND_ARCH_SNR 2      /* SNRnet         */
/* Creates a network; weights are set tensor-by-tensor with the reference's state_dict keys. */
int yond_net_create(int arch, int in_nc, int out_nc, int nf, int res, int norm, yond_net_t** out);
void yond_net_destroy(yond_net_t* net);
/* `host_data`: f32 tensor in the reference's (PyTorch) layout — Conv2d (Cout,Cin,kh,kw), ConvTranspose2d
 * (Cin,Cout,2,2), bias (C).  Repacked once to the kernels' layouts (per-tap K-major bf16).  utils/utils.py:160-209. */
int yond_net_set_tensor(yond_net_t* net, const char* key, const float* host_data, const int64_t* shape, int ndim);
/* The state_dict this network expects, in the reference's registration order (keys, shapes). */
int yond_net_num_keys(yond_net_t* net);
const char* yond_net_key(yond_net_t* net, int i);
int yond_net_key_shape(yond_net_t* net, int i, int64_t* shape4); /* returns ndim */
/* Returns the number of state-dict tensors still unset (0 = ready); `missing` receives a ';'-joined key list. */
int yond_net_missing(yond_net_t* net, char* missing, size_t cap);
size_t yond_net_workspace_bytes(yond_net_t* net, int B, int H, int W);
/* net(x[, t]) on NHWC input: z (B,H,W,4) f32 in [0,1] (H,W multiples of 16), ub (B) = per-sample max (used when
 * norm=1; pass the buffer yond_vst_fwd / yond_pack_pad filled), t (B) f32 per-sample guidance value BEFORE the
 * division by ub (archs/Unet.py:427-429; NULL for UNetSeeInDark).  y (B,H,W,4) f32 = network output (not clamped). */
int yond_net_forward(yond_net_t* net, const float* z, const float* ub, const float* t, float* y, int B, int H,
                     int W, void* workspace, size_t workspace_bytes, void* stream);
/* Module-level drop-in: NCHW f32 in / out like nn.Module.forward; computes ub itself (data_normalize). */
int yond_net_forward_nchw(yond_net_t* net, const float* x, const float* t, float* y, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* stream);
/* FLOPs of the tensor-core conv stack for one forward of this shape (2*MAC, algorithmic, no halo / padding). */
double yond_net_flops(yond_net_t* net, int B, int H, int W);
/* 0: tcgen05 implicit-GEMM kernels (product path).  1: CUDA-core direct convolution (debug cross-check only). */
int yond_net_set_conv_impl(yond_net_t* net, int impl);
/* Device time (ms) accumulated by the conv-stack kernels since the last reset, measured with CUDA events on
 * `stream` when profiling is enabled (bench.py's live roofline). */
int yond_net_profile(yond_net_t* net, int enable);
int yond_net_profile_read(yond_net_t* net, double* conv_ms, double* conv_flops, int* launches, int reset);

/* ---- single conv layer on NHWC bf16 activations (the building block of the networks above; torch.nn.Conv2d /
 * ConvTranspose2d in archs/Unet.py).  mode: 0 = 3x3 s1 p1, 1 = 1x1, 2 = 3x3 s2 p1, 3 = ConvTranspose 2x2 s2.
 * The input is the channel concatenation of src0 (Cin0) and src1 (Cin1, may be 0/NULL).  `weight_host`: f32 in the
 * PyTorch layout; bias/scale/shift/res/out are device pointers.  Epilogue: v = acc + bias; v = v*scale[b] + shift[b];
 * act (0 none, 1 LeakyReLU(slope), 2 SiLU); v += res; out0 = bf16(v); out1 = bf16(SiLU(v)) if given.
 * impl: 0 = tcgen05 kernel, 1 = CUDA-core cross-check.  Synchronous (packs and uploads the weights per call). */
int yond_conv2d(int mode, int impl, int B, int Hin, int Win, int Cin0, int Cin1, const void* src0, const void* src1,
                int Cout, const float* weight_host, const float* bias, const float* scale, const float* shift, int act,
                float slope, const void* res, void* out0, void* out1, void* stream);

/* ---- tiling helpers (new design; reference semantics utils/utils.py:254-268 + whole-frame forward) ----
 * Copies a halo-extended tile out of / back into a padded NHWC4 frame; out-of-frame halo pixels are zero
 * (what the network's own zero padding would have seen). */
int yond_tile_extract(const float* frame, float* tile, int H, int W, int y0, int x0, int th, int tw, void* stream);
int yond_tile_insert(const float* tile, float* frame, int H, int W, int y0, int x0, int th, int tw, int halo_t,
                     int halo_l, int core_h, int core_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOND_B200_H */
