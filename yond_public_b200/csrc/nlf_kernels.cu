// Noise-level-function estimation kernels (SimpleNLF): box statistics, exact order statistics, the
// 'score3' bin-occupancy count and the masked regression sums.
//   reference: utils/isp_algos.py:234-242 (stdfilt via cv2.blur), :345-365 (polyfit);
//              YOND_SIDD.py:22-49 (get_threshold 'score3'), :62-115 (SelfNLF / CollabNLF).
// cv2.blur on float32 = normalised box, BORDER_REFLECT_101, float64 running sums, result rounded to float32;
// the kernels below keep float64 sums (vertical pass -> float64 scratch -> horizontal pass) to match it.
#include "common.cuh"

namespace {

constexpr int kMaxK = 31;

struct D4 {
  double x, y, z, w;
};
__device__ __forceinline__ void d4_add(D4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ void d4_sub(D4& a, const float4& v) { a.x -= v.x; a.y -= v.y; a.z -= v.z; a.w -= v.w; }
__device__ __forceinline__ float4 sq4(const float4& v) {
  return make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w));
}

// Vertical pass: one thread owns one pixel column (4 channels) of a strip of `rows_per_strip` output rows and
// slides a k-tall window down it.  Writes float64 column sums of x (and of fl32(x*x) when S2 != nullptr).
__global__ void __launch_bounds__(128) box_v_kernel(const float4* __restrict__ x, D4* __restrict__ S1, D4* __restrict__ S2,
                                                    int h, int w, int k, int rows_per_strip, int square_input) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  if (col >= w) return;
  const int r = k / 2;
  const int i0 = blockIdx.y * rows_per_strip;
  const int i1 = min(h, i0 + rows_per_strip);
  const float4* xb = x + (size_t)b * h * w;
  D4 s1{0, 0, 0, 0}, s2{0, 0, 0, 0};
  auto ld = [&](int i) {
    float4 v = __ldg(xb + (size_t)reflect101(i, h) * w + col);
    return square_input ? sq4(v) : v;
  };
  for (int i = i0 - r; i <= i0 + r; ++i) {
    const float4 v = ld(i);
    d4_add(s1, v);
    if (S2) d4_add(s2, sq4(v));
  }
  for (int i = i0; i < i1; ++i) {
    const size_t o = ((size_t)b * h + i) * w + col;
    S1[o] = s1;
    if (S2) S2[o] = s2;
    if (i + 1 < i1) {
      const float4 vn = ld(i + 1 + r), vo = ld(i - r);
      d4_add(s1, vn);
      d4_sub(s1, vo);
      if (S2) {
        d4_add(s2, sq4(vn));
        d4_sub(s2, sq4(vo));
      }
    }
  }
}

// Horizontal pass over the float64 column sums.  A block stages kHCols pixels (+halo) of kHRows rows in shared
// memory; each thread owns kSeg consecutive pixels of one row: one k-wide window sum, then it slides (add the
// entering column, subtract the leaving one — exact in float64 for sums of float32 values).  Rows are padded by one
// element per kSeg so that the threads of a warp hit different banks.  Epilogue selected by `op`:
//   OP_MEAN: out0 = mean | OP_MEAN_STD: out0 = mean, out1 = std = sqrt(max(m2 - mean^2, 0)) | OP_STD: out0 = std
enum { OP_MEAN = 0, OP_MEAN_STD = 1, OP_STD = 2 };
constexpr int kSeg = 8;
constexpr int kHThreads = 64;
constexpr int kHSpan = kHThreads * kSeg;                  // 512 pixels of output per block
constexpr int kHStage = kHSpan + kMaxK - 1;               // + halo
constexpr int kHPadded = kHStage + kHStage / kSeg + 1;    // one pad element every kSeg
__device__ __forceinline__ int hpad(int i) { return i + i / kSeg; }
__global__ void __launch_bounds__(kHThreads) box_h_kernel(const D4* __restrict__ S1, const D4* __restrict__ S2,
                                                          float4* __restrict__ out0, float4* __restrict__ out1, int h, int w, int k,
                                                          int op, int cols, int rows_per_block) {
  __shared__ D4 t1[kHPadded];
  __shared__ D4 t2[kHPadded];
  const int r = k / 2;
  const int b = blockIdx.z;
  const int row0 = blockIdx.y * rows_per_block;
  const int c0 = blockIdx.x * cols;                 // first output column of this block
  const int stage_w = cols + 2 * r;                 // staged columns per row
  // stage: rows_per_block rows x stage_w columns, row-major in the padded index space
  for (int i = threadIdx.x; i < rows_per_block * stage_w; i += blockDim.x) {
    const int rr = i / stage_w, cc = i - rr * stage_w;
    const int row = row0 + rr;
    if (row < h) {
      const int c = reflect101(c0 + cc - r, w);
      const size_t g = ((size_t)b * h + row) * w + c;
      t1[hpad(i)] = S1[g];
      if (S2) t2[hpad(i)] = S2[g];
    }
  }
  __syncthreads();
  const int segs_per_row = cols / kSeg;
  const int rr = threadIdx.x / segs_per_row, sg = threadIdx.x - rr * segs_per_row;
  const int row = row0 + rr;
  if (rr >= rows_per_block || row >= h) return;
  const int cfirst = c0 + sg * kSeg;
  if (cfirst >= w) return;
  const int base = rr * stage_w + sg * kSeg;        // staged index of the window start for the first pixel
  const double inv = 1.0 / ((double)k * (double)k);
  D4 a{0, 0, 0, 0}, q{0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const D4 v = t1[hpad(base + j)];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    if (S2) {
      const D4 u = t2[hpad(base + j)];
      q.x += u.x; q.y += u.y; q.z += u.z; q.w += u.w;
    }
  }
  const size_t rbase = ((size_t)b * h + row) * w;
  // stdfilt (isp_algos.py:236-241): float32 square of the blurred image, float32 difference, sqrt; explicit
  // round-to-nearest mul/sub: an FMA contraction would skip the float32 rounding of mean^2 the reference has
  auto sd = [](float e2, float e1) { return sqrtf(fmaxf(__fsub_rn(e2, __fmul_rn(e1, e1)), 0.f)); };
#pragma unroll
  for (int i = 0; i < kSeg; ++i) {
    const int c = cfirst + i;
    if (c < w) {
      const float4 m = make_float4((float)(a.x * inv), (float)(a.y * inv), (float)(a.z * inv), (float)(a.w * inv));
      if (op == OP_MEAN) {
        out0[rbase + c] = m;
      } else {
        const float4 m2 = make_float4((float)(q.x * inv), (float)(q.y * inv), (float)(q.z * inv), (float)(q.w * inv));
        const float4 st = make_float4(sd(m2.x, m.x), sd(m2.y, m.y), sd(m2.z, m.z), sd(m2.w, m.w));
        if (op == OP_MEAN_STD) {
          out0[rbase + c] = m;
          out1[rbase + c] = st;
        } else {
          out0[rbase + c] = st;
        }
      }
    }
    if (i + 1 < kSeg) {  // slide the window one pixel to the right
      const D4 vn = t1[hpad(base + i + k)], vo = t1[hpad(base + i)];
      a.x += vn.x - vo.x; a.y += vn.y - vo.y; a.z += vn.z - vo.z; a.w += vn.w - vo.w;
      if (S2) {
        const D4 un = t2[hpad(base + i + k)], uo = t2[hpad(base + i)];
        q.x += un.x - uo.x; q.y += un.y - uo.y; q.z += un.z - uo.z; q.w += un.w - uo.w;
      }
    }
  }
}

// var = std^2 (self)  |  var = std_lr^2 - std_hr^2 (collab), elementwise in float32 like the reference
__global__ void var_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ var, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float s = a[i];
    float v = __fmul_rn(s, s);  // no FMA contraction: the reference rounds each square to float32 first
    if (b) {
      const float t = b[i];
      v = __fsub_rn(v, __fmul_rn(t, t));
    }
    var[i] = v;
  }
}

// ------------------------------------------------------------------ exact order statistics (radix select)
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

constexpr int kMaxRanks = 64;
struct SelectWork {
  unsigned long long hist0[2048];
  unsigned long long hist1[kMaxRanks][2048];
  unsigned long long hist2[kMaxRanks][1024];
  int slot1_of_prefix[2048];             // top-11-bit prefix -> slot (or -1)
  int slot2_of[kMaxRanks][2048];         // (slot1, middle 11 bits) -> slot2 (or -1)
  unsigned long long rank_in[kMaxRanks];  // residual rank of each query inside its current bucket
  int q_slot1[kMaxRanks], q_slot2[kMaxRanks];
  uint32_t q_prefix[kMaxRanks];
  int nslot1, nslot2;
};

__global__ void __launch_bounds__(256) hist0_kernel(const float* __restrict__ d, size_t n, SelectWork* wk) {
  d += (size_t)blockIdx.y * n;  // one segment (image) per blockIdx.y
  wk += blockIdx.y;
  __shared__ unsigned int h[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(&h[f2key(d[i]) >> 21], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (h[i]) atomicAdd(&wk->hist0[i], (unsigned long long)h[i]);
}
// Warp-cooperative search of the bin holding rank r in a histogram: returns the bin, *before = #elements below it.
__device__ __forceinline__ int warp_find_bin(const unsigned long long* __restrict__ hist, int nbins, unsigned long long r,
                                             unsigned long long* before) {
  const int lane = threadIdx.x & 31;
  unsigned long long running = 0;
  for (int base = 0; base < nbins; base += 32) {
    const unsigned long long h = hist[base + lane];
    unsigned long long incl = h;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    if (r < running + total) {
      const unsigned int m = __ballot_sync(0xffffffffu, running + incl > r);
      const int l = __ffs(m) - 1;
      const unsigned long long excl = __shfl_sync(0xffffffffu, incl - h, l);
      *before = running + excl;
      return base + l;
    }
    running += total;
  }
  *before = running - hist[nbins - 1];
  return nbins - 1;
}

// One block of 32 warps; warp w resolves queries w, w+32.  Thread 0 then assigns slots (deduplicated buckets).
__global__ void __launch_bounds__(1024) select0_kernel(SelectWork* wk, const unsigned long long* __restrict__ ranks, int nranks) {
  wk += blockIdx.x;
  __shared__ int s_bin[kMaxRanks];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) wk->slot1_of_prefix[i] = -1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < nranks; q += 32) {
    unsigned long long before;
    const int bin = warp_find_bin(wk->hist0, 2048, ranks[q], &before);
    if (lane == 0) {
      s_bin[q] = bin;
      wk->q_prefix[q] = (uint32_t)bin;
      wk->rank_in[q] = ranks[q] - before;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ns = 0;
    for (int q = 0; q < nranks; ++q) {
      const int bin = s_bin[q];
      if (wk->slot1_of_prefix[bin] < 0) wk->slot1_of_prefix[bin] = ns++;
      wk->q_slot1[q] = wk->slot1_of_prefix[bin];
    }
    wk->nslot1 = ns;
  }
}
__global__ void __launch_bounds__(256) hist1_kernel(const float* __restrict__ d, size_t n, SelectWork* wk) {
  d += (size_t)blockIdx.y * n;
  wk += blockIdx.y;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = f2key(d[i]);
    const int s = wk->slot1_of_prefix[k >> 21];
    if (s >= 0) atomicAdd(&wk->hist1[s][(k >> 10) & 2047u], 1ull);
  }
}
__global__ void __launch_bounds__(1024) select1_kernel(SelectWork* wk, int nranks) {
  wk += blockIdx.x;
  __shared__ int s_bin[kMaxRanks];
  for (int i = threadIdx.x; i < kMaxRanks * 2048; i += blockDim.x) (&wk->slot2_of[0][0])[i] = -1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < nranks; q += 32) {
    unsigned long long before;
    const int bin = warp_find_bin(wk->hist1[wk->q_slot1[q]], 2048, wk->rank_in[q], &before);
    __syncwarp();
    if (lane == 0) {
      s_bin[q] = bin;
      wk->q_prefix[q] = (wk->q_prefix[q] << 11) | (uint32_t)bin;
      wk->rank_in[q] -= before;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ns = 0;
    for (int q = 0; q < nranks; ++q) {
      const int s1 = wk->q_slot1[q], bin = s_bin[q];
      if (wk->slot2_of[s1][bin] < 0) wk->slot2_of[s1][bin] = ns++;
      wk->q_slot2[q] = wk->slot2_of[s1][bin];
    }
    wk->nslot2 = ns;
  }
}
__global__ void __launch_bounds__(256) hist2_kernel(const float* __restrict__ d, size_t n, SelectWork* wk) {
  d += (size_t)blockIdx.y * n;
  wk += blockIdx.y;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = f2key(d[i]);
    const int s1 = wk->slot1_of_prefix[k >> 21];
    if (s1 < 0) continue;
    const int s2 = wk->slot2_of[s1][(k >> 10) & 2047u];
    if (s2 >= 0) atomicAdd(&wk->hist2[s2][k & 1023u], 1ull);
  }
}
__global__ void __launch_bounds__(1024) select2_kernel(SelectWork* wk, int nranks, float* __restrict__ out) {
  wk += blockIdx.x;
  out += (size_t)blockIdx.x * nranks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < nranks; q += 32) {
    unsigned long long before;
    const int bin = warp_find_bin(wk->hist2[wk->q_slot2[q]], 1024, wk->rank_in[q], &before);
    if (lane == 0) out[q] = key2f((wk->q_prefix[q] << 10) | (uint32_t)bin);
  }
}

// ------------------------------------------------------------------ score3 bin occupancy
// minj[bin] = smallest threshold index i such that some pixel with lap <= ths[i] falls in `bin`
__global__ void __launch_bounds__(256) score3_kernel(const float* __restrict__ lap, const float* __restrict__ mean, size_t n,
                                                     const double* __restrict__ ths, int nth, int* __restrict__ minj) {
  lap += (size_t)blockIdx.y * n;
  mean += (size_t)blockIdx.y * n;
  ths += (size_t)blockIdx.y * nth;
  minj += (size_t)blockIdx.y * 1001;
  __shared__ int smin[1001];
  __shared__ double sth[32];
  for (int i = threadIdx.x; i < 1001; i += blockDim.x) smin[i] = 0x7fffffff;
  if (threadIdx.x < nth) sth[threadIdx.x] = ths[threadIdx.x];
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double l = (double)lap[i];
    int j = 0;
    while (j < nth && !(l <= sth[j])) ++j;  // ths ascending: first threshold that admits this pixel
    if (j < nth) {
      const int bin = (int)(fminf(fmaxf(mean[i], 0.f), 1.f) * 1000.f);
      if (smin[bin] > j) atomicMin(&smin[bin], j);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1001; i += blockDim.x)
    if (smin[i] != 0x7fffffff) atomicMin(&minj[i], smin[i]);
}
__global__ void score3_count_kernel(const int* __restrict__ minj, int nth, int* __restrict__ npeaks) {
  minj += (size_t)blockIdx.x * 1001;
  npeaks += (size_t)blockIdx.x * nth;
  __shared__ int cnt[32];
  if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < 1001; b += blockDim.x) {
    const int j = minj[b];
    if (j < nth) atomicAdd(&cnt[j], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < nth; ++i) {
      c += cnt[i];
      npeaks[i] = c;
    }
  }
}
__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ masked regression sums
__global__ void __launch_bounds__(256) masked_sums_kernel(const float* __restrict__ lap, const float* __restrict__ mean,
                                                          const float* __restrict__ var, size_t n,
                                                          const double* __restrict__ ths, double* __restrict__ sums) {
  lap += (size_t)blockIdx.y * n;
  mean += (size_t)blockIdx.y * n;
  var += (size_t)blockIdx.y * n;
  sums += (size_t)blockIdx.y * 12;
  const double th = ths[blockIdx.y];
  double s[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) s[i] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if ((double)lap[i] < th) {
      const float xf = mean[i];
      const double x = xf, y = var[i];
      s[0] += 1.0; s[1] += x; s[2] += y; s[3] += x * x; s[4] += x * y; s[5] += y * y;
      if (xf > 1e-4f && xf < 0.8f) {
        s[6] += 1.0; s[7] += x; s[8] += y; s[9] += x * x; s[10] += x * y; s[11] += y * y;
      }
    }
  }
  __shared__ double red[8][12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    double v = s[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0;
    for (int wq = 0; wq < 8; ++wq) v += red[wq][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], v);
  }
}

inline int stream_grid(size_t n) {
  size_t g = (n + 256 * 8 - 1) / (256 * 8);
  const size_t cap = (size_t)yond_num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int box_pass(const float* x, float* out0, float* out1, int B, int h, int w, int k, int square_input, bool with_sq, int op,
             void* work, cudaStream_t s) {
  const size_t npix = (size_t)B * h * w;
  D4* S1 = reinterpret_cast<D4*>(work);
  D4* S2 = with_sq ? S1 + npix : nullptr;
  const int rows_per_strip = 64;
  dim3 gv(ceil_div(w, 128), ceil_div(h, rows_per_strip), B);
  box_v_kernel<<<gv, 128, 0, s>>>(reinterpret_cast<const float4*>(x), S1, S2, h, w, k, rows_per_strip, square_input);
  YOND_LAUNCH_CHECK();
  // block = `cols` output columns (multiple of kSeg, <= kHSpan) x as many rows as fit in kHThreads threads
  int cols = ceil_div(w, kSeg) * kSeg;
  if (cols > kHSpan) cols = kHSpan;
  int rpb = kHThreads / (cols / kSeg);
  while (rpb > 1 && rpb * (cols + k - 1) > kHStage) --rpb;
  dim3 gh(ceil_div(w, cols), ceil_div(h, rpb), B);
  box_h_kernel<<<gh, kHThreads, 0, s>>>(S1, S2, reinterpret_cast<float4*>(out0), reinterpret_cast<float4*>(out1), h, w, k, op, cols, rpb);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

}  // namespace

extern "C" {

size_t yond_nlf_work_bytes(int B, int h, int w, int C) {
  (void)C;
  const size_t npix = (size_t)B * h * w;
  // two float64x4 column-sum planes + two float32x4 temporaries (blur_k2(x), std of the second input)
  return 2 * npix * sizeof(D4) + 2 * npix * sizeof(float4) + 4096;
}

int yond_box_blur(const float* x, float* out, int B, int h, int w, int C, int k, int square_input, void* work, void* stream) {
  YOND_REQUIRE(C == 4, "yond_box_blur: packed 4-channel frames only (pass SIDD block stacks as a batch)");
  YOND_REQUIRE(k % 2 == 1 && k >= 1 && k <= kMaxK, "yond_box_blur: odd k <= %d required (got %d)", kMaxK, k);
  YOND_REQUIRE(h > k / 2 && w > k / 2, "yond_box_blur: frame smaller than the filter radius");
  return box_pass(x, out, nullptr, B, h, w, k, square_input, false, OP_MEAN, work, (cudaStream_t)stream);
}

int yond_nlf_maps(const float* x, const float* y, float* var, float* mean, float* lap, int B, int h, int w, int C, int k,
                  int mode, void* work, void* stream) {
  YOND_REQUIRE(C == 4, "yond_nlf_maps: packed 4-channel frames only (pass SIDD block stacks as a batch)");
  YOND_REQUIRE(k % 2 == 1 && k >= 3 && k <= kMaxK, "yond_nlf_maps: odd k <= %d required (got %d)", kMaxK, k);
  YOND_REQUIRE(h > k / 2 && w > k / 2, "yond_nlf_maps: frame smaller than the filter radius");
  YOND_REQUIRE(mode == 0 || (mode == 1 && y != nullptr), "yond_nlf_maps: collab mode needs the second input");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t npix = (size_t)B * h * w, n = npix * 4;
  uint8_t* wb = reinterpret_cast<uint8_t*>(work);
  float* tmpA = reinterpret_cast<float*>(wb + 2 * npix * sizeof(D4));
  float* tmpB = tmpA + n;
  int rc;
  if (mode == 0) {
    // mean = blur_k(x), std -> tmpA ; var = std^2
    if ((rc = box_pass(x, mean, tmpA, B, h, w, k, 0, true, OP_MEAN_STD, work, s))) return rc;
    var_kernel<<<stream_grid(n), 256, 0, s>>>(tmpA, nullptr, var, n);
    YOND_LAUNCH_CHECK();
    // lap = std_k(blur_k2(x)), k2 = k//3*2+1 (YOND_SIDD.py:70)
    const int k2 = k / 3 * 2 + 1;
    if ((rc = box_pass(x, tmpB, nullptr, B, h, w, k2, 0, false, OP_MEAN, work, s))) return rc;
    if ((rc = box_pass(tmpB, lap, nullptr, B, h, w, k, 0, true, OP_STD, work, s))) return rc;
  } else {
    // std_k(lr) -> tmpA ; mean = blur_k(hr), lap = std_k(hr) ; var = std_lr^2 - std_hr^2
    if ((rc = box_pass(x, tmpA, nullptr, B, h, w, k, 0, true, OP_STD, work, s))) return rc;
    if ((rc = box_pass(y, mean, lap, B, h, w, k, 0, true, OP_MEAN_STD, work, s))) return rc;
    var_kernel<<<stream_grid(n), 256, 0, s>>>(tmpA, lap, var, n);
    YOND_LAUNCH_CHECK();
  }
  return YOND_OK;
}

size_t yond_select_work_bytes(int nseg) { return (size_t)(nseg < 1 ? 1 : nseg) * sizeof(SelectWork) + 256; }

int yond_order_stats(const float* data, size_t seg_len, int nseg, const uint64_t* ranks_dev, int nranks, float* out_dev,
                     void* work, void* stream) {
  YOND_REQUIRE(nranks > 0 && nranks <= kMaxRanks, "yond_order_stats: 1..%d ranks (got %d)", kMaxRanks, nranks);
  YOND_REQUIRE(seg_len > 0 && nseg > 0 && nseg <= 65535, "yond_order_stats: empty input");
  cudaStream_t s = (cudaStream_t)stream;
  SelectWork* wk = reinterpret_cast<SelectWork*>(work);
  for (int i = 0; i < nseg; ++i)
    YOND_CUDA_CHECK(cudaMemsetAsync(wk + i, 0, offsetof(SelectWork, slot1_of_prefix), s));
  dim3 g(stream_grid(seg_len), nseg);
  if ((size_t)g.x * nseg > (size_t)yond_num_sms() * 16) g.x = (unsigned)((yond_num_sms() * 16 + nseg - 1) / nseg);
  hist0_kernel<<<g, 256, 0, s>>>(data, seg_len, wk);
  YOND_LAUNCH_CHECK();
  select0_kernel<<<nseg, 1024, 0, s>>>(wk, reinterpret_cast<const unsigned long long*>(ranks_dev), nranks);
  YOND_LAUNCH_CHECK();
  hist1_kernel<<<g, 256, 0, s>>>(data, seg_len, wk);
  YOND_LAUNCH_CHECK();
  select1_kernel<<<nseg, 1024, 0, s>>>(wk, nranks);
  YOND_LAUNCH_CHECK();
  hist2_kernel<<<g, 256, 0, s>>>(data, seg_len, wk);
  YOND_LAUNCH_CHECK();
  select2_kernel<<<nseg, 1024, 0, s>>>(wk, nranks, out_dev);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

int yond_score3_bins(const float* lap, const float* mean, size_t seg_len, int nseg, const double* ths_dev, int nth,
                     int32_t* npeaks_dev, void* work, void* stream) {
  YOND_REQUIRE(nth > 0 && nth <= 32, "yond_score3_bins: 1..32 thresholds (got %d)", nth);
  YOND_REQUIRE(nseg > 0 && nseg <= 65535, "yond_score3_bins: bad segment count");
  cudaStream_t s = (cudaStream_t)stream;
  int* minj = reinterpret_cast<int*>(work);
  fill_int_kernel<<<ceil_div(1001 * nseg, 256), 256, 0, s>>>(minj, 1001 * nseg, 0x7fffffff);
  YOND_LAUNCH_CHECK();
  dim3 g(stream_grid(seg_len), nseg);
  if ((size_t)g.x * nseg > (size_t)yond_num_sms() * 16) g.x = (unsigned)((yond_num_sms() * 16 + nseg - 1) / nseg);
  score3_kernel<<<g, 256, 0, s>>>(lap, mean, seg_len, ths_dev, nth, minj);
  YOND_LAUNCH_CHECK();
  score3_count_kernel<<<nseg, 256, 0, s>>>(minj, nth, npeaks_dev);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

int yond_masked_sums(const float* lap, const float* mean, const float* var, size_t seg_len, int nseg, const double* ths_dev,
                     double* sums_dev, void* stream) {
  YOND_REQUIRE(nseg > 0 && nseg <= 65535, "yond_masked_sums: bad segment count");
  cudaStream_t s = (cudaStream_t)stream;
  YOND_CUDA_CHECK(cudaMemsetAsync(sums_dev, 0, (size_t)nseg * 12 * sizeof(double), s));
  dim3 g(stream_grid(seg_len), nseg);
  if ((size_t)g.x * nseg > (size_t)yond_num_sms() * 16) g.x = (unsigned)((yond_num_sms() * 16 + nseg - 1) / nseg);
  masked_sums_kernel<<<g, 256, 0, s>>>(lap, mean, var, seg_len, ths_dev, sums_dev);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

}  // extern "C"
