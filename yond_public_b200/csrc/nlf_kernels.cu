// Noise-level-function estimation kernels (SimpleNLF): box statistics, exact order statistics, the
// 'score3' bin-occupancy count and the masked regression sums.
//   reference: utils/isp_algos.py:234-242 (stdfilt via cv2.blur), :345-365 (polyfit);
//              YOND_SIDD.py:22-49 (get_threshold 'score3'), :62-115 (SelfNLF / CollabNLF).
// cv2.blur on float32 = normalised box, BORDER_REFLECT_101, float64 running sums, result rounded to float32;
// the box kernel below keeps float64 sums (vertical sliding window in registers, horizontal prefix in shared memory).
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kMaxK = 31;

__device__ __forceinline__ float4 sq4(const float4& v) {
  return make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w));
}

// Fused box statistics: one kernel, no float64 scratch in HBM.
// A block owns kCols consecutive columns (output columns + a halo of k/2 on each side) of a strip of rows.  Thread c
// slides a k-tall window down column c (float64 sums of x and of fl32(x*x), 4 channels = NQ running sums).  For each
// row the k-wide horizontal sums come from an inclusive prefix over the block's columns, window(c) = P[c+r] - P[c-r-1],
// computed in shared memory as a chunked scan: warp q owns quantity q, lane j scans the CH = kCols/32 consecutive
// columns of chunk j sequentially, one warp shuffle-scan of the 32 chunk totals gives the offsets.  That is one
// 5-step shuffle scan per quantity and row instead of one per quantity and WARP OF COLUMNS (8x fewer shuffles, ~3x
// fewer instructions than scanning the column sums where they live).  All sums are float64 sums of float32 values,
// so the result equals cv2.blur's float64 accumulation rounded to float32 up to the (far below float32) float64
// rounding of the summation order.  S is double-buffered by row parity: two __syncthreads per row.
// OP_MEAN_VAR: out0 = mean, out1 = std^2 (SelfNLF's var).  OP_COLLAB: out0 = mean, out1 = std (lap), out2 = aux^2 - std^2
// with aux = the std map of the other frame (CollabNLF's var).  Squares / difference in float32 with explicit rounding,
// like the reference's elementwise float32 expressions.
enum { OP_MEAN = 0, OP_MEAN_STD = 1, OP_STD = 2, OP_MEAN_VAR = 3, OP_COLLAB = 4 };
constexpr int kBoxThreads = 256;
template <bool kSq, int kCols>  // kCols = 256, or 160 when the whole row plus both halos fits (SIDD blocks: 128 + 28)
__global__ void __launch_bounds__(kBoxThreads) box_fused_kernel(const float4* __restrict__ x, float4* __restrict__ out0,
                                                                float4* __restrict__ out1, int h, int w, int k, int op,
                                                                int rows_per_strip, const float4* __restrict__ aux,
                                                                float4* __restrict__ out2) {
  constexpr int NQ = kSq ? 8 : 4;
  constexpr int CH = kCols / 32;           // columns per scan chunk (8 or 5)
  constexpr int PITCH = kCols + 32 + 1;    // padded: column c sits at c + c / CH, so chunk reads are bank-conflict free
  __shared__ double S[2][NQ][PITCH];
  const int r = k / 2;
  const int outc = kCols - 2 * r;
  const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
  const bool colthread = c < kCols;
  const int b = blockIdx.z;
  const int col_out = blockIdx.x * outc + (c - r);          // image column this thread's window is centred on
  const int col_src = reflect101(col_out, w);               // BORDER_REFLECT_101
  const int i0 = blockIdx.y * rows_per_strip;
  const int i1 = min(h, i0 + rows_per_strip);
  const float4* xb = x + (size_t)b * h * w;
  double s[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) s[q] = 0.0;
  auto accum = [&](const float4& v, double sign) {
    s[0] += sign * v.x; s[1] += sign * v.y; s[2] += sign * v.z; s[3] += sign * v.w;
    if (kSq) {
      const float4 q2 = sq4(v);
      s[4] += sign * q2.x; s[5] += sign * q2.y; s[6] += sign * q2.z; s[7] += sign * q2.w;
    }
  };
  auto load = [&](int i) { return __ldg(xb + (size_t)reflect101(i, h) * w + col_src); };
  if (colthread)
    for (int i = i0 - r; i <= i0 + r; ++i) accum(load(i), 1.0);
  const double inv = 1.0 / ((double)k * (double)k);
  // stdfilt (isp_algos.py:236-241): float32 square of the blurred image, float32 difference, sqrt; explicit
  // round-to-nearest mul/sub: an FMA contraction would skip the float32 rounding of mean^2 the reference has
  auto sd = [](float e2, float e1) { return sqrtf(fmaxf(__fsub_rn(e2, __fmul_rn(e1, e1)), 0.f)); };
  const bool writer = (c >= r) && (c < r + outc) && (col_out < w);
  const int my_idx = c + c / CH;
  const int hi_idx = (c + r) + (c + r) / CH;
  const int lo_col = c - r - 1;
  const int lo_idx = lo_col >= 0 ? lo_col + lo_col / CH : 0;
  for (int i = i0; i < i1; ++i) {
    double(*Sb)[PITCH] = S[i & 1];
    // next row's two pixels: in flight while this row is scanned
    float4 vin = make_float4(0.f, 0.f, 0.f, 0.f), vout = vin;
    const bool more = colthread && (i + 1 < i1);
    if (more) {
      vin = load(i + 1 + r);
      vout = load(i - r);
    }
    if (colthread) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) Sb[q][my_idx] = s[q];
    }
    __syncthreads();
    if (warp < NQ) {  // chunked inclusive scan of quantity `warp` over the block's columns
      double* row = Sb[warp] + lane * (CH + 1);  // chunk `lane` starts at column lane*CH, i.e. index lane*CH + lane
      double loc[CH];
      double run = 0.0;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        run += row[j];
        loc[j] = run;
      }
      double incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const double excl = incl - run;
#pragma unroll
      for (int j = 0; j < CH; ++j) row[j] = loc[j] + excl;
    }
    __syncthreads();
    if (writer) {
      double a[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) a[q] = Sb[q][hi_idx] - (lo_col >= 0 ? Sb[q][lo_idx] : 0.0);
      const size_t o = ((size_t)b * h + i) * w + col_out;
      const float4 m = make_float4((float)(a[0] * inv), (float)(a[1] * inv), (float)(a[2] * inv), (float)(a[3] * inv));
      if (!kSq || op == OP_MEAN) {
        out0[o] = m;
      } else {
        const float4 m2 = make_float4((float)(a[4] * inv), (float)(a[5] * inv), (float)(a[6] * inv), (float)(a[7] * inv));
        const float4 st = make_float4(sd(m2.x, m.x), sd(m2.y, m.y), sd(m2.z, m.z), sd(m2.w, m.w));
        if (op == OP_MEAN_STD) {
          out0[o] = m;
          out1[o] = st;
        } else if (op == OP_MEAN_VAR) {
          out0[o] = m;
          out1[o] = sq4(st);
        } else if (op == OP_COLLAB) {
          out0[o] = m;
          out1[o] = st;
          const float4 a = __ldg(aux + o), a2 = sq4(a), s2 = sq4(st);
          out2[o] = make_float4(__fsub_rn(a2.x, s2.x), __fsub_rn(a2.y, s2.y), __fsub_rn(a2.z, s2.z), __fsub_rn(a2.w, s2.w));
        } else {
          out0[o] = st;
        }
      }
    }
    if (more) {
      accum(vin, 1.0);
      accum(vout, -1.0);
    }
    // the other S buffer is written next; this one is rewritten two rows later, after the next row's barriers
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
  const float4* d4 = reinterpret_cast<const float4*>(d);  // segments hold 4-channel pixels: n % 4 == 0, 16-byte aligned
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream_f4(d4 + i);
    atomicAdd(&h[f2key(v.x) >> 21], 1u);
    atomicAdd(&h[f2key(v.y) >> 21], 1u);
    atomicAdd(&h[f2key(v.z) >> 21], 1u);
    atomicAdd(&h[f2key(v.w) >> 21], 1u);
  }
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
// Second radix level.  The queried ranks (percentiles 5..95) sit in a handful of top-11-bit buckets that together
// hold most of the data, so nearly every element increments a counter: the counters of the first kHist1Slots buckets
// are privatised in shared memory (32-bit, flushed once per block), later buckets fall back to global atomics.
constexpr int kHist1Slots = 16;
constexpr int kHist1Threads = 1024;
__global__ void __launch_bounds__(kHist1Threads) hist1_kernel(const float* __restrict__ d, size_t n, SelectWork* wk) {
  extern __shared__ unsigned int hs[];  // [kHist1Slots][2048]
  d += (size_t)blockIdx.y * n;
  wk += blockIdx.y;
  const int nsh = min(wk->nslot1, kHist1Slots);
  for (int i = threadIdx.x; i < nsh * 2048; i += blockDim.x) hs[i] = 0u;
  __syncthreads();
  const float4* d4 = reinterpret_cast<const float4*>(d);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream_f4(d4 + i);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t k = f2key(e[j]);
      const int s = __ldg(&wk->slot1_of_prefix[k >> 21]);
      if (s >= 0) {
        const uint32_t mid = (k >> 10) & 2047u;
        if (s < kHist1Slots) atomicAdd(&hs[s * 2048 + mid], 1u);
        else atomicAdd(&wk->hist1[s][mid], 1ull);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nsh * 2048; i += blockDim.x)
    if (hs[i]) atomicAdd(&wk->hist1[i >> 11][i & 2047], (unsigned long long)hs[i]);
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
  const float4* d4 = reinterpret_cast<const float4*>(d);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream_f4(d4 + i);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t k = f2key(e[j]);
      const int s1 = wk->slot1_of_prefix[k >> 21];
      if (s1 < 0) continue;
      const int s2 = wk->slot2_of[s1][(k >> 10) & 2047u];
      if (s2 >= 0) atomicAdd(&wk->hist2[s2][k & 1023u], 1ull);
    }
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
  __shared__ float sthf[32];
  for (int i = threadIdx.x; i < 1001; i += blockDim.x) smin[i] = 0x7fffffff;
  if (threadIdx.x < 32) {
    // lap is float32: (double)l <= t  <=>  l <= the largest float32 not above t, so the comparison runs in float32
    float f = __int_as_float(0x7f800000);  // +inf pads the table
    if ((int)threadIdx.x < nth) {
      const double t = ths[threadIdx.x];
      f = __double2float_rn(t);
      if ((double)f > t) f = nextafterf(f, -__int_as_float(0x7f800000));
    }
    sthf[threadIdx.x] = f;
  }
  __syncthreads();
  const float4* l4 = reinterpret_cast<const float4*>(lap);
  const float4* m4 = reinterpret_cast<const float4*>(mean);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 lv = ldg_stream_f4(l4 + i), mv = ldg_stream_f4(m4 + i);
    const float le[4] = {lv.x, lv.y, lv.z, lv.w}, me[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float l = le[e];
      // ths ascending: j = number of thresholds that do not admit this pixel = index of the first that does
      int j = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const int m = j + step;
        if (m <= nth && !(l <= sthf[m - 1])) j = m;
      }
      if (j < nth) {
        const int bin = (int)__fmul_rn(fminf(fmaxf(me[e], 0.f), 1.f), 1000.f);
        if (smin[bin] > j) atomicMin(&smin[bin], j);
      }
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
  const float4* l4 = reinterpret_cast<const float4*>(lap);
  const float4* m4 = reinterpret_cast<const float4*>(mean);
  const float4* v4 = reinterpret_cast<const float4*>(var);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 lv = ldg_stream_f4(l4 + i), mv = ldg_stream_f4(m4 + i), vv = ldg_stream_f4(v4 + i);
    const float le[4] = {lv.x, lv.y, lv.z, lv.w}, me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if ((double)le[e] < th) {
        const float xf = me[e];
        const double x = xf, y = ve[e];
        s[0] += 1.0; s[1] += x; s[2] += y; s[3] += x * x; s[4] += x * y; s[5] += y * y;
        if (xf > 1e-4f && xf < 0.8f) {
          s[6] += 1.0; s[7] += x; s[8] += y; s[9] += x * x; s[10] += x * y; s[11] += y * y;
        }
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
  size_t g = (n + 256 * 16 - 1) / (256 * 16);
  const size_t cap = (size_t)yond_num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int box_pass(const float* x, float* out0, float* out1, int B, int h, int w, int k, int square_input, bool with_sq, int op,
             void* work, cudaStream_t s, const float* aux = nullptr, float* out2 = nullptr) {
  (void)work;
  (void)square_input;
  const bool narrow = w + 2 * (k / 2) <= 160;
  static const int env_rows = getenv("YOND_BOX_ROWS") ? atoi(getenv("YOND_BOX_ROWS")) : 0;
  const int rows_per_strip = env_rows > 0 ? env_rows : 64;
  const int cols = narrow ? 160 : 256;
  const int outc = cols - 2 * (k / 2);
  dim3 g(ceil_div(w, outc), ceil_div(h, rows_per_strip), B);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* o0 = reinterpret_cast<float4*>(out0);
  float4* o1 = reinterpret_cast<float4*>(out1);
  const float4* ax = reinterpret_cast<const float4*>(aux);
  float4* o2 = reinterpret_cast<float4*>(out2);
  if (with_sq) {
    if (narrow) box_fused_kernel<true, 160><<<g, kBoxThreads, 0, s>>>(x4, o0, o1, h, w, k, op, rows_per_strip, ax, o2);
    else box_fused_kernel<true, 256><<<g, kBoxThreads, 0, s>>>(x4, o0, o1, h, w, k, op, rows_per_strip, ax, o2);
  } else {
    if (narrow) box_fused_kernel<false, 160><<<g, kBoxThreads, 0, s>>>(x4, o0, o1, h, w, k, OP_MEAN, rows_per_strip, ax, o2);
    else box_fused_kernel<false, 256><<<g, kBoxThreads, 0, s>>>(x4, o0, o1, h, w, k, OP_MEAN, rows_per_strip, ax, o2);
  }
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

}  // namespace

extern "C" {

size_t yond_nlf_work_bytes(int B, int h, int w, int C) {
  (void)C;
  const size_t npix = (size_t)B * h * w;
  // two float32x4 temporaries (blur_k2(x) / std of the first input)
  return 2 * npix * sizeof(float4) + 4096;
}

int yond_box_blur(const float* x, float* out, int B, int h, int w, int C, int k, int square_input, void* work, void* stream) {
  YOND_REQUIRE(C == 4, "yond_box_blur: packed 4-channel frames only (pass SIDD block stacks as a batch)");
  YOND_REQUIRE(k % 2 == 1 && k >= 1 && k <= kMaxK, "yond_box_blur: odd k <= %d required (got %d)", kMaxK, k);
  YOND_REQUIRE(h > k / 2 && w > k / 2, "yond_box_blur: frame smaller than the filter radius");
  YOND_REQUIRE(square_input == 0, "yond_box_blur: square_input is not supported");
  return box_pass(x, out, nullptr, B, h, w, k, 0, false, OP_MEAN, work, (cudaStream_t)stream);
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
  float* tmpA = reinterpret_cast<float*>(wb);
  float* tmpB = tmpA + n;
  int rc;
  if (mode == 0) {
    // mean = blur_k(x), var = std_k(x)^2 in one pass
    if ((rc = box_pass(x, mean, var, B, h, w, k, 0, true, OP_MEAN_VAR, work, s))) return rc;
    // lap = std_k(blur_k2(x)), k2 = k//3*2+1 (YOND_SIDD.py:70)
    const int k2 = k / 3 * 2 + 1;
    if ((rc = box_pass(x, tmpB, nullptr, B, h, w, k2, 0, false, OP_MEAN, work, s))) return rc;
    if ((rc = box_pass(tmpB, lap, nullptr, B, h, w, k, 0, true, OP_STD, work, s))) return rc;
  } else {
    // std_k(lr) -> tmpA ; mean = blur_k(hr), lap = std_k(hr) ; var = std_lr^2 - std_hr^2
    if ((rc = box_pass(x, tmpA, nullptr, B, h, w, k, 0, true, OP_STD, work, s))) return rc;
    if ((rc = box_pass(y, mean, lap, B, h, w, k, 0, true, OP_COLLAB, work, s, tmpA, var))) return rc;
  }
  return YOND_OK;
}

size_t yond_select_work_bytes(int nseg) { return (size_t)(nseg < 1 ? 1 : nseg) * sizeof(SelectWork) + 256; }

int yond_order_stats(const float* data, size_t seg_len, int nseg, const uint64_t* ranks_dev, int nranks, float* out_dev,
                     void* work, void* stream) {
  YOND_REQUIRE(nranks > 0 && nranks <= kMaxRanks, "yond_order_stats: 1..%d ranks (got %d)", kMaxRanks, nranks);
  YOND_REQUIRE(seg_len > 0 && nseg > 0 && nseg <= 65535, "yond_order_stats: empty input");
  YOND_REQUIRE(seg_len % 4 == 0 && (uintptr_t)data % 16 == 0, "yond_order_stats: segments must hold whole 4-channel pixels (16-byte aligned)");
  cudaStream_t s = (cudaStream_t)stream;
  SelectWork* wk = reinterpret_cast<SelectWork*>(work);
  // one memset for all segments (the slot tables behind the histograms are rewritten by the select kernels anyway)
  YOND_CUDA_CHECK(cudaMemsetAsync(wk, 0, (size_t)nseg * sizeof(SelectWork), s));
  dim3 g(stream_grid(seg_len), nseg);
  if ((size_t)g.x * nseg > (size_t)yond_num_sms() * 16) g.x = (unsigned)((yond_num_sms() * 16 + nseg - 1) / nseg);
  hist0_kernel<<<g, 256, 0, s>>>(data, seg_len, wk);
  YOND_LAUNCH_CHECK();
  select0_kernel<<<nseg, 1024, 0, s>>>(wk, reinterpret_cast<const unsigned long long*>(ranks_dev), nranks);
  YOND_LAUNCH_CHECK();
  {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    const size_t smem = (size_t)kHist1Slots * 2048 * sizeof(unsigned int);
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(hist1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (attr_err != cudaSuccess) return yond_set_error(YOND_ERR_CUDA, "cudaFuncSetAttribute(hist1_kernel) failed: %s", cudaGetErrorString(attr_err));
    int bx = yond_num_sms() / nseg;  // one 128 KB block per SM
    if (bx < 1) bx = 1;
    const size_t need = (seg_len / 4 + kHist1Threads - 1) / kHist1Threads;
    if ((size_t)bx > need) bx = (int)need;
    hist1_kernel<<<dim3(bx, nseg), kHist1Threads, smem, s>>>(data, seg_len, wk);
    YOND_LAUNCH_CHECK();
  }
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
  YOND_REQUIRE(seg_len % 4 == 0 && (uintptr_t)lap % 16 == 0 && (uintptr_t)mean % 16 == 0, "yond_score3_bins: 4-channel pixel segments required");
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
  YOND_REQUIRE(seg_len % 4 == 0, "yond_masked_sums: 4-channel pixel segments required");
  cudaStream_t s = (cudaStream_t)stream;
  YOND_CUDA_CHECK(cudaMemsetAsync(sums_dev, 0, (size_t)nseg * 12 * sizeof(double), s));
  dim3 g(stream_grid(seg_len), nseg);
  if ((size_t)g.x * nseg > (size_t)yond_num_sms() * 16) g.x = (unsigned)((yond_num_sms() * 16 + nseg - 1) / nseg);
  masked_sums_kernel<<<g, 256, 0, s>>>(lap, mean, var, seg_len, ths_dev, sums_dev);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

}  // extern "C"
