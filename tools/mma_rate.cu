// Bring-up microbenchmark: tcgen05.mma issue / execution rate on B200 for M=128, K=16, various N and accumulator
// patterns.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu && ./mma_rate
// Operands are uninitialised shared memory (values are irrelevant for timing).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo = 1024u, uint64_t layout = 2ull) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}

// pattern: 0 = one accumulator, 1 = rotate over 4 accumulators, 2 = one accumulator + descriptor low-word increments,
//          3 = A operand from TMEM, one accumulator, 4 = A from TMEM, rotate 4 accumulators
template <int N, int PATTERN>
__global__ void __launch_bounds__(320, 1) rate_kernel(long long* out, int iters, int spin, uint32_t inc) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  if (threadIdx.x < 32) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = desc(base), b0 = desc(base + 65536);
    const uint64_t a128 = desc(base, 1280u, 2ull), a64 = desc(base, 640u, 4ull), b64 = desc(base + 65536, 512u, 4ull);
    long long t0 = clock64();
    uint64_t ra = a64, rb = b64;  // running descriptors (pattern 12/13): every MMA depends on a fresh uniform add
    if (leader) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (PATTERN == 0) mma(tm, a0, b0, idesc, 1u);
          if (PATTERN == 1) mma(tm + (u & 3) * 64, a0, b0, idesc, 1u);
          if (PATTERN == 2) mma(tm, a0 + 2 * (u & 3), b0 + 2 * (u & 3), idesc, 1u);
          if (PATTERN == 3) mma_ts(tm, tm + 256 + 8 * (u & 3), b0, idesc, 1u);
          if (PATTERN == 4) mma_ts(tm + (u & 3) * 64, tm + 256 + 8 * (u & 3), b0, idesc, 1u);
          // conv-style A operands: halo slab (pitch 10 pixels), start shifted by whole pixels, K advance inside the row
          if (PATTERN == 5) mma(tm, a128 + 8 * ((u >> 2) * 11) + 2 * (u & 3), b0 + 2 * (u & 3), idesc, 1u);   // SW128, sbo 1280
          if (PATTERN == 6) mma(tm, a64 + 4 * ((u >> 1) * 11) + 2 * (u & 1), b64 + 2 * (u & 1), idesc, 1u);    // SW64, sbo 640
          if (PATTERN == 7) mma(tm, a128 + 2 * (u & 3), b0 + 2 * (u & 3), idesc, 1u);                          // SW128, sbo 1280, aligned start
          if (PATTERN == 8) mma(tm, a0 + 8 * (u >> 2) + 2 * (u & 3), b0 + 2 * (u & 3), idesc, 1u);             // sbo 1024, start shifted by rows
          if (PATTERN == 12) { mma(tm, ra, rb, idesc, 1u); ra += inc; rb += inc; if ((u & 3) == 3) { ra -= 4 * inc; rb -= 4 * inc; } }
          if (PATTERN == 13) { mma(tm + (u & 3) * 32, ra, rb, idesc, 1u); ra += inc; if ((u & 3) == 3) { ra -= 4 * inc; rb += 2; } if (u == 15) rb -= 8; }
          if (PATTERN >= 9 && PATTERN <= 11) {  // bursts of 8 MMAs followed by commit (9), fence (10), fresh accumulation (11)
            mma(tm, a64 + 4 * ((u >> 1) * 11) + 2 * (u & 1), b64 + 2 * (u & 1), idesc, (PATTERN == 11 && (u & 7) == 0) ? 0u : 1u);
            if ((u & 7) == 7) {
              if (PATTERN == 9) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
              if (PATTERN == 10) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    long long t_issue = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = t_issue - t0;
      out[1] = t1 - t0;
    }
  }
  else if (spin) {  // the other warps wait on the same mbarrier, like the epilogue / producer warps of the conv kernel
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

template <int N, int P>
void run(const char* name, int grid, int spin = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 256;
  cudaFuncSetAttribute(rate_kernel<N, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<N, P><<<grid, 320, 200 * 1024>>>(d, iters, spin, 2u);
  rate_kernel<N, P><<<grid, 320, 200 * 1024>>>(d, iters, spin, 2u);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA%s\n", name, N, grid, (double)h[0] / (iters * 16),
         (double)h[1] / (iters * 16), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<32, 0>("SS one accumulator", grid);
    run<32, 1>("SS rotate 4 accumulators", grid);
    run<32, 2>("SS one accumulator, advancing descriptors", grid);
    run<32, 3>("TS (A in TMEM) one accumulator", grid);
    run<32, 4>("TS rotate 4 accumulators", grid);
    run<64, 0>("SS one accumulator", grid);
    run<64, 1>("SS rotate 4 accumulators", grid);
    run<64, 3>("TS one accumulator", grid);
    run<128, 0>("SS one accumulator", grid);
    run<128, 3>("TS one accumulator", grid);
    run<256, 0>("SS one accumulator", grid);
    run<256, 3>("TS one accumulator", grid);
    run<32, 6>("SS SW64 slab (sbo 640, pixel-shifted starts)", grid);
    run<64, 5>("SS SW128 slab (sbo 1280, pixel-shifted)", grid);
    run<64, 7>("SS SW128 sbo 1280, aligned start", grid);
    run<64, 8>("SS SW128 sbo 1024, row-shifted start", grid);
    run<128, 5>("SS SW128 slab (sbo 1280, pixel-shifted)", grid);
    run<32, 12>("SS SW64 running descriptors (dependent adds)", grid);
    run<32, 13>("SS SW64 running descs, 4 accumulators", grid);
    run<32, 9>("SS SW64 slab, commit every 8 MMAs", grid);
    run<32, 10>("SS SW64 slab, fence::after every 8 MMAs", grid);
    run<32, 11>("SS SW64 slab, accumulate=0 every 8 MMAs", grid);
    run<32, 6>("SS SW64 slab + 9 warps in mbarrier.try_wait", grid, 1);
    run<64, 5>("SS SW128 slab + 9 warps in mbarrier.try_wait", grid, 1);
  }
  return 0;
}
