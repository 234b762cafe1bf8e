// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M=128, bf16) issued back-to-back from
// resident shared memory, for the two operand layouts (no-swizzle "interleaved" K-major vs 128B
// swizzle K-major), several N, 1/2/4 independent accumulators and a misaligned A start (tap shift).
// Build + run on the GPU box:  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 tools/umma_bench.cu -o /tmp/ub && /tmp/ub
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

struct Cfg {
  int swizzle;   // 0 none, 1 = 128B
  int N;
  int nacc;      // independent accumulators used round-robin
  int ashift;    // A start row shift (rows); no-swizzle only
  int iters;
};

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // zero the operands (values do not matter for timing; keep them finite)
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 96 * 1024);
    uint64_t ahi, bhi;
    uint32_t a_lo, b_lo, kstep_a, kstep_b;
    if (c.swizzle == 0) {
      // [chunk][row][16B]: LBO = 312*16 (A) / N*16 (B), SBO = 128
      const uint32_t lbo_a = 312 * 16, lbo_b = (uint32_t)c.N * 16;
      uint64_t da = ((uint64_t)(lbo_a >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
      uint64_t db = ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
      ahi = da >> 32; bhi = db >> 32;
      a_lo = (uint32_t)da + (abase >> 4) + (uint32_t)c.ashift;
      b_lo = (uint32_t)db + (bbase >> 4);
      kstep_a = (2 * lbo_a) >> 4; kstep_b = (2 * lbo_b) >> 4;
    } else {
      // 128B swizzle K-major: rows of 128 B (64 bf16), 8-row atoms of 1024 B: SBO = 1024, LBO unused (1)
      uint64_t d = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      ahi = d >> 32; bhi = d >> 32;
      a_lo = (uint32_t)d + (abase >> 4);
      b_lo = (uint32_t)d + (bbase >> 4);
      kstep_a = 32 >> 4; kstep_b = 32 >> 4;   // +16 elements of K inside the 128-byte row
    }
    const int stride_cols = 512 / c.nacc;
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      // one "k-block": 4 k16 steps on each accumulator, accumulators alternating in the inner loop
      uint32_t al = a_lo, bl = b_lo;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        for (int j = 0; j < c.nacc; ++j)
          umma(tmem + (uint32_t)(j * stride_cols), ((uint64_t)ahi << 32) | (al + (c.swizzle ? 0u : (uint32_t)j * 128u)),
               ((uint64_t)bhi << 32) | bl, idesc, (uint32_t)(it | k));
        al += kstep_a; bl += kstep_b;
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Variant 2: W issuing warps (lane 0 of each), each with its own accumulator; descriptors precomputed,
// predicate fixed, inner loop = 4 bare tcgen05.mma instructions.
__device__ __forceinline__ void umma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc)
               : "memory");
}
__global__ void __launch_bounds__(128, 1) bench2(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  long long dt = 0;
  if (lane == 0 && warp < c.nacc) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 96 * 1024);
    const uint32_t lbo_a = 312 * 16, lbo_b = (uint32_t)c.N * 16;
    uint64_t da = ((uint64_t)(lbo_a >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    uint64_t db = ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    uint64_t a[4], b[4];
    for (int k = 0; k < 4; ++k) {
      a[k] = da + (abase >> 4) + (uint32_t)c.ashift + (uint32_t)warp * 128u + (uint32_t)k * ((2 * lbo_a) >> 4);
      b[k] = db + (bbase >> 4) + (uint32_t)k * ((2 * lbo_b) >> 4);
    }
    const uint32_t d = tmem + (uint32_t)(warp * (512 / c.nacc));
    long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      umma_acc(d, a[0], b[0], idesc); umma_acc(d, a[1], b[1], idesc);
      umma_acc(d, a[2], b[2], idesc); umma_acc(d, a[3], b[3], idesc);
    }
    umma_commit(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    dt = clock64() - t0;
    atomicMax((unsigned long long*)&out[blockIdx.x], (unsigned long long)dt);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ns[] = {32, 64, 96, 128, 256};
  printf("swz    N nacc shift   cyc/MMA  floor  (148 CTAs, 1 issuing thread each)\n");
  for (int sw = 0; sw < 2; ++sw)
    for (int N : Ns)
      for (int nacc : {1, 2, 4})
        for (int shift : {0, 3}) {
          if (sw == 1 && shift) continue;
          if (nacc * N > 512) continue;
          Cfg c{sw, N, nacc, shift, 2000};
          bench<<<148, 128, 200 * 1024>>>(c, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
          double mx = 0;
          for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("%3d %4d %4d %5d  %8.1f  %5.0f\n", sw, N, nacc, shift, mx / (c.iters * 4.0 * nacc), N / 2.0);
        }
  printf("\nvariant 2: W issuing warps, bare MMA loop\n   N warps   cyc/MMA(all warps)  floor\n");
  for (int N : Ns)
    for (int nacc : {1, 2, 4}) {
      if (nacc * N > 512) continue;
      Cfg c{0, N, nacc, 3, 2000};
      cudaMemset(out, 0, 148 * sizeof(long long));
      bench2<<<148, 128, 200 * 1024>>>(c, out);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%4d %5d  %8.1f  %5.0f\n", N, nacc, mx / (c.iters * 4.0 * nacc), N / 2.0);
    }
  return 0;
}
