// Micro-benchmark 2: the conv kernel's MMA issue pattern in isolation (one elected thread, resident
// operands): per "tile", ntaps taps x nk16 K-steps x msub sub-tile accumulators, then a commit.
// Answers: what does one M=128 x N x K=16 MMA cost in this pattern, and does the loop order matter?
// Build + run on the GPU box:
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 tools/umma_bench2.cu -o /tmp/ub2 && /tmp/ub2
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}

struct Cfg { int N, msub, nk16, ntaps, order, iters, M; };

template <int MS>
__device__ __forceinline__ void issue_tap(uint32_t d0, uint32_t bnc, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accum, int nk16, uint32_t a_kstep, uint32_t b_kstep) {
#pragma unroll 4
  for (int k16 = 0; k16 < nk16; ++k16) {
    const uint32_t af = accum | (uint32_t)k16;
    umma(d0, adesc, bdesc, idesc, af);
    if (MS >= 2) umma(d0 + bnc, adesc + 128u, bdesc, idesc, af);
    if (MS == 4) {
      umma(d0 + 2 * bnc, adesc + 256u, bdesc, idesc, af);
      umma(d0 + 3 * bnc, adesc + 384u, bdesc, idesc, af);
    }
    adesc += a_kstep; bdesc += b_kstep;
  }
}

__global__ void __launch_bounds__(128, 1) bench3(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (200 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
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
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
      const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 120 * 1024);
      const int astride = 128 * c.msub + 56;
      const uint32_t lbo_a = (uint32_t)astride * 16, lbo_b = (uint32_t)c.N * 16;
      const uint64_t da = ((uint64_t)(lbo_a >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
      const uint64_t db = ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
      const uint32_t a_kstep = (2 * lbo_a) >> 4, b_kstep = (2 * lbo_b) >> 4;
      const uint32_t bnc = (uint32_t)((c.N + 31) & ~31);
      const uint32_t b_stage = (uint32_t)(c.nk16 * 2 * c.N * 16) >> 4;
      uint32_t ph = 0;   // bit b: parity to wait for on bar[b]
      long long t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const uint32_t d0 = tmem + (uint32_t)((it & 1) * (256));
        uint32_t accum = 0;
        if (c.order == 0) {
          for (int tap = 0; tap < c.ntaps; ++tap) {
            const uint64_t adesc = da + (abase >> 4) + (uint32_t)(tap * 3);
            const uint64_t bdesc = db + (bbase >> 4) + (uint32_t)tap * b_stage;
            if (c.msub == 4) issue_tap<4>(d0, bnc, adesc, bdesc, idesc, accum, c.nk16, a_kstep, b_kstep);
            else if (c.msub == 2) issue_tap<2>(d0, bnc, adesc, bdesc, idesc, accum, c.nk16, a_kstep, b_kstep);
            else issue_tap<1>(d0, bnc, adesc, bdesc, idesc, accum, c.nk16, a_kstep, b_kstep);
            accum = 1;
          }
        } else {
          // sub-tile outermost: all taps / K-steps of one accumulator back to back
          for (int j = 0; j < c.msub; ++j) {
            accum = 0;
            for (int tap = 0; tap < c.ntaps; ++tap) {
              const uint64_t adesc = da + (abase >> 4) + (uint32_t)(tap * 3) + (uint32_t)j * 128u;
              const uint64_t bdesc = db + (bbase >> 4) + (uint32_t)tap * b_stage;
              issue_tap<1>(d0 + j * bnc, bnc, adesc, bdesc, idesc, accum, c.nk16, a_kstep, b_kstep);
              accum = 1;
            }
          }
        }
        const int b = it & 1;
        umma_commit(smem_u32(&bar[b]));
        if (it > 0) {   // keep one tile in flight: wait for the previous tile (double-buffered accumulators)
          const int pb = b ^ 1;
          mbar_wait(smem_u32(&bar[pb]), (ph >> pb) & 1u);
          ph ^= 1u << pb;
        }
      }
      { const int pb = (c.iters - 1) & 1; mbar_wait(smem_u32(&bar[pb]), (ph >> pb) & 1u); }
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
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
  cudaFuncSetAttribute(bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("   M    N msub nk16 taps order   cyc/MMA   N/2   cyc/tile\n");
  const int cases[][4] = {{32, 4, 2, 3}, {32, 4, 2, 11}, {32, 2, 2, 11}, {32, 1, 2, 11}, {48, 4, 3, 7}, {48, 2, 3, 7},
                          {96, 2, 6, 3}, {96, 2, 6, 11}, {96, 1, 6, 11}, {192, 1, 4, 11}, {256, 1, 4, 11}, {64, 4, 4, 7}, {128, 2, 4, 7}};
  for (auto& cs : cases)
    for (int M : {128, 64})
    for (int order : {0, 1}) {
      Cfg c{cs[0], cs[1], cs[2], cs[3], order, 400, M};
      if (c.msub * ((c.N + 31) & ~31) > 256) continue;
      if (M == 64 && order == 1) continue;
      cudaMemset(out, 0, 148 * sizeof(long long));
      bench3<<<148, 128, 200 * 1024>>>(c, out);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double nm = (double)c.iters * c.ntaps * c.nk16 * c.msub;
      printf("%4d %4d %4d %4d %4d %5d  %8.1f  %5.0f  %9.0f\n", M, c.N, c.msub, c.nk16, c.ntaps, order, mx / nm, c.N / 2.0, mx / c.iters);
    }
  return 0;
}
