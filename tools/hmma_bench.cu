// Micro-benchmark: warp-level mma.sync.m16n8k16 (bf16 -> fp32) issue rate on sm_100a, per SM, at 4..32
// warps per SM.  Question: is the legacy warp-MMA path fast enough to take the FIR taps of Activation1d
// off the FP32 pipe (10 MMAs per 128 output elements)?
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 tools/hmma_bench.cu -o /tmp/hb && /tmp/hb
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 4;   // 4 independent accumulator chains per warp

__global__ void k(float* out, long long* cyc, uint32_t seed) {
  float c[CH][4];
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = seed, a1 = seed ^ 0x3f803f80u, a2 = seed + 7, a3 = seed ^ 0x3c003c00u, b0 = 0x3f803f80u, b1 = 0x3c003c00u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    k<<<148, warps * 32>>>(out, cyc, 0x3f003f00u);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const double n = (double)warps * ITERS * CH;
    printf("mma.sync m16n8k16 bf16  warps/SM %2d  MMAs/clk/SM %.3f  (cycles per MMA per SM %.2f, %.0f MAC/clk/SM)\n", warps, n / mx, mx / n,
           n / mx * 2048);
  }
  return 0;
}
