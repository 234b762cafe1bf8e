// Micro-benchmark: per-SM issue throughput of the instructions the Activation1d kernel is made of
// (FFMA reg/imm, FFMA2 = fma.rn.f32x2, HFMA2, MUFU.COS incl. its range-reduction FMUL), at 8/16/32 warps
// per SM.  Prints warp-instructions per cycle per SM and the equivalent scalar FMA lanes per clock.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 tools/fma_bench.cu -o /tmp/fb && /tmp/fb
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 8;   // 8 independent chains per thread

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[CH], b = seed, c = seed * 0.5f;
  unsigned long long p[CH];
  __half2 h[CH], hb = __floats2half2_rn(seed, seed), hc = __floats2half2_rn(0.5f, 0.25f);
  for (int i = 0; i < CH; ++i) { a[i] = seed + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i]); h[i] = __floats2half2_rn(a[i], a[i]); }
  const unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b), pc = pb ^ 0x1000000010ull;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (OP == 0) a[i] = fmaf(a[i], b, c);                       // FFMA reg
      if (OP == 1) a[i] = fmaf(a[i], 0.99f, c);                   // FFMA imm
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
      if (OP == 3) h[i] = __hfma2(h[i], hb, hc);                  // HFMA2
      if (OP == 4) a[i] = __cosf(a[i]);                           // FMUL + MUFU.COS
      if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
      if (OP == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __low2float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int elems) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<148, warps * 32>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const double wi = (double)warps * ITERS * CH;     // warp-instructions per SM
    printf("%-10s warps/SM %2d  warp-instr/clk/SM %.2f   scalar-op lanes/clk/SM %.0f\n", name, warps, wi / mx, wi / mx * 32 * elems);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA reg", 1); run<1>("FFMA imm", 1); run<2>("FFMA2", 2); run<5>("FADD2", 2); run<6>("FMUL2", 2); run<3>("HFMA2", 2); run<4>("cos.approx", 1);
  return 0;
}
