// Timing experiments on the tensor-core Activation1d kernel (csrc/bvg_act3.cu compiled with -DBVG_ACT_EXP=<bits>):
// which resource bounds it?  Build + run: tools/gpu_act_exp.sh.  Results with EXP != 0 are numerically wrong on purpose.
#include <cstdio>
#include <vector>
#include "../index-tts-dubbing_b200/csrc/bvg_act3.cu"

bool bvg_pdl_enabled() { return false; }

int main(int argc, char** argv) {
  const int B = 16;
  const int C = argc > 1 ? atoi(argv[1]) : 24, T = argc > 2 ? atoi(argv[2]) : 240640;
  std::vector<SegDesc> seg(B);
  int R = BVG_GUARD;
  for (int b = 0; b < B; ++b) { seg[b] = SegDesc{R, T}; R += T + BVG_GUARD; }
  R += BVG_TAIL_SLACK;
  SegDesc* sd; __nv_bfloat16 *x, *y; float* prm; char* flush;
  const size_t n = (size_t)C * R * 8 / 8;
  cudaMalloc(&sd, B * sizeof(SegDesc)); cudaMemcpy(sd, seg.data(), B * sizeof(SegDesc), cudaMemcpyHostToDevice);
  cudaMalloc(&x, n * 2); cudaMalloc(&y, n * 2); cudaMalloc(&prm, 2 * C * 4); cudaMalloc(&flush, 256 << 20);
  cudaMemset(x, 0x3c, n * 2);   // bf16 0x3c3c ~ 0.0115
  std::vector<float> p(2 * C, 1.0f);
  cudaMemcpy(prm, p.data(), 2 * C * 4, cudaMemcpyHostToDevice);
  ActArgs a{x, y, prm, prm + C, sd, R, C, B, T};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, sum = 0.f; const int iters = 8;
  for (int i = 0; i < iters + 2; ++i) {
    cudaMemsetAsync(flush, i, 256 << 20);
    cudaEventRecord(e0);
    cudaError_t e = launch_act_c8_mma(a, 1, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 2) { best = ms < best ? ms : best; sum += ms; }
  }
  cudaError_t e = cudaDeviceSynchronize();
  const double bytes = 2.0 * C * (double)T * B * 2;
  printf("EXP %2d  C %3d T %6d: mean %.1f us  best %.1f us  %.0f GB/s  (%s)\n", BVG_ACT_EXP, C, T, sum / iters * 1e3, best * 1e3,
         bytes / (best * 1e-3) / 1e9, cudaGetErrorString(e));
  return 0;
}
