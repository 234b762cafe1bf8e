// Activation1d (alias_free_torch/act.py:24-29) with both FIR filters on the tensor cores -- bf16 performance
// mode, packed c8 layout.  Version 4 of the activation kernel.
//
// Why: the register-streamed kernel (bvg_act2.cu) is bound by the FP32 pipe, and 25 of its 31 fp32
// operations per element are FIR taps (tools/fma_bench.cu: 128 FMA lanes per SM per clock; packed f32x2
// does not raise that).  Both FIRs are banded Toeplitz products over time, so they map onto
// warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate; tools/hmma_bench.cu: 0.49 MMAs per clock per
// SM) and leave the CUDA cores the snake (3 fp32 + 1 MUFU per activated sample):
//
//   U^T[row, m]  = sum_i X^T[row, i] * G^T[i, m]      up-FIR   (A = input rows, B = constant taps)
//   s'           = u - h cos(2 alpha u)               snake without its +h (added to y: sum of taps = 1)
//   Y^T[row, t]  = sum_m S'^T[row, m] * F^T[m, t]     down-FIR (A = activated samples, B = constant taps)
//
// The 16 MMA rows are two independent "streams" of 8 channels (an 8-channel chunk x a 128-row time
// tile each), the MMA columns are time.  The accumulator fragment of two consecutive up-FIR column
// tiles IS the A fragment of one down-FIR K-step (same trick as P = softmax(S) in attention kernels),
// so the 2x activated signal lives only in registers.  Operand movement: ldmatrix.trans turns the
// staged [time][8 channels] rows into A fragments, stmatrix.trans writes the result rows back.
// The up-FIR runs in bf16 (its A operand is the stored bf16 tensor; taps rounded to bf16, optionally split hi + lo,
// BVG_ACT_MMA_UPLO=1); the down-FIR runs in fp16 (activated samples and taps rounded to 11 bits: finer than the 8 bits the
// stored result keeps anyway, and fp16's range is ample for activation magnitudes).
//
// Formulas:  u[m] = 2 sum_k f[k] x[(m+5-k)/2],  y[t] = sum_k f[k] s[2t+k-5], replicate
// padding of the input (staged rows are clamped) and of the activated signal (the three outputs next to
// each segment end are recomputed exactly, actcore::exact_clamped).
#include <cstdlib>

#include <cuda_fp16.h>

#include "bvg_act_core.cuh"
#include "bvg_common.cuh"

namespace {

constexpr int TW = 224;            // output rows per stream tile
constexpr int XROWS = TW + 32;     // staged rows per stream: local time -8 .. TW+23
constexpr int NJ = TW / 8;         // output column tiles per stream tile
constexpr int WPB = 8;             // warps per block
constexpr int MINB = 3;            // resident blocks per SM the kernel is compiled for (<= 64 registers)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stsm_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// saturating (F2FP.SATFINITE, same cost): an activated sample beyond fp16's range (tiny beta) becomes +-65504 instead of
// inf, which the symmetric down-FIR taps would turn into NaN across the output
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float tap(int k) {   // f[k], 0 outside 0..11 (kaiser_sinc_filter1d(0.25, 0.3, 12), symmetric)
  const float f[6] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5};
  if (k < 0 || k > 11) return 0.f;
  return f[k < 6 ? k : 11 - k];
}
// constant B fragment {B[2t][g], B[2t+1][g]} / {B[2t+8][g], B[2t+9][g]} of a tap matrix B[k][n] = scale * f[idx(k, n)],
// split into a bf16 "hi" part and the bf16 rounding residual "lo"
template <typename F>
__device__ __forceinline__ void tap_frag(F idx, float scale, int g, int t, uint32_t (&hi)[2], uint32_t (&lo)[2]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int k0 = 2 * t + 8 * half;
    const float v0 = scale * tap(idx(k0, g)), v1 = scale * tap(idx(k0 + 1, g));
    const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
    hi[half] = pack_bf16(h0, h1);
    lo[half] = pack_bf16(v0 - h0, v1 - h1);
  }
}

struct Stream {
  int chunk, tile0, L;        // 8-channel chunk, first output row, segment length (tile0 >= L: nothing to do)
};

template <bool UP_LO>   // UP_LO: add the bf16 rounding residual of the up-FIR taps (second MMA per column tile)
__global__ void __launch_bounds__(WPB * 32, MINB)
act1d_c8_mma_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ alpha,
                    const float* __restrict__ inv_beta, const SegDesc* __restrict__ seg, int R, int ntiles, int nchunks,
                    int GT /* consecutive time tiles a warp processes per stream: amortises the constant set-up */) {
  extern __shared__ uint4 smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  const SegDesc sd = seg[b];
  const int L = sd.len;
  // two streams per warp: consecutive (chunk, group of GT tiles) items of this segment (the second may not exist)
  const int ngroups = (ntiles + GT - 1) / GT;
  const int nitems = ngroups * nchunks;
  const int item0 = 2 * (blockIdx.x * WPB + warp);
  if (item0 >= nitems) return;   // warp-uniform; no block barrier is ever used
  Stream st[2];
  int first_tile[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int it = item0 + s < nitems ? item0 + s : item0;
    st[s].chunk = it / ngroups;
    first_tile[s] = (it - st[s].chunk * ngroups) * GT * TW;
    st[s].L = item0 + s < nitems ? L : 0;   // a missing second stream computes on a copy of the first and stores nothing
  }
  if (first_tile[0] >= L && (st[1].L == 0 || first_tile[1] >= L)) return;

  __nv_bfloat16* region = reinterpret_cast<__nv_bfloat16*>(smem4) + (size_t)warp * (2 * XROWS * 8);   // [stream][row][8 ch]
  // ---- constants: tap fragments and this thread's two channel rows (stream 0 / stream 1, channel g) ----
  uint32_t gup_hi[2], gup_lo[2];                      // up-FIR:   B[k][n] = 2 f[n + 11 - 2k]
  tap_frag([](int k, int n) { return n + 11 - 2 * k; }, 2.f, g, t, gup_hi, gup_lo);
  uint32_t fdn[3][2];                                 // down-FIR: B_d[k][n] = f[16 d + k - 2n + 5], d = -1, 0, +1, as fp16
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k0 = 2 * t + 8 * half, off = 16 * (d - 1) + 5 - 2 * g;
      fdn[d][half] = pack_f16(tap(k0 + off), tap(k0 + 1 + off));
    }
  float a2[2], hh[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    a2[s] = 2.f * alpha[st[s].chunk * 8 + g];
    hh[s] = 0.5f * inv_beta[st[s].chunk * 8 + g];
  }
  const uint32_t reg_s = smem_addr(region);
  // ldmatrix row address of this lane: matrices 0..3 = (stream 0, rows +0..7), (stream 1, +0..7), (stream 0, +8..15), (stream 1, +8..15)
  const uint32_t ld_base = reg_s + (uint32_t)(((lane >> 3) & 1) * XROWS + (lane >> 4) * 8 + (lane & 7)) * 16;
  // stmatrix row address: matrices 0 / 1 = stream 0 / 1, row = output row within the column tile
  const uint32_t st_base = reg_s + (uint32_t)(((lane >> 3) & 1) * XROWS + 8 + (lane & 7)) * 16;

#pragma unroll 1
  for (int gt = 0; gt < GT; ++gt) {
  st[0].tile0 = first_tile[0] + gt * TW;
  st[1].tile0 = first_tile[1] + gt * TW;
  if (st[0].tile0 >= L && (st[1].L == 0 || st[1].tile0 >= L)) break;
  // ---- stage the raw rows: region row r of stream s = x[clamp(tile0 - 8 + r, 0, L-1)] (replicate padding of the input) ----
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const __nv_bfloat16* xs = x + ((size_t)st[s].chunk * R + sd.off) * 8;
    __nv_bfloat16* rs = region + (size_t)s * XROWS * 8;
    const int r0 = st[s].tile0 - 8;
#pragma unroll
    for (int i = 0; i < XROWS / 32; ++i) {
      const int r = lane + 32 * i;
      const int row = min(max(r0 + r, 0), L - 1);
      cp_async16(rs + r * 8, xs + (size_t)row * 8);
    }
  }
  if (gt + 1 < GT) {   // pull the next tile's rows towards L2 while this one is processed
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = st[s].tile0 + TW - 8 + lane * 8;   // one 128-byte line per lane covers 8 rows
      if (row < L && lane * 8 < XROWS)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + ((size_t)st[s].chunk * R + sd.off + row) * 8));
    }
  }
  cp_async_wait_all();
  __syncwarp();

  // ---- exact values of the outputs that see the replicate padding of the activated signal (computed from
  //      the raw rows before the in-place output pass overwrites them): lane = (stream, end, channel pair) ----
  float2 fix[3];
  int fix_t[3] = {-1, -1, -1};
  if (lane < 16) {
    const int s = lane >> 3, end = (lane >> 2) & 1, cp = lane & 3;
    if (st[s].L > 0) {
      const int ch = st[s].chunk * 8 + 2 * cp;
      const float fa0 = 2.f * alpha[ch], fa1 = 2.f * alpha[ch + 1], fh0 = 0.5f * inv_beta[ch], fh1 = 0.5f * inv_beta[ch + 1];
      const __nv_bfloat16* col0 = region + (size_t)s * XROWS * 8 + 2 * cp;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ts = end ? L - 3 + j : j;
        if (ts >= 0 && ts < L && ts >= st[s].tile0 && ts < st[s].tile0 + TW && (end || ts < L - 3)) {
          fix_t[j] = ts;
          fix[j] = actcore::exact_clamped(col0, st[s].tile0 - 8, XROWS, ts, L, fa0, fa1, fh0, fh1);
        }
      }
    }
  }
  __syncwarp();

  // ---- main pass ----

  // one up-FIR column tile j (activated samples 8j .. 8j+7 of both streams) -> packed bf16 pair per stream
  auto up_tile = [&](int j, uint32_t& p0, uint32_t& p1) {
    uint32_t xa[4];
    ldsm_x4_trans(ld_base + (uint32_t)(4 * j + 5) * 16, xa);   // input rows 4j-3 .. 4j+12 (region row = local time + 8)
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(c, xa, gup_hi[0], gup_hi[1]);
    if (UP_LO) mma16816(c, xa, gup_lo[0], gup_lo[1]);
    const float s0 = fmaf(-hh[0], __cosf(a2[0] * c[0]), c[0]), s1 = fmaf(-hh[0], __cosf(a2[0] * c[1]), c[1]);
    const float s2 = fmaf(-hh[1], __cosf(a2[1] * c[2]), c[2]), s3 = fmaf(-hh[1], __cosf(a2[1] * c[3]), c[3]);
    p0 = pack_f16(s0, s1);   // fp16 keeps 11 bits of the activated sample (bf16: 8); its range is ample for activations
    p1 = pack_f16(s2, s3);
  };
  uint32_t ap[4], ac[4], an[4];   // down-FIR A fragments of K-steps J-1, J, J+1 (16 activated samples each)
  ap[0] = ap[1] = 0u;             // samples -16 .. -9 are never used (zero taps): skip their column tile
  up_tile(-1, ap[2], ap[3]);
  up_tile(0, ac[0], ac[1]);
  up_tile(1, ac[2], ac[3]);
  // one output column tile J (rows 8J .. 8J+7 of both streams); P / C / N = K-steps J-1, J, J+1 (N is produced here)
  auto step = [&](int J, const uint32_t (&P)[4], const uint32_t (&C)[4], uint32_t (&N)[4], bool last) {
    up_tile(2 * J + 2, N[0], N[1]);
    if (!last) up_tile(2 * J + 3, N[2], N[3]);
    else N[2] = N[3] = 0u;        // beyond the last sample any output of this tile needs
    float c[4] = {hh[0], hh[0], hh[1], hh[1]};
    mma16816_f16(c, P, fdn[0][0], fdn[0][1]);
    mma16816_f16(c, C, fdn[1][0], fdn[1][1]);
    mma16816_f16(c, N, fdn[2][0], fdn[2][1]);
    // the raw rows these outputs overwrite were consumed by the column tiles above
    stsm_x2_trans(st_base + (uint32_t)(8 * J) * 16, pack_bf16(c[0], c[1]), pack_bf16(c[2], c[3]));
  };
  // the J loop is unrolled by 3 with rotating fragment names; the last one to three column tiles follow separately
  constexpr int NJ3 = ((NJ - 1) / 3) * 3;
#pragma unroll 1
  for (int J = 0; J < NJ3; J += 3) {
    step(J, ap, ac, an, false);
    step(J + 1, ac, an, ap, false);
    step(J + 2, an, ap, ac, false);
  }
  if constexpr (NJ - NJ3 == 1) {
    step(NJ - 1, ap, ac, an, true);
  } else if constexpr (NJ - NJ3 == 2) {
    step(NJ - 2, ap, ac, an, false);
    step(NJ - 1, ac, an, ap, true);
  } else {
    step(NJ - 3, ap, ac, an, false);
    step(NJ - 2, ac, an, ap, false);
    step(NJ - 1, an, ap, ac, true);
  }
  __syncwarp();
  // ---- patch the exact edge values in ----
  if (lane < 16) {
    const int s = lane >> 3, cp = lane & 3;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (fix_t[j] >= 0) actcore::stpair(region + ((size_t)s * XROWS + 8 + fix_t[j] - st[s].tile0) * 8 + 2 * cp, fix[j]);
  }
  __syncwarp();
  // ---- copy the result rows out (16 bytes per lane, consecutive rows) ----
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    __nv_bfloat16* ys = y + ((size_t)st[s].chunk * R + sd.off + st[s].tile0) * 8;
    const __nv_bfloat16* rs = region + ((size_t)s * XROWS + 8) * 8;
    const int nrow = st[s].L - st[s].tile0;   // valid rows of this tile (<= 0: none)
#pragma unroll
    for (int i = 0; i < TW / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < nrow) *reinterpret_cast<uint4*>(ys + (size_t)r * 8) = *reinterpret_cast<const uint4*>(rs + r * 8);
    }
  }
  __syncwarp();   // the region is re-staged by the next tile
  }
}

}  // namespace

// bf16 packed layout only; same contract as launch_act_c8 (bvg_act.cu).
cudaError_t launch_act_c8_mma(const ActArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.max_len <= 0) return cudaSuccess;
  const int ntiles = (a.max_len + TW - 1) / TW, nchunks = a.C / 8;
  const int GT = ntiles >= 2 ? 2 : 1;
  const int nitems = ((ntiles + GT - 1) / GT) * nchunks;
  dim3 grid((nitems + 2 * WPB - 1) / (2 * WPB), 1, a.B), block(WPB * 32);
  const size_t smem = (size_t)WPB * 2 * XROWS * 16;
  static bool attr_done_dev[64] = {false};   // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  bool& attr_done = attr_done_dev[dev];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(act1d_c8_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // keep the SMs at the maximum shared-memory carve-out, the one the persistent conv kernel needs: alternating
    // act / conv launches then never wait for an L1 / shared-memory reconfiguration (BVG_CARVEOUT=0 to compare)
    static const int carve = [] { const char* c = getenv("BVG_CARVEOUT"); return c ? atoi(c) : 1; }();
    if (carve && e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (carve && e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  // BVG_ACT_MMA_UPLO=1 adds the bf16 rounding residual of the up-FIR taps (a second MMA per column tile): +0.4 dB of SNR
  // (31.9 instead of 31.5 dB on config 1) for ~4 % of the step time; off by default.
  static const int up_lo = [] { const char* e = getenv("BVG_ACT_MMA_UPLO"); return e ? atoi(e) : 0; }();
  if (!up_lo)
    act1d_c8_mma_kernel<false><<<grid, block, smem, s>>>((const __nv_bfloat16*)a.x, (__nv_bfloat16*)a.y, a.alpha, a.inv_beta, a.seg, a.R,
                                                        ntiles, nchunks, GT);
  else
  act1d_c8_mma_kernel<true><<<grid, block, smem, s>>>((const __nv_bfloat16*)a.x, (__nv_bfloat16*)a.y, a.alpha, a.inv_beta, a.seg, a.R, ntiles,
                                               nchunks, GT);
  return cudaGetLastError();
}
