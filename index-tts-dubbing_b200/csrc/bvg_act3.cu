// Activation1d (alias_free_torch/act.py:24-29) with both FIR filters on the tensor cores -- the 16-bit performance
// modes (bf16 or fp16 storage), packed c8 layout.  Version 5 of the activation kernel.
//
// Why: the register-streamed kernel (bvg_act2.cu) is bound by the FP32 pipe, and 25 of its 31 fp32
// operations per element are FIR taps (tools/fma_bench.cu: 128 FMA lanes per SM per clock; packed f32x2
// does not raise that).  Both FIRs are banded Toeplitz products over time, so they map onto
// warp-level mma.sync.m16n8k16 (16-bit in, fp32 accumulate; tools/hmma_bench.cu: 0.49 MMAs per clock per
// SM) and leave the CUDA cores the snake (3 fp32 + 1 MUFU per activated sample):
//
//   U^T[row, m]  = sum_i X^T[row, i] * G^T[i, m]      up-FIR   (A = input rows, B = constant taps)
//   s'           = u - h cos(2 alpha u)               snake without its +h (added to y: sum of taps = 1)
//   Y^T[row, t]  = sum_m S'^T[row, m] * F^T[m, t]     down-FIR (A = activated samples, B = constant taps)
//
// The 16 MMA rows are two independent "streams" of 8 channels (an 8-channel chunk x a time tile each), the
// MMA columns are time.  The accumulator fragment of two consecutive up-FIR column tiles IS the A fragment of
// one down-FIR K-step (same trick as P = softmax(S) in attention kernels), so the 2x activated signal lives
// only in registers.  Operand movement: ldmatrix.trans turns the staged [time][8 channels] rows into A
// fragments, stmatrix.trans writes the result rows back.  The up-FIR runs in the storage type (its A operand
// is the stored tensor; taps rounded to it, optionally split hi + lo, BVG_ACT_MMA_UPLO=1); the down-FIR runs in
// fp16 (activated samples and taps rounded to 11 bits, saturating); its K-steps are offset by half a step against the output
// column tiles, so that 8 outputs need two K-steps of activated samples, not three.
//
// Formulas:  u[m] = 2 sum_k f[k] x[(m+5-k)/2],  y[t] = sum_k f[k] s[clamp(2t+k-5, 0, 2L-1)],
// with replicate padding of the input (staged rows are clamped to the segment) AND of the activated 2x signal.
// Version 5 handles the second padding inside the MMA pipeline: in the (at most two) tiles of a stream that
// touch a segment end, the activated samples outside [0, 2L) are replaced by s[0] / s[2L-1] (one warp shuffle
// each) before they enter the down-FIR -- the separate exact edge pass of version 4 (36 serial FIR evaluations
// per edge tile, ~16 us per launch on short segments) is gone.  Also new: the tile length is chosen per launch
// (balanced tiles, a whole number of waves when the problem is small), column tiles past a segment's end are
// not computed, and the storage type is a template parameter (bf16 / fp16).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include <cuda_fp16.h>

#include "bvg_common.cuh"

namespace {

constexpr int TWMAX = 240;         // most output rows per stream tile (3 blocks x 8 warps x 8.5 KB = 204 KB of shared memory per SM)
constexpr int XROWS = TWMAX + 32;  // staged rows per stream: local time -8 .. TW+23
constexpr int WPB = 8;             // warps per block
constexpr int MINB = 3;            // resident blocks per SM the kernel is compiled for (<= 85 registers)
// Timing experiments (tools/act_exp.cu compiles this file with -DBVG_ACT_EXP=<bits>; results are numerically WRONG):
// 1 = no cosine (MUFU), 2 = no down-FIR MMAs, 4 = no staging loads, 8 = no copy-out, 16 = no up-FIR MMA
#ifndef BVG_ACT_EXP
#define BVG_ACT_EXP 0
#endif

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// bulk (TMA-engine) copies: one instruction moves a whole contiguous slab of rows
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "ACT_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra ACT_WAIT_DONE;\n\t"
      "bra ACT_WAIT_LOOP;\n\t"
      "ACT_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stsm_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
// D = A B + C with separate C / D registers (a constant C needs no per-step moves)
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1,
                                         float c2, float c3) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
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
template <typename T> __device__ __forceinline__ uint32_t pack_io(float lo, float hi) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) return pack_bf16(lo, hi);
  else return pack_f16(lo, hi);
}
template <typename T> __device__ __forceinline__ float round_io(float v) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) return __bfloat162float(__float2bfloat16_rn(v));
  else return __half2float(__float2half_rn(v));
}
__device__ __forceinline__ float tap(int k) {   // f[k], 0 outside 0..11 (kaiser_sinc_filter1d(0.25, 0.3, 12), symmetric)
  const float f[6] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5};
  if (k < 0 || k > 11) return 0.f;
  return f[k < 6 ? k : 11 - k];
}

struct ActExtraJobs {
  const void* x[2];
  void* y[2];
  const float* alpha[2];
  const float* inv_beta[2];
};

// UP_LO: add the rounding residual of the up-FIR taps as a second MMA per column tile.
// PRE: the input is pre-scaled by 2 alpha per channel (u' = 2 alpha u comes straight out of the up-FIR), the output stays
// scaled:  v = u' - (2 alpha h) cos(u') = 2 alpha (s - h),  y' = down(v) + 2 alpha h = 2 alpha y  -- one multiply per sample less.
template <typename T, bool UP_LO, bool PRE = false>
__global__ void __launch_bounds__(WPB * 32, MINB)
act1d_c8_mma_kernel(const T* __restrict__ x0, T* __restrict__ y0, const float* __restrict__ alpha0,
                    const float* __restrict__ inv_beta0, const SegDesc* __restrict__ seg, int R, int ntiles, int nchunks,
                    int tw /* output rows per tile (multiple of 8, <= TWMAX) */,
                    int GT /* consecutive time tiles a warp processes per stream: amortises the constant set-up */,
                    const ActExtraJobs extra /* blockIdx.y = 1, 2: the same geometry on other tensors / parameters */) {
  extern __shared__ uint4 smem4[];
  const int job = blockIdx.y;
  const T* __restrict__ x = job == 0 ? x0 : reinterpret_cast<const T*>(job == 1 ? extra.x[0] : extra.x[1]);
  T* __restrict__ y = job == 0 ? y0 : reinterpret_cast<T*>(job == 1 ? extra.y[0] : extra.y[1]);
  const float* __restrict__ alpha = job == 0 ? alpha0 : (job == 1 ? extra.alpha[0] : extra.alpha[1]);
  const float* __restrict__ inv_beta = job == 0 ? inv_beta0 : (job == 1 ? extra.inv_beta[0] : extra.inv_beta[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  const SegDesc sd = seg[b];
  const int L = sd.len;
  // two streams per warp: consecutive (chunk, group of GT tiles) items of this segment (the second may not exist)
  const int ngroups = (ntiles + GT - 1) / GT;
  const int nitems = ngroups * nchunks;
  const int item0 = 2 * (blockIdx.x * WPB + warp);
  if (item0 >= nitems) return;   // warp-uniform; no block barrier is ever used
  int chunk[2], first_tile[2];
  bool have[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int it = item0 + s < nitems ? item0 + s : item0;
    chunk[s] = it / ngroups;
    first_tile[s] = (it - chunk[s] * ngroups) * GT * tw;
    have[s] = item0 + s < nitems;   // a missing second stream computes on a copy of the first and stores nothing
  }
  if (first_tile[0] >= L && (!have[1] || first_tile[1] >= L)) return;

  T* region = reinterpret_cast<T*>(smem4) + (size_t)warp * (2 * XROWS * 8);   // [stream][row][8 ch]
  // one mbarrier per warp (after the regions) for the bulk loads of interior tiles
  const uint32_t bar = smem_addr(reinterpret_cast<uint8_t*>(smem4) + (size_t)WPB * 2 * XROWS * 16) + 8u * warp;
  uint32_t bar_phase = 0;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // ---- constants: tap fragments and this thread's two channel rows (stream 0 / stream 1, channel g) ----
  // constant B fragment {B[2t][g], B[2t+1][g]} / {B[2t+8][g], B[2t+9][g]} of the up-FIR tap matrix B[k][n] = 2 f[n + 11 - 2k]
  uint32_t gup_hi[2], gup_lo[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int k0 = 2 * t + 8 * half;
    const float v0 = 2.f * tap(g + 11 - 2 * k0), v1 = 2.f * tap(g + 11 - 2 * (k0 + 1));
    const float h0 = round_io<T>(v0), h1 = round_io<T>(v1);
    gup_hi[half] = pack_io<T>(h0, h1);
    gup_lo[half] = pack_io<T>(v0 - h0, v1 - h1);
  }
  // down-FIR: K-step J' holds the activated samples 16J'-8 .. 16J'+7 (column tiles 2J'-1 and 2J'), so the 8 outputs of
  // column tile J (samples 16J-5 .. 16J+20) need exactly the two K-steps J and J+1:  B_d[k][n] = f[16 d + k - 2n - 3], d = 0, 1 (fp16)
  uint32_t fdn[2][2];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k0 = 2 * t + 8 * half, off = 16 * d - 3 - 2 * g;
      fdn[d][half] = pack_f16(tap(k0 + off), tap(k0 + 1 + off));
    }
  float a2[2], hh[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    a2[s] = 2.f * alpha[chunk[s] * 8 + g];
    hh[s] = 0.5f * inv_beta[chunk[s] * 8 + g];
    if (PRE) hh[s] *= a2[s];   // 2 alpha h
  }
  const uint32_t reg_s = smem_addr(region);
  // ldmatrix row address of this lane: matrices 0..3 = (stream 0, rows +0..7), (stream 1, +0..7), (stream 0, +8..15), (stream 1, +8..15)
  const uint32_t ld_base = reg_s + (uint32_t)(((lane >> 3) & 1) * XROWS + (lane >> 4) * 8 + (lane & 7)) * 16;
  // stmatrix row address: matrices 0 / 1 = stream 0 / 1, row = output row within the column tile
  const uint32_t st_base = reg_s + (uint32_t)(((lane >> 3) & 1) * XROWS + 8 + (lane & 7)) * 16;

  // everything above read launch constants only (segment table, snake parameters): from here on the previous kernel's
  // output is read and its input buffer overwritten
  pdl_trigger();
  pdl_wait();
#pragma unroll 1
  for (int gt = 0; gt < GT; ++gt) {
    int tile0[2], nrow[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      tile0[s] = first_tile[s] + gt * tw;
      nrow[s] = have[s] ? min(L - tile0[s], tw) : 0;   // valid output rows of this tile (<= 0: none)
    }
    const int nrmax = max(nrow[0], nrow[1]);
    if (nrmax <= 0) break;
    const int nJ = (nrmax + 7) >> 3;          // output column tiles to compute
    const int need = 8 * nJ + 24;             // staged rows the main pass reads (local time -8 .. 8 nJ + 15)
    // ---- stage the raw rows: region row r of stream s = x[clamp(tile0 - 8 + r, 0, L-1)] (replicate padding of the input).
    //      Interior tiles (no clamping, both streams present): ONE bulk copy per stream, issued by lane 0 ----
    const bool interior = have[1] && tile0[0] >= 8 && tile0[1] >= 8 && tile0[0] - 8 + need <= L && tile0[1] - 8 + need <= L;
    if (lane == 0) bulk_wait_read();   // the previous tile's bulk store has finished reading the region
    __syncwarp();
    if (interior && !(BVG_ACT_EXP & 4)) {
      if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(2 * need * 16));
#pragma unroll
        for (int s = 0; s < 2; ++s)
          bulk_g2s(reg_s + (uint32_t)(s * XROWS * 16), x + ((size_t)chunk[s] * R + sd.off + tile0[s] - 8) * 8, (uint32_t)(need * 16), bar);
      }
    } else {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const T* xs = x + ((size_t)chunk[s] * R + sd.off) * 8;
        T* rs = region + (size_t)s * XROWS * 8;
        const int r0 = tile0[s] - 8;
#pragma unroll
        for (int i = 0; i < (XROWS + 31) / 32; ++i) {
          const int r = lane + 32 * i;
          const int row = min(max(r0 + r, 0), L - 1);
          if (r < need && !(BVG_ACT_EXP & 4)) cp_async16(rs + r * 8, xs + (size_t)row * 8);
        }
      }
    }
    if (gt + 1 < GT) {   // pull the next tile's rows towards L2 while this one is processed
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int row = tile0[s] + tw - 8 + lane * 8;   // one 128-byte line per lane covers 8 rows
        if (row < L && lane * 8 < XROWS)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(x + ((size_t)chunk[s] * R + sd.off + row) * 8));
      }
    }
    if (interior && !(BVG_ACT_EXP & 4)) {
      mbar_wait(bar, bar_phase);
      bar_phase ^= 1u;
    } else {
      cp_async_wait_all();
      __syncwarp();
    }

    // ---- main pass ----
    // does a stream's tile touch a segment end?  start: samples < 0 exist only in column tile -1 of the tile at time 0;
    // end: a sample > hi = 2 (L - tile0) - 1 is read when the tile's last computed output row is within 3 rows of L
    const bool edge = tile0[0] == 0 || tile0[1] == 0 || L - tile0[0] < 8 * nJ + 3 || L - tile0[1] < 8 * nJ + 3;

    // one up-FIR column tile j: activated samples 8j .. 8j+7 (local) of both streams, v[0..1] stream 0, v[2..3] stream 1
    auto up_vals = [&](int j, float (&v)[4]) {
      uint32_t xa[4];
      ldsm_x4_trans(ld_base + (uint32_t)(4 * j + 5) * 16, xa);   // input rows 4j-3 .. 4j+12 (region row = local time + 8)
      float c[4];
      if (BVG_ACT_EXP & 16) { c[0] = __uint_as_float(xa[0]); c[1] = __uint_as_float(xa[1]); c[2] = __uint_as_float(xa[2]); c[3] = __uint_as_float(xa[3]); }
      else mma16816<T>(c, xa, gup_hi[0], gup_hi[1], 0.f, 0.f, 0.f, 0.f);
      if (UP_LO) mma16816<T>(c, xa, gup_lo[0], gup_lo[1], c[0], c[1], c[2], c[3]);
      if (BVG_ACT_EXP & 1) {
        v[0] = fmaf(-hh[0], a2[0] * c[0], c[0]); v[1] = fmaf(-hh[0], a2[0] * c[1], c[1]);
        v[2] = fmaf(-hh[1], a2[1] * c[2], c[2]); v[3] = fmaf(-hh[1], a2[1] * c[3], c[3]);
        return;
      }
      if (PRE) {
        v[0] = fmaf(-hh[0], __cosf(c[0]), c[0]); v[1] = fmaf(-hh[0], __cosf(c[1]), c[1]);
        v[2] = fmaf(-hh[1], __cosf(c[2]), c[2]); v[3] = fmaf(-hh[1], __cosf(c[3]), c[3]);
        return;
      }
      v[0] = fmaf(-hh[0], __cosf(a2[0] * c[0]), c[0]); v[1] = fmaf(-hh[0], __cosf(a2[0] * c[1]), c[1]);
      v[2] = fmaf(-hh[1], __cosf(a2[1] * c[2]), c[2]); v[3] = fmaf(-hh[1], __cosf(a2[1] * c[3]), c[3]);
    };
    // the accumulator's initial value (h per channel row) as four registers the compiler treats as distinct values, so
    // that it keeps the aligned quad alive instead of rebuilding it with moves in every step
    float hq[4];
    asm volatile("mov.f32 %0, %4; mov.f32 %1, %4; mov.f32 %2, %5; mov.f32 %3, %5;"
                 : "=f"(hq[0]), "=f"(hq[1]), "=f"(hq[2]), "=f"(hq[3]) : "f"(hh[0]), "f"(hh[1]));
    auto down = [&](int J, const uint32_t (&C)[4], const uint32_t (&N)[4]) {
      float c[4];
      if (BVG_ACT_EXP & 2) {
        c[0] = __uint_as_float(C[0] ^ N[0]); c[1] = __uint_as_float(C[1] ^ N[1]);
        c[2] = __uint_as_float(C[2] ^ N[2]); c[3] = __uint_as_float(C[3] ^ N[3]);
      } else {
        mma16816<__half>(c, C, fdn[0][0], fdn[0][1], hq[0], hq[1], hq[2], hq[3]);
        mma16816<__half>(c, N, fdn[1][0], fdn[1][1], c[0], c[1], c[2], c[3]);
      }
      // the raw rows these outputs overwrite (region rows 8J+8 .. 8J+15) were consumed by column tiles <= 2J+2; later
      // column tiles start at region row 8J+17
      stsm_x2_trans(st_base + (uint32_t)(8 * J) * 16, pack_io<T>(c[0], c[1]), pack_io<T>(c[2], c[3]));
    };
    uint32_t ka[4], kb[4];   // down-FIR A fragments of two consecutive K-steps: {tile 2J'-1: stream 0, stream 1, tile 2J': stream 0, stream 1}

    if (!edge) {
      auto up_tile = [&](int j, uint32_t& p0, uint32_t& p1) {
        float v[4];
        up_vals(j, v);
        p0 = pack_f16(v[0], v[1]);   // fp16 keeps 11 bits of the activated sample (bf16: 8); its range is ample for activations
        p1 = pack_f16(v[2], v[3]);
      };
      up_tile(-1, ka[0], ka[1]);
      up_tile(0, ka[2], ka[3]);
      // one output column tile J (rows 8J .. 8J+7 of both streams): C = K-step J, N = K-step J+1 (produced here)
      auto step = [&](int J, const uint32_t (&C)[4], uint32_t (&N)[4]) {
        up_tile(2 * J + 1, N[0], N[1]);
        up_tile(2 * J + 2, N[2], N[3]);
        down(J, C, N);
      };
      // the J loop is unrolled by 2 with alternating fragment names; an odd last column tile follows
      int J = 0;
#pragma unroll 1
      for (; J + 2 <= nJ; J += 2) {
        step(J, ka, kb);
        step(J + 1, kb, ka);
      }
      if (J < nJ) step(J, ka, kb);
    } else {
      // Tiles at a segment end: replicate padding of the ACTIVATED signal.  Column tiles are produced in increasing
      // order, so the last in-segment sample s[hi] (column tile hi >> 3) is known before any sample beyond it.
      int hi[2];
      float e[2] = {0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 2; ++s) hi[s] = 2 * (L - tile0[s]) - 1;
      auto up_edge = [&](int j, float (&v)[4]) {
        up_vals(j, v);
        const int m0 = 8 * j + 2 * t;   // local index of v[0] / v[2]; v[1] / v[3] are m0 + 1
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (j == (hi[s] >> 3))   // warp-uniform
            e[s] = __shfl_sync(0xffffffffu, (hi[s] & 1) ? v[2 * s + 1] : v[2 * s], (lane & ~3) | ((hi[s] & 7) >> 1));
          if (m0 > hi[s]) v[2 * s] = e[s];
          if (m0 + 1 > hi[s]) v[2 * s + 1] = e[s];
        }
      };
      float v0[4], vm[4];
      up_edge(0, v0);
      up_vals(-1, vm);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float first = __shfl_sync(0xffffffffu, v0[2 * s], lane & ~3);   // s[0] of this channel
        if (tile0[s] == 0) vm[2 * s] = vm[2 * s + 1] = first;
        else if (hi[s] < 0) vm[2 * s] = vm[2 * s + 1] = 0.f;   // stream past its segment: nothing is stored
      }
      ka[0] = pack_f16(vm[0], vm[1]); ka[1] = pack_f16(vm[2], vm[3]);
      ka[2] = pack_f16(v0[0], v0[1]); ka[3] = pack_f16(v0[2], v0[3]);
#pragma unroll 1
      for (int J = 0; J < nJ; ++J) {
        float va[4], vb[4];
        up_edge(2 * J + 1, va);
        up_edge(2 * J + 2, vb);
        kb[0] = pack_f16(va[0], va[1]); kb[1] = pack_f16(va[2], va[3]);
        kb[2] = pack_f16(vb[0], vb[1]); kb[3] = pack_f16(vb[2], vb[3]);
        down(J, ka, kb);
#pragma unroll
        for (int i = 0; i < 4; ++i) ka[i] = kb[i];
      }
    }
    // ---- copy the result rows out: one bulk store per stream (the rows are contiguous in the packed layout) ----
    fence_async_smem();   // the stmatrix results (generic proxy) become visible to the bulk-copy engine
    __syncwarp();
    if (lane == 0 && !(BVG_ACT_EXP & 8)) {
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (nrow[s] > 0)
          bulk_s2g(y + ((size_t)chunk[s] * R + sd.off + tile0[s]) * 8, reg_s + (uint32_t)((s * XROWS + 8) * 16), (uint32_t)(nrow[s] * 16));
      bulk_commit();
    }
    __syncwarp();   // the region is re-staged by the next tile (after bulk_wait_read)
  }
  if (lane == 0) bulk_wait_read();   // shared memory must stay valid until the last bulk store has read it
}

// Tile length / tiles per warp for a launch (measured: tools/gpu_act_tiling.sh, profiles/r2_act_tiling_sweep.txt).
//   * more than one wave of warps: balanced tiles of at most TWMAX rows (long tiles amortise the 32 halo rows and the
//     per-tile set-up; between 1 and 8 waves the tile length hardly matters, wave-quantisation models did not predict
//     the measurements), two consecutive tiles per warp from 4 waves on (with the three AMP blocks of a stage in one launch
//     stage 1 of config 2 -- 5.2 waves -- runs 0.60 -> 0.51 ms per step with two tiles per warp; profiles/r2g_act_tiling.txt);
//   * at most one wave (short segments, small batches): the kernel is bound by one warp's latency, so as many tiles
//     as still fit one wave (tiles of at least 32 rows).
struct ActTiling { int tw, ntiles, GT; };
ActTiling choose_tiling(int max_len, int nchunks, int B, long long slots /* resident warps on the device */) {
  const int nt_min = (max_len + TWMAX - 1) / TWMAX;
  auto tw_of = [&](int nt) { return ((max_len + nt - 1) / nt + 7) / 8 * 8; };
  auto warps_of = [&](int nt) { return (long long)((nt * nchunks + 1) / 2) * B; };
  int nt = nt_min;
  if (warps_of(nt_min) <= slots) {
    const int nt_max = max_len <= 32 ? 1 : (max_len + 31) / 32;
    while (nt < nt_max && warps_of(nt + 1) <= slots && tw_of(nt + 1) >= 32) ++nt;
  }
  const int tw = tw_of(nt);
  const int nt_eff = (max_len + tw - 1) / tw;
  return ActTiling{tw, nt_eff, warps_of(nt_eff) >= 4 * slots ? 2 : 1};
}

template <typename T>
cudaError_t launch_t(const ActArgs& a, cudaStream_t s) {
  static long long sms_of_dev[64] = {0};   // per device: resident warps of this kernel (and the function attributes are set)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const size_t smem = (size_t)WPB * 2 * XROWS * 16 + 8 * WPB;   // row regions + one mbarrier per warp
  if (!sms_of_dev[dev]) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(act1d_c8_mma_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = MINB;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, act1d_c8_mma_kernel<T, false>, WPB * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    sms_of_dev[dev] = (long long)(n > 0 ? n : 148) * per_sm * WPB;
  }
  const int nchunks = a.C / 8;
  static const int force_tw = [] { const char* e = getenv("BVG_ACT_TW"); return e ? atoi(e) : 0; }();
  const int njobs = 1 + (a.extra_jobs > 0 ? (a.extra_jobs > 2 ? 2 : a.extra_jobs) : 0);
  ActExtraJobs extra{};
  for (int j = 0; j + 1 < njobs; ++j) {
    extra.x[j] = a.xj[j]; extra.y[j] = a.yj[j]; extra.alpha[j] = a.alphaj[j]; extra.inv_beta[j] = a.inv_betaj[j];
  }
  ActTiling tl = choose_tiling(a.max_len, nchunks, a.B * njobs, sms_of_dev[dev]);
  static const int force_gt = [] { const char* e = getenv("BVG_ACT_GT"); return e ? atoi(e) : 0; }();
  if (force_tw >= 8 && force_tw <= TWMAX) {
    tl.tw = force_tw / 8 * 8; tl.ntiles = (a.max_len + tl.tw - 1) / tl.tw; tl.GT = tl.ntiles >= 2 ? 2 : 1;
  }
  if (force_gt == 1 || (force_gt == 2 && tl.ntiles >= 2)) tl.GT = force_gt;
  static const int dbg = [] { const char* e = getenv("BVG_ACT_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg) fprintf(stderr, "act: C %d max_len %d B %d -> tw %d ntiles %d GT %d (resident warps %lld)\n", a.C, a.max_len, a.B, tl.tw, tl.ntiles, tl.GT, sms_of_dev[dev]);
  const int nitems = ((tl.ntiles + tl.GT - 1) / tl.GT) * nchunks;
  dim3 grid((nitems + 2 * WPB - 1) / (2 * WPB), njobs, a.B), block(WPB * 32);
  // BVG_ACT_MMA_UPLO=1 adds the rounding residual of the up-FIR taps (a second MMA per column tile): +0.4 dB of SNR
  // on config 1 in the bf16 mode for ~4 % of the step time; off by default.
  static const int up_lo = [] { const char* e = getenv("BVG_ACT_MMA_UPLO"); return e ? atoi(e) : 0; }();
  if (a.prescaled)   // (the hi + lo tap split is a bf16-accuracy aid of the plain variant only)
    return launch_pdl(act1d_c8_mma_kernel<T, false, true>, grid, block, smem, s, (const T*)a.x, (T*)a.y, a.alpha, a.inv_beta, a.seg, a.R,
                      tl.ntiles, nchunks, tl.tw, tl.GT, extra);
  if (!up_lo)
    return launch_pdl(act1d_c8_mma_kernel<T, false>, grid, block, smem, s, (const T*)a.x, (T*)a.y, a.alpha, a.inv_beta, a.seg, a.R,
                      tl.ntiles, nchunks, tl.tw, tl.GT, extra);
  return launch_pdl(act1d_c8_mma_kernel<T, true>, grid, block, smem, s, (const T*)a.x, (T*)a.y, a.alpha, a.inv_beta, a.seg, a.R,
                    tl.ntiles, nchunks, tl.tw, tl.GT, extra);
}

}  // namespace

// 16-bit packed layout only (dtype 1 = bf16, 2 = fp16); same contract as launch_act_c8 (bvg_act.cu).
cudaError_t launch_act_c8_mma(const ActArgs& a, int dtype, cudaStream_t s) {
  if (a.B <= 0 || a.max_len <= 0) return cudaSuccess;
  if (dtype == 1) return launch_t<__nv_bfloat16>(a, s);
  if (dtype == 2) return launch_t<__half>(a, s);
  return cudaErrorInvalidValue;
}
