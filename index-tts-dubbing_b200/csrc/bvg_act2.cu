// Fused Activation1d, version 3: warp-autonomous, register-streamed (packed c8 layout only).
//
// Same function as act1d_kernel in bvg_act.cu (see the formulas and reference citations there);
// different mapping, chosen after the round-1 ncu captures (v1: issue-bound at ~95 thread-
// instructions and ~50 B of shared-memory traffic per element; v2: staging latency + slow edge path):
//
//   * one WARP = one 8-channel group x (8*RT) consecutive time steps; no block-level barrier at all.
//   * lane = (time block tb = lane>>2, channel pair cp = lane&3): a thread owns RT consecutive
//     outputs of TWO channels and streams along time with everything in registers:
//       x window (6 rows) -> two new activated samples per step -> s window (12) -> one output.
//     The up-FIR halo (3 pairs on each side) is recomputed per thread: (RT+5)/RT extra snake work.
//   * shared memory only stages the raw rows (verbatim 16-byte c8 rows fetched with cp.async, one
//     region per time block with an odd row stride so the 32 lanes hit 32 distinct banks) and carries
//     the outputs back for coalesced 16-byte stores.  A thread only ever touches ITS OWN 4-byte
//     column of a row, so outputs are written in place without any cross-lane hazard.
//   * replicate padding of the activated 2x signal only matters for the first 3 and last 3 outputs of
//     a segment: those (at most 6 per thread, only in the threads that hold a segment end) are
//     recomputed exactly by a generic routine BEFORE the streaming pass and patched in afterwards.
#include <cstdlib>

#include "bvg_common.cuh"

namespace {

__constant__ float c_taps[12] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5,
                                 BVG_F5, BVG_F4, BVG_F3, BVG_F2, BVG_F1, BVG_F0};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T> struct RowOps;
template <> struct RowOps<__nv_bfloat16> {
  static __device__ __forceinline__ void stage_row(__nv_bfloat16* d, const __nv_bfloat16* s) { cp_async16(d, s); }
  static __device__ __forceinline__ void copy_row(__nv_bfloat16* d, const __nv_bfloat16* s) {
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
  }
  static __device__ __forceinline__ float2 ldpair(const __nv_bfloat16* p) {
    uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
  static __device__ __forceinline__ void stpair(__nv_bfloat16* p, float2 v) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct RowOps<float> {
  static __device__ __forceinline__ void stage_row(float* d, const float* s) {
    cp_async16(d, s);
    cp_async16(d + 4, s + 4);
  }
  static __device__ __forceinline__ void copy_row(float* d, const float* s) {
    reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(s)[0];
    reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(s)[1];
  }
  static __device__ __forceinline__ float2 ldpair(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void stpair(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};

// s = u + sin^2(alpha u) / (beta + 1e-9).  Fast form: sin^2 z = (1 - cos 2z)/2, so
// s = u + h - h cos(2 alpha u) with h = inv_beta/2; the constant h is added once per OUTPUT instead
// (the down-FIR taps sum to 1 and replicate padding passes constants through).
template <int PRECISE>
__device__ __forceinline__ float snake1(float u, float a, float ib_or_h) {
  if (PRECISE == 1) {
    float sn = sinf(a * u);
    return fmaf(ib_or_h * sn, sn, u);
  } else if (PRECISE == 2) {
    // fast form with an fp32-grade cosine: z = k pi + r, |r| <= pi/2 (two-constant Cody-Waite, exact for |z| < 2^12 pi),
    // cos z = (-1)^k cos r, cos r by its degree-12 even Taylor polynomial (truncation error 7e-9 at r = pi/2)
    const float z = a * u;
    const float k = rintf(z * 0.318309886183790672f);
    float r = fmaf(k, -3.140625f, z);
    r = fmaf(k, -9.67653589793e-4f, r);
    const float r2 = r * r;
    float c = fmaf(r2, 2.08767569878681e-9f, -2.75573192239859e-7f);
    c = fmaf(c, r2, 2.48015873015873e-5f);
    c = fmaf(c, r2, -1.38888888888889e-3f);
    c = fmaf(c, r2, 4.16666666666667e-2f);
    c = fmaf(c, r2, -0.5f);
    c = fmaf(c, r2, 1.f);
    c = __int_as_float(__float_as_int(c) ^ (__float2int_rn(k) << 31));
    return fmaf(-ib_or_h, c, u);             // a = 2 alpha, ib_or_h = h
  } else if (PRECISE == 3) {
    // fast form, MUFU cosine of the argument reduced to [-pi, pi] (two-constant Cody-Waite): the hardware's 2^-21.2
    // absolute error bound holds whatever the size of alpha * u
    const float z = a * u;
    const float k = rintf(z * 0.159154943091895336f);
    float r = fmaf(k, -6.28125f, z);
    r = fmaf(k, -1.935307179586e-3f, r);
    return fmaf(-ib_or_h, __cosf(r), u);
  } else {
    return fmaf(-ib_or_h, __cosf(a * u), u);   // a = 2 alpha, ib_or_h = h
  }
}

constexpr float G0 = 2.f * BVG_F0, G1 = 2.f * BVG_F1, G2 = 2.f * BVG_F2, G3 = 2.f * BVG_F3, G4 = 2.f * BVG_F4,
                G5 = 2.f * BVG_F5;

// Exact output t of this thread (rows relative to r0) with both replicate paddings, reading the raw
// rows from the thread's staging region (row index of x[j] is j - r0 + 5; staged rows are clamped).
template <typename T, int NX, int PRECISE>
__device__ __noinline__ float2 exact_output(const T* xr, int r0, int t, int L, float a0, float a1, float h0, float h1) {
  float accx = 0.f, accy = 0.f;
#pragma unroll 1
  for (int k = 0; k < 12; ++k) {
    const int m = min(max(2 * (r0 + t) - 5 + k, 0), 2 * L - 1);   // replicate pad of the activated signal
    const int i = m >> 1;
    const int base = i - r0 + 5;
    float2 p[7];
#pragma unroll
    for (int d = -3; d <= 3; ++d) p[d + 3] = RowOps<T>::ldpair(xr + min(max(base + d, 0), NX - 1) * 8);
    float ux, uy;
    if (m & 1) {   // u[2i+1]
      ux = G0 * p[6].x + G2 * p[5].x + G4 * p[4].x + G5 * p[3].x + G3 * p[2].x + G1 * p[1].x;
      uy = G0 * p[6].y + G2 * p[5].y + G4 * p[4].y + G5 * p[3].y + G3 * p[2].y + G1 * p[1].y;
    } else {       // u[2i]
      ux = G1 * p[5].x + G3 * p[4].x + G5 * p[3].x + G4 * p[2].x + G2 * p[1].x + G0 * p[0].x;
      uy = G1 * p[5].y + G3 * p[4].y + G5 * p[3].y + G4 * p[2].y + G2 * p[1].y + G0 * p[0].y;
    }
    accx = fmaf(c_taps[k], snake1<PRECISE>(ux, a0, h0), accx);
    accy = fmaf(c_taps[k], snake1<PRECISE>(uy, a1, h1), accy);
  }
  if (PRECISE != 1) { accx += h0; accy += h1; }
  return make_float2(accx, accy);
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float2 upk2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// SPLIT (fp32 only): the result is written as a split fp16 tensor [2 C/8][R][8] = [hi chunks | lo chunks] (split_f32 in
// bvg_common.cuh), the A operand of the fp32 tensor-core convolutions (bvg_conv_umma.cu, F32IO), instead of fp32 rows.
template <typename T, int RT, int WPB, int PRECISE, bool PACKED, int MINB = (sizeof(T) == 2 ? 2 : 1), bool SPLIT = false>
__global__ void __launch_bounds__(WPB * 32, MINB)
act1d_c8_v3_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ alpha,
                   const float* __restrict__ inv_beta, const SegDesc* __restrict__ seg, int R, int tiles, int nchunks) {
  constexpr int RS = RT + 11;   // region stride in rows: odd, >= RT + 10
  constexpr int NX = RT + 10;   // staged rows per region: r0-5 .. r0+RT+4
  extern __shared__ uint4 smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item = (8-channel chunk, tile of 8*RT rows) of one segment, flattened so that the warps of a
  // block stay busy when a segment is only a few tiles long (the first stages: 940 rows = 4 tiles)
  // (tiles == 0 selects the plain mapping blockIdx = (tile group, chunk, segment) used for long segments)
  const int item = blockIdx.x * WPB + warp, b = blockIdx.z;
  const int chunk = tiles ? item / tiles : (int)blockIdx.y, tile = tiles ? item - chunk * tiles : item;
  if (chunk >= nchunks) return;   // warp-uniform; no block barrier is ever used
  const SegDesc sd = seg[b];
  const int L = sd.len;
  const int tile0 = tile * (8 * RT);
  if (tile0 >= L) return;
  T* region = reinterpret_cast<T*>(smem4) + (size_t)warp * (8 * RS * 8);
  const T* xb = x + ((size_t)chunk * R + sd.off) * 8;

  // ---- stage the raw rows (clamped = replicate padding of the input), all copies in flight at once --
  constexpr int NSTAGE = (8 * NX + 31) / 32;
#pragma unroll
  for (int it = 0; it < NSTAGE; ++it) {
    const int idx = lane + 32 * it;
    if (idx < 8 * NX) {
      const int tb = idx / NX, n = idx - tb * NX;
      const int row = min(max(tile0 + tb * RT - 5 + n, 0), L - 1);
      RowOps<T>::stage_row(region + (tb * RS + n) * 8, xb + (size_t)row * 8);
    }
  }
  const int cp = lane & 3, tb = lane >> 2;
  const int r0 = tile0 + tb * RT;
  T* xr = region + (tb * RS) * 8 + 2 * cp;   // row n, this thread's channel pair: xr[n*8], xr[n*8+1]
  const int ch = chunk * 8 + 2 * cp;
  float a0 = alpha[ch], a1 = alpha[ch + 1], h0 = inv_beta[ch], h1 = inv_beta[ch + 1];
  if (PRECISE != 1) { a0 *= 2.f; a1 *= 2.f; h0 *= 0.5f; h1 *= 0.5f; }
  cp_async_wait_all();
  __syncwarp();

  if (r0 < L) {
    // ---- exact values for the outputs that see the replicate padding of the activated signal ------
    const int nvalid = min(RT, L - r0);
    const int te = max(L - 3 - r0, 0);        // first output of this thread within 3 rows of the end
    float2 fs[3], fe[3];
    const bool fix_start = r0 == 0, fix_end = te < nvalid;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (fix_start && j < nvalid) fs[j] = exact_output<T, NX, PRECISE>(xr, r0, j, L, a0, a1, h0, h1);
      if (fix_end && te + j < nvalid) fe[j] = exact_output<T, NX, PRECISE>(xr, r0, te + j, L, a0, a1, h0, h1);
    }

    // ---- streaming pass (interior formulas; rows beyond the segment are clamped garbage) ----------
    if constexpr (PACKED) {
      // Blackwell packed-fp32 math (fma/mul/add.rn.f32x2): the thread's two channels ride in one
      // 64-bit register pair, halving the FIR / snake instruction count.
      u64 xw[NX];
      u64 s[2 * RT + 10];
      const u64 GG0 = pk2(G0, G0), GG1 = pk2(G1, G1), GG2 = pk2(G2, G2), GG3 = pk2(G3, G3), GG4 = pk2(G4, G4),
                GG5 = pk2(G5, G5);
      const u64 FF0 = pk2(BVG_F0, BVG_F0), FF1 = pk2(BVG_F1, BVG_F1), FF2 = pk2(BVG_F2, BVG_F2),
                FF3 = pk2(BVG_F3, BVG_F3), FF4 = pk2(BVG_F4, BVG_F4), FF5 = pk2(BVG_F5, BVG_F5);
      const u64 AA = pk2(a0, a1), NH = pk2(-h0, -h1), HH = pk2(h0, h1);
      auto snake2 = [&](u64 u) {
        const float2 tt = upk2(mul2(AA, u));
        return fma2(NH, pk2(__cosf(tt.x), __cosf(tt.y)), u);
      };
      auto up_odd = [&](int j) {
        u64 u = mul2(GG1, xw[j]);
        u = fma2(GG3, xw[j + 1], u); u = fma2(GG5, xw[j + 2], u); u = fma2(GG4, xw[j + 3], u);
        u = fma2(GG2, xw[j + 4], u); u = fma2(GG0, xw[j + 5], u);
        return snake2(u);
      };
      auto up_even = [&](int j) {
        u64 u = mul2(GG0, xw[j - 1]);
        u = fma2(GG2, xw[j], u); u = fma2(GG4, xw[j + 1], u); u = fma2(GG5, xw[j + 2], u);
        u = fma2(GG3, xw[j + 3], u); u = fma2(GG1, xw[j + 4], u);
        return snake2(u);
      };
#pragma unroll
      for (int n = 0; n < 10; ++n) { const float2 v = RowOps<T>::ldpair(xr + n * 8); xw[n] = pk2(v.x, v.y); }
#pragma unroll
      for (int n = 0; n < 10; ++n) s[n] = (n & 1) ? up_even((n + 1) / 2) : up_odd(n / 2);
#pragma unroll
      for (int t = 0; t < RT; ++t) {
        { const float2 v = RowOps<T>::ldpair(xr + (t + 10) * 8); xw[t + 10] = pk2(v.x, v.y); }
        s[2 * t + 10] = up_odd(t + 5);
        s[2 * t + 11] = up_even(t + 6);
        u64 acc = mul2(FF0, add2(s[2 * t], s[2 * t + 11]));
        acc = fma2(FF1, add2(s[2 * t + 1], s[2 * t + 10]), acc);
        acc = fma2(FF2, add2(s[2 * t + 2], s[2 * t + 9]), acc);
        acc = fma2(FF3, add2(s[2 * t + 3], s[2 * t + 8]), acc);
        acc = fma2(FF4, add2(s[2 * t + 4], s[2 * t + 7]), acc);
        acc = fma2(FF5, add2(s[2 * t + 5], s[2 * t + 6]), acc);
        acc = add2(acc, HH);
        RowOps<T>::stpair(xr + t * 8, upk2(acc));
      }
    } else {
    float2 xw[NX];
    float2 s[2 * RT + 10];
#pragma unroll
    for (int n = 0; n < 10; ++n) xw[n] = RowOps<T>::ldpair(xr + n * 8);
    // s[n] is the activated sample m = 2 r0 - 5 + n
    auto up_odd = [&](int j) {   // n even: sample 2i+1 of pair i = r0-3+j, taps on xw[j..j+5]
      float2 u;
      u.x = G1 * xw[j].x; u.y = G1 * xw[j].y;
      u.x = fmaf(G3, xw[j + 1].x, u.x); u.y = fmaf(G3, xw[j + 1].y, u.y);
      u.x = fmaf(G5, xw[j + 2].x, u.x); u.y = fmaf(G5, xw[j + 2].y, u.y);
      u.x = fmaf(G4, xw[j + 3].x, u.x); u.y = fmaf(G4, xw[j + 3].y, u.y);
      u.x = fmaf(G2, xw[j + 4].x, u.x); u.y = fmaf(G2, xw[j + 4].y, u.y);
      u.x = fmaf(G0, xw[j + 5].x, u.x); u.y = fmaf(G0, xw[j + 5].y, u.y);
      return make_float2(snake1<PRECISE>(u.x, a0, h0), snake1<PRECISE>(u.y, a1, h1));
    };
    auto up_even = [&](int j) {  // n odd: sample 2i of pair i = r0-3+j, taps on xw[j-1..j+4]
      float2 u;
      u.x = G0 * xw[j - 1].x; u.y = G0 * xw[j - 1].y;
      u.x = fmaf(G2, xw[j].x, u.x); u.y = fmaf(G2, xw[j].y, u.y);
      u.x = fmaf(G4, xw[j + 1].x, u.x); u.y = fmaf(G4, xw[j + 1].y, u.y);
      u.x = fmaf(G5, xw[j + 2].x, u.x); u.y = fmaf(G5, xw[j + 2].y, u.y);
      u.x = fmaf(G3, xw[j + 3].x, u.x); u.y = fmaf(G3, xw[j + 3].y, u.y);
      u.x = fmaf(G1, xw[j + 4].x, u.x); u.y = fmaf(G1, xw[j + 4].y, u.y);
      return make_float2(snake1<PRECISE>(u.x, a0, h0), snake1<PRECISE>(u.y, a1, h1));
    };
#pragma unroll
    for (int n = 0; n < 10; ++n) s[n] = (n & 1) ? up_even((n + 1) / 2) : up_odd(n / 2);
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      xw[t + 10] = RowOps<T>::ldpair(xr + (t + 10) * 8);
      s[2 * t + 10] = up_odd(t + 5);
      s[2 * t + 11] = up_even(t + 6);
      float2 acc;
      acc.x = BVG_F0 * (s[2 * t].x + s[2 * t + 11].x); acc.y = BVG_F0 * (s[2 * t].y + s[2 * t + 11].y);
      acc.x = fmaf(BVG_F1, s[2 * t + 1].x + s[2 * t + 10].x, acc.x); acc.y = fmaf(BVG_F1, s[2 * t + 1].y + s[2 * t + 10].y, acc.y);
      acc.x = fmaf(BVG_F2, s[2 * t + 2].x + s[2 * t + 9].x, acc.x); acc.y = fmaf(BVG_F2, s[2 * t + 2].y + s[2 * t + 9].y, acc.y);
      acc.x = fmaf(BVG_F3, s[2 * t + 3].x + s[2 * t + 8].x, acc.x); acc.y = fmaf(BVG_F3, s[2 * t + 3].y + s[2 * t + 8].y, acc.y);
      acc.x = fmaf(BVG_F4, s[2 * t + 4].x + s[2 * t + 7].x, acc.x); acc.y = fmaf(BVG_F4, s[2 * t + 4].y + s[2 * t + 7].y, acc.y);
      acc.x = fmaf(BVG_F5, s[2 * t + 5].x + s[2 * t + 6].x, acc.x); acc.y = fmaf(BVG_F5, s[2 * t + 5].y + s[2 * t + 6].y, acc.y);
      if (PRECISE != 1) { acc.x += h0; acc.y += h1; }
      RowOps<T>::stpair(xr + t * 8, acc);   // row t is dead: in-place, own column only
    }
    }
    // ---- patch the exact edge values in ---------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (fix_start && j < nvalid) RowOps<T>::stpair(xr + j * 8, fs[j]);
      if (fix_end && te + j < nvalid) RowOps<T>::stpair(xr + (te + j) * 8, fe[j]);
    }
  }
  __syncwarp();

  // ---- coalesced write-back of the 8*RT output rows ----------------------------------------------
  T* yb = y + ((size_t)chunk * R + sd.off) * 8;
#pragma unroll
  for (int it = 0; it < (8 * RT) / 32; ++it) {
    const int idx = lane + 32 * it;
    const int tb2 = idx / RT, t = idx - tb2 * RT;
    const int row = tile0 + tb2 * RT + t;
    if (row < L) {
      if constexpr (SPLIT) {
        const float* src = reinterpret_cast<const float*>(region + (tb2 * RS + t) * 8);
        Vec8<__half> hi, lo;
#pragma unroll
        for (int c = 0; c < 8; ++c) split_f32(src[c], hi.v[c], lo.v[c]);
        __half* ys = reinterpret_cast<__half*>(y);
        hi.store(ys + ((size_t)chunk * R + sd.off + row) * 8);
        lo.store(ys + ((size_t)(nchunks + chunk) * R + sd.off + row) * 8);
      } else {
        RowOps<T>::copy_row(yb + (size_t)row * 8, region + (tb2 * RS + t) * 8);
      }
    }
  }
}

template <typename T, int RT, int WPB, int PRECISE, bool PACKED = false, int MINB = (sizeof(T) == 2 ? 2 : 1), bool SPLIT = false>
cudaError_t launch_v3(const ActArgs& a, cudaStream_t s) {
  const int tiles = (a.max_len + 8 * RT - 1) / (8 * RT);
  const int nchunks = a.C / 8;
  const bool flat = tiles < 4 * WPB;   // short segments: flatten (chunk, tile) so no warp of a block idles
  dim3 grid(flat ? (tiles * nchunks + WPB - 1) / WPB : (tiles + WPB - 1) / WPB, flat ? 1 : nchunks, a.B), block(WPB * 32);
  const size_t smem = (size_t)WPB * 8 * (RT + 11) * 8 * sizeof(T);
  auto kern = act1d_c8_v3_kernel<T, RT, WPB, PRECISE, PACKED, MINB, SPLIT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, block, smem, s>>>((const T*)a.x, (T*)a.y, a.alpha, a.inv_beta, a.seg, a.R, flat ? tiles : 0, nchunks);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_act_c8_v2(const ActArgs& a, int dtype, bool precise, int rt, cudaStream_t s) {
  if (a.B <= 0 || a.max_len <= 0) return cudaSuccess;
  if (dtype == 0) {
    if (a.split_out || a.fast_fp32) {
      // fp32 tensor-core mode: the fast form of SnakeBeta, s = u + h - h cos(2 alpha u), with the MUFU cosine of the
      // Cody-Waite-reduced argument.  Measured on cfg2 (109 launches, profiles/r2g_fp32tc.md): sinf 44.9 ms, polynomial
      // cosine 26.2 ms, plain MUFU 18.0 ms -- and the waveform error is the same 8.8e-5 for all of them (the tensor core's
      // accumulator truncation dominates).  BVG_TC32_ACT = 1 sinf, 2 polynomial, 0 unreduced MUFU select the others.
      static const int sel = [] { const char* e = getenv("BVG_TC32_ACT"); return e ? atoi(e) : 3; }();
      if (!precise) return cudaErrorInvalidValue;
      if (a.split_out) {
        switch (sel) {
          case 0: return launch_v3<float, 16, 4, 0, false, 1, true>(a, s);
          case 1: return launch_v3<float, 16, 4, 1, false, 1, true>(a, s);
          case 2: return launch_v3<float, 16, 4, 2, false, 1, true>(a, s);
          default: return launch_v3<float, 16, 4, 3, false, 1, true>(a, s);
        }
      }
      switch (sel) {
        case 0: return launch_v3<float, 16, 4, 0>(a, s);
        case 1: return launch_v3<float, 16, 4, 1>(a, s);
        case 2: return launch_v3<float, 16, 4, 2>(a, s);
        default: return launch_v3<float, 16, 4, 3>(a, s);
      }
    }
    return precise ? launch_v3<float, 16, 4, true>(a, s) : launch_v3<float, 16, 4, false>(a, s);
  }
  if (a.split_out) return cudaErrorInvalidValue;
  if (precise) return launch_v3<__nv_bfloat16, 16, 8, true>(a, s);
  static const int packed = [] { const char* e = getenv("BVG_ACT_PACKED"); return e ? atoi(e) : 1; }();
  if (packed) {
    // warps per block x resident blocks -> registers per thread.  Measured on cfg2 (round 1): 4 warps x 6
    // blocks (24 warps/SM at 80 registers, 24 bytes of spill) beats 8 x 2 (16 warps at 128 registers) by
    // 6 % overall -- smaller blocks fill the SMs better on the short early stages; the long stages sit on
    // the FP32-pipe plateau either way.  BVG_ACT_OCC selects the other variants for experiments.
    static const int occ = [] { const char* e = getenv("BVG_ACT_OCC"); return e ? atoi(e) : 3; }();
    if (rt == 32 && occ == 1) return launch_v3<__nv_bfloat16, 32, 6, false, true, 3>(a, s);   // 18 warps, <= 112 regs
    if (rt == 32 && occ == 2) return launch_v3<__nv_bfloat16, 32, 5, false, true, 4>(a, s);   // 20 warps, <= 96 regs
    if (rt == 32 && occ == 3) return launch_v3<__nv_bfloat16, 32, 4, false, true, 6>(a, s);   // 24 warps, <= 80 regs
    if (rt == 32 && occ == 4) return launch_v3<__nv_bfloat16, 32, 4, false, true, 5>(a, s);   // 20 warps, <= 96 regs
    if (rt == 16) return launch_v3<__nv_bfloat16, 16, 8, false, true>(a, s);
    if (rt == 24) return launch_v3<__nv_bfloat16, 24, 8, false, true>(a, s);
    return launch_v3<__nv_bfloat16, 32, 8, false, true>(a, s);
  }
  if (rt == 16) return launch_v3<__nv_bfloat16, 16, 8, false>(a, s);
  if (rt == 24) return launch_v3<__nv_bfloat16, 24, 8, false>(a, s);
  return launch_v3<__nv_bfloat16, 32, 8, false>(a, s);
}
