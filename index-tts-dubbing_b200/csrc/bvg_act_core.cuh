// Device-side core of the fused Activation1d for bf16 rows staged in shared memory with a 16-byte row
// pitch (8 channels per row): used by the conv kernel's activation warps (bvg_conv_umma.cu), which
// turn the raw input tile into the activated UMMA A operand without a round trip through HBM.
// Formulas and reference citations: see bvg_act.cu (closed form of SURVEY.md section 8a); the
// stand-alone kernel with the same streaming structure is act1d_c8_v3_kernel in bvg_act2.cu.
#pragma once
#include "bvg_common.cuh"

namespace actcore {

typedef unsigned long long u64;

constexpr float G0 = 2.f * BVG_F0, G1 = 2.f * BVG_F1, G2 = 2.f * BVG_F2, G3 = 2.f * BVG_F3, G4 = 2.f * BVG_F4,
                G5 = 2.f * BVG_F5;

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
// channel pair (2 bf16 = 4 bytes) of one row
__device__ __forceinline__ u64 ldpair(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void stpair(__nv_bfloat16* p, float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h);
}

// Interior streaming pass: RT consecutive outputs of one channel pair.
//   xr: raw rows, xr[n*8] = x[r0 - 5 + n], n in [0, RT+10)      yr: outputs, yr[t*8] = y[r0 + t]
//   a0,a1 = 2*alpha, h0,h1 = inv_beta/2 (s = u + h - h cos(2 alpha u), the +h is added per output).
template <int RT>
__device__ __forceinline__ void stream_packed(const __nv_bfloat16* xr, __nv_bfloat16* yr, float a0, float a1, float h0,
                                              float h1) {
  u64 xw[RT + 10];
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
  auto up_odd = [&](int j) {   // activated sample 2i+1 of pair i = r0-3+j: taps on xw[j..j+5]
    u64 u = mul2(GG1, xw[j]);
    u = fma2(GG3, xw[j + 1], u); u = fma2(GG5, xw[j + 2], u); u = fma2(GG4, xw[j + 3], u);
    u = fma2(GG2, xw[j + 4], u); u = fma2(GG0, xw[j + 5], u);
    return snake2(u);
  };
  auto up_even = [&](int j) {  // activated sample 2i of pair i = r0-3+j: taps on xw[j-1..j+4]
    u64 u = mul2(GG0, xw[j - 1]);
    u = fma2(GG2, xw[j], u); u = fma2(GG4, xw[j + 1], u); u = fma2(GG5, xw[j + 2], u);
    u = fma2(GG3, xw[j + 3], u); u = fma2(GG1, xw[j + 4], u);
    return snake2(u);
  };
#pragma unroll
  for (int n = 0; n < 10; ++n) xw[n] = ldpair(xr + n * 8);
#pragma unroll
  for (int n = 0; n < 10; ++n) s[n] = (n & 1) ? up_even((n + 1) / 2) : up_odd(n / 2);
#pragma unroll
  for (int t = 0; t < RT; ++t) {
    xw[t + 10] = ldpair(xr + (t + 10) * 8);
    s[2 * t + 10] = up_odd(t + 5);
    s[2 * t + 11] = up_even(t + 6);
    u64 acc = mul2(FF0, add2(s[2 * t], s[2 * t + 11]));
    acc = fma2(FF1, add2(s[2 * t + 1], s[2 * t + 10]), acc);
    acc = fma2(FF2, add2(s[2 * t + 2], s[2 * t + 9]), acc);
    acc = fma2(FF3, add2(s[2 * t + 3], s[2 * t + 8]), acc);
    acc = fma2(FF4, add2(s[2 * t + 4], s[2 * t + 7]), acc);
    acc = fma2(FF5, add2(s[2 * t + 5], s[2 * t + 6]), acc);
    acc = add2(acc, HH);
    stpair(yr + t * 8, upk2(acc));
  }
}

// Exact output at segment time t (0 <= t < L) with both replicate paddings (input and activated 2x
// signal).  col0 points at this thread's channel pair in slab row 0; slab row r holds x[slab_t0 + r].
static __device__ __noinline__ float2 exact_clamped(const __nv_bfloat16* col0, int slab_t0, int slab_rows, int t, int L,
                                             float a0, float a1, float h0, float h1) {
  const float taps[6] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5};
  float accx = 0.f, accy = 0.f;
#pragma unroll 1
  for (int k = 0; k < 12; ++k) {
    const int m = min(max(2 * t - 5 + k, 0), 2 * L - 1);   // replicate pad of the activated signal
    const int i = m >> 1;
    float2 p[7];
#pragma unroll
    for (int d = -3; d <= 3; ++d) {
      int idx = min(max(i + d, 0), L - 1) - slab_t0;         // replicate pad of the input
      idx = min(max(idx, 0), slab_rows - 1);
      p[d + 3] = upk2(ldpair(col0 + idx * 8));
    }
    float ux, uy;
    if (m & 1) {   // u[2i+1]
      ux = G0 * p[6].x + G2 * p[5].x + G4 * p[4].x + G5 * p[3].x + G3 * p[2].x + G1 * p[1].x;
      uy = G0 * p[6].y + G2 * p[5].y + G4 * p[4].y + G5 * p[3].y + G3 * p[2].y + G1 * p[1].y;
    } else {       // u[2i]
      ux = G1 * p[5].x + G3 * p[4].x + G5 * p[3].x + G4 * p[2].x + G2 * p[1].x + G0 * p[0].x;
      uy = G1 * p[5].y + G3 * p[4].y + G5 * p[3].y + G4 * p[2].y + G2 * p[1].y + G0 * p[0].y;
    }
    const float f = taps[k < 6 ? k : 11 - k];
    accx = fmaf(f, fmaf(-h0, __cosf(a0 * ux), ux), accx);
    accy = fmaf(f, fmaf(-h1, __cosf(a1 * uy), uy), accy);
  }
  return make_float2(accx + h0, accy + h1);
}

}  // namespace actcore
