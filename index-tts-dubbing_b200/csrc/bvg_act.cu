// Fused anti-aliased SnakeBeta activation (Activation1d): 2x kaiser-sinc upsample -> snake ->
// 2x low-pass downsample, one pass over HBM (the 2x intermediate lives only in shared memory).
//
// Replaces the reference's alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-181 and
// follows the semantics of the reference's torch path (alias_free_torch/act.py:24-29,
// resample.py:10-49, filter.py:61-96, activations.py:109-122) -- replicate padding of the input
// for the up-FIR and of the ACTIVATED 2x signal for the down-FIR:
//
//   u[2i]   = 2 (f1 x[i+2] + f3 x[i+1] + f5 x[i] + f4 x[i-1] + f2 x[i-2] + f0 x[i-3])
//   u[2i+1] = 2 (f0 x[i+3] + f2 x[i+2] + f4 x[i+1] + f5 x[i] + f3 x[i-1] + f1 x[i-2])   (x clamped)
//   s[m]    = u[m] + sin^2(alpha u[m]) / (beta + 1e-9)
//   y[t]    = sum_{k<12} f[k] s[clamp(2t + k - 5, 0, 2L-1)]
//
// One CTA = one 8-channel group x TR consecutive time steps of one segment.  The raw tile
// (+-6 halo rows) is staged in shared memory channel-major so the FIRs read conflict-free
// along time; each warp owns one channel.  HBM traffic is the algorithmic 2 * elements.
#include <cstdlib>

#include "bvg_common.cuh"

namespace {

constexpr int TR = 256;             // output rows per CTA
constexpr int XROWS = TR + 12;      // staged input rows  (t0-6 .. t0+TR+5)
constexpr int XS = XROWS + 1;       // smem stride (odd: the transposing stores hit 8 distinct banks)
constexpr int NPAIR = TR + 6;       // (u[2i], u[2i+1]) pairs, i = t0-3 .. t0+TR+2
constexpr int SS = 2 * NPAIR + 4;   // smem stride of the activated 2x signal
constexpr int NTHREADS = 256;

template <bool PRECISE>
__device__ __forceinline__ float snake(float u, float alpha, float inv_beta) {
  float sn = PRECISE ? sinf(alpha * u) : __sinf(alpha * u);
  return fmaf(inv_beta * sn, sn, u);
}

template <typename T, bool NCT, bool PRECISE>
__global__ void __launch_bounds__(NTHREADS)
act1d_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ alpha,
             const float* __restrict__ inv_beta, const SegDesc* __restrict__ seg, int R, int C, int Tn) {
  __shared__ float xs[8][XS];
  __shared__ __align__(16) float ss[8][SS];

  const int b = blockIdx.z, chunk = blockIdx.y, t0 = blockIdx.x * TR;
  int off, L;
  if constexpr (NCT) { off = 0; L = Tn; } else { SegDesc sd = seg[b]; off = sd.off; L = sd.len; }
  if (t0 >= L) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- phase 1: stage x[clamp(t0-6+r)] for r < XROWS, channel-major ----------------------
  if constexpr (!NCT) {
    const T* xb = x + ((size_t)chunk * R + off) * 8;
    for (int r = tid; r < XROWS; r += NTHREADS) {
      int t = min(max(t0 - 6 + r, 0), L - 1);
      Vec8<T> v;
      v.load(xb + (size_t)t * 8);
#pragma unroll
      for (int c = 0; c < 8; ++c) xs[c][r] = v.v[c];
    }
  } else {
    for (int idx = tid; idx < 8 * XROWS; idx += NTHREADS) {
      int c = idx / XROWS, r = idx - c * XROWS;
      int ch = chunk * 8 + c;
      int t = min(max(t0 - 6 + r, 0), L - 1);
      xs[c][r] = (ch < C) ? to_f32(x[((size_t)b * C + ch) * Tn + t]) : 0.f;
    }
  }
  __syncthreads();

  // ---- phase 2: up-FIR + snake, warp <-> channel, lane <-> pair ---------------------------
  {
    const int c = warp, ch = chunk * 8 + c;
    const float al = (ch < C) ? alpha[ch] : 1.f, ib = (ch < C) ? inv_beta[ch] : 1.f;
    const float* xr = xs[c];
    for (int pi = lane; pi < NPAIR; pi += 32) {
      int i = t0 - 3 + pi;
      int ic = min(max(i, 0), L - 1);
      const float* p = xr + (ic - t0 + 6);   // p[d] = x[clamp(ic + d)]
      float xm3 = p[-3], xm2 = p[-2], xm1 = p[-1], x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3];
      float ue = BVG_F1 * x2;
      ue = fmaf(BVG_F3, x1, ue); ue = fmaf(BVG_F5, x0, ue); ue = fmaf(BVG_F4, xm1, ue);
      ue = fmaf(BVG_F2, xm2, ue); ue = fmaf(BVG_F0, xm3, ue);
      float uo = BVG_F0 * x3;
      uo = fmaf(BVG_F2, x2, uo); uo = fmaf(BVG_F4, x1, uo); uo = fmaf(BVG_F5, x0, uo);
      uo = fmaf(BVG_F3, xm1, uo); uo = fmaf(BVG_F1, xm2, uo);
      float se = snake<PRECISE>(2.f * ue, al, ib);
      float so = snake<PRECISE>(2.f * uo, al, ib);
      if (i < 0) so = se;        // replicate padding of the activated signal: s[m<0] = s[0]
      if (i >= L) se = so;       //                                            s[m>=2L] = s[2L-1]
      *reinterpret_cast<float2*>(&ss[c][2 * pi]) = make_float2(se, so);
    }
  }
  __syncthreads();

  // ---- phase 3: down-FIR (stride 2), results back into xs as ys[c][r] ----------------------
  {
    const int c = warp;
    for (int r = lane; r < TR; r += 32) {
      const float2* sp = reinterpret_cast<const float2*>(&ss[c][2 * r]);
      float2 a0 = sp[0], a1 = sp[1], a2 = sp[2], a3 = sp[3], a4 = sp[4], a5 = sp[5], a6 = sp[6];
      // taps k = 0..11 read ss[2r+1 .. 2r+12]
      float acc = BVG_F0 * (a0.y + a6.x);
      acc = fmaf(BVG_F1, a1.x + a5.y, acc);
      acc = fmaf(BVG_F2, a1.y + a5.x, acc);
      acc = fmaf(BVG_F3, a2.x + a4.y, acc);
      acc = fmaf(BVG_F4, a2.y + a4.x, acc);
      acc = fmaf(BVG_F5, a3.x + a3.y, acc);
      xs[c][r] = acc;
    }
  }
  __syncthreads();

  // ---- phase 4: write back ---------------------------------------------------------------
  if constexpr (!NCT) {
    T* yb = y + ((size_t)chunk * R + off) * 8;
    for (int r = tid; r < TR; r += NTHREADS) {
      int t = t0 + r;
      if (t < L) {
        Vec8<T> v;
#pragma unroll
        for (int c = 0; c < 8; ++c) v.v[c] = xs[c][r];
        v.store(yb + (size_t)t * 8);
      }
    }
  } else {
    for (int idx = tid; idx < 8 * TR; idx += NTHREADS) {
      int c = idx / TR, r = idx - c * TR;
      int ch = chunk * 8 + c, t = t0 + r;
      if (ch < C && t < L) y[((size_t)b * C + ch) * Tn + t] = from_f32<T>(xs[c][r]);
    }
  }
}

}  // namespace

cudaError_t launch_act_c8(const ActArgs& a, int dtype, bool precise, cudaStream_t s) {
  if (a.B <= 0 || a.max_len <= 0) return cudaSuccess;
  // 16-bit modes: the tensor-core kernel (both FIRs as warp-level MMAs, bvg_act3.cu).  It is used for EVERY length so that
  // a segment's samples do not depend on what else is in the batch (the kernels round differently).  fp32 parity mode (and
  // bf16 with BVG_ACT_MMA=0, a tuning aid): the register-streamed kernel (bvg_act2.cu), BVG_ACT_RT=16|24|32 rows per thread.
  static const int rt = [] { const char* e = getenv("BVG_ACT_RT"); int v = e ? atoi(e) : 32; return (v == 16 || v == 24) ? v : 32; }();
  static const int use_mma = [] { const char* e = getenv("BVG_ACT_MMA"); return e ? atoi(e) : 1; }();
  if (dtype == 2) return precise ? cudaErrorInvalidValue : launch_act_c8_mma(a, dtype, s);   // fp16 storage: tensor-core kernel only
  if (use_mma && dtype == 1 && !precise) return launch_act_c8_mma(a, dtype, s);
  if (a.prescaled || a.extra_jobs) return cudaErrorInvalidValue;   // only the tensor-core kernel has the pre-scaled variant / grouped launches
  return launch_act_c8_v2(a, dtype, precise, rt, s);
}

cudaError_t launch_act_nct(const void* x, void* y, const float* alpha, const float* inv_beta, int B, int C,
                           int T, int dtype, cudaStream_t s) {
  if (B <= 0 || C <= 0 || T <= 0) return cudaSuccess;
  dim3 grid((T + TR - 1) / TR, (C + 7) / 8, B), block(NTHREADS);
  if (dtype == 0)
    act1d_kernel<float, true, true><<<grid, block, 0, s>>>((const float*)x, (float*)y, alpha, inv_beta, nullptr, 0, C, T);
  else if (dtype == 1)
    act1d_kernel<__nv_bfloat16, true, false><<<grid, block, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, alpha, inv_beta, nullptr, 0, C, T);
  else
    act1d_kernel<__half, true, false><<<grid, block, 0, s>>>((const __half*)x, (__half*)y, alpha, inv_beta, nullptr, 0, C, T);
  return cudaGetLastError();
}
