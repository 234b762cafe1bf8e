// CUDA-core (FFMA, fp32 accumulate) implicit-GEMM 1-D convolution over the packed c8 layout.
//
// This is the fp32 PARITY path (BVG_MODE_FP32) and the fallback-free reference for the tcgen05
// kernel's tests; it also runs the few layers the tensor-core kernel does not cover.
//
// One kernel covers the three convolution kinds of the generator (reference models.py):
//   * Conv1d k in {3,7,11}, dilation d in {1,3,5}, zero "same" padding (models.py:26-41,149):
//         D[q, co] = sum_j sum_ci X[q + (j - (k-1)/2) d, ci] W[co, ci, j]
//   * ConvTranspose1d (k,u,p) (models.py:155-161), as u interleaved phase convolutions
//     (SURVEY.md 8a closed form):  D[q, (phi,co)] = sum_m sum_ci X[q - m, ci] W[ci, co, phi + m u],
//     stored to output row  q u + phi - p.
// Both are the same GEMM: M = time rows q, N = u*Cout columns, K = taps x Cin, the taps being
// row shifts `tap_off[j]` of the activation tile; weights come pre-packed as [tap][Cin][N] fp32.
#include "bvg_common.cuh"

namespace {

constexpr int TM = 128, TN = 64, NT = 256;
constexpr int MAXSPAN = 50;                 // (k-1) d  <=  10 * 5
constexpr int XW = TM + MAXSPAN + 2;

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = r;
}

template <typename T>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const ConvArgs a) {
  __shared__ float xs[8][XW];
  __shared__ __align__(16) float ws[BVG_MAX_TAPS][8][TN];

  const int b = blockIdx.z;
  const SegDesc si = a.seg_in[b], so = a.seg_out[b];
  const int q0 = blockIdx.x * TM;
  const int Lq = si.len + a.q_extra;
  if (q0 >= Lq) return;
  const int N = a.u * a.Cout;
  const int n0 = blockIdx.y * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  int minoff = a.tap_off[0], maxoff = a.tap_off[0];
  for (int j = 1; j < a.ntaps; ++j) { minoff = min(minoff, a.tap_off[j]); maxoff = max(maxoff, a.tap_off[j]); }
  const int xrows = TM + (maxoff - minoff);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* xg = reinterpret_cast<const T*>(a.x);
  const float* wg = reinterpret_cast<const float*>(a.w);
  const int nchunks = a.Cin >> 3;
  for (int kc = 0; kc < nchunks; ++kc) {
    // stage the activation tile (8 channels, xrows rows), channel-major
    const T* xb = xg + ((size_t)kc * a.Rx + si.off + q0 + minoff) * 8;
    for (int r = tid; r < xrows; r += NT) {
      Vec8<T> v;
      v.load(xb + (size_t)r * 8);
#pragma unroll
      for (int c = 0; c < 8; ++c) xs[c][r] = v.v[c];
    }
    // stage the weights of this channel group: [tap][8][TN]
    for (int idx = tid; idx < a.ntaps * 8 * (TN / 4); idx += NT) {
      int n4 = idx % (TN / 4), rest = idx / (TN / 4);
      int c = rest & 7, tap = rest >> 3;
      int n = n0 + n4 * 4;
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) w4 = *reinterpret_cast<const float4*>(wg + ((size_t)tap * a.Cin + kc * 8 + c) * N + n);
      *reinterpret_cast<float4*>(&ws[tap][c][n4 * 4]) = w4;
    }
    __syncthreads();
    for (int tap = 0; tap < a.ntaps; ++tap) {
      const int o = a.tap_off[tap] - minoff + ty * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 w4 = *reinterpret_cast<const float4*>(&ws[tap][c][tx * 4]);
        const float* xr = &xs[c][o];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = xr[i];
          acc[i][0] = fmaf(av, w4.x, acc[i][0]);
          acc[i][1] = fmaf(av, w4.y, acc[i][1]);
          acc[i][2] = fmaf(av, w4.z, acc[i][2]);
          acc[i][3] = fmaf(av, w4.w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue: bias (+ per-segment speaker conditioning), residual, scale, optional accumulate
  const int n = n0 + tx * 4;
  if (n >= N) return;
  const int phase = n / a.Cout, co = n - phase * a.Cout;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
    const float* bp = a.bias + (size_t)b * a.bias_bstride + co;
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[j] = bp[j];
  }
  T* yg = reinterpret_cast<T*>(a.y);
  const T* rg = reinterpret_cast<const T*>(a.res);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + ty * 8 + i;
    const int orow = q * a.u + phase - a.p;
    if (q < Lq && orow >= 0 && orow < so.len) {
      const size_t o = ((size_t)(co >> 3) * a.Ry + so.off + orow) * 8 + (co & 7);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bs[j];
      if (rg) {
        float r4[4];
        load4<T>(rg + o, r4);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += r4[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= a.out_scale;
      if (a.accumulate) {
        float r4[4];
        load4<T>(yg + o, r4);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += r4[j];
      }
      store4<T>(yg + o, v);
    }
  }
}

}  // namespace

cudaError_t launch_conv_simt(const ConvArgs& a, int dtype, cudaStream_t s) {
  if (a.B <= 0 || a.max_q <= 0) return cudaSuccess;
  const int N = a.u * a.Cout;
  dim3 grid((a.max_q + TM - 1) / TM, (N + TN - 1) / TN, a.B), block(NT);
  if (dtype == 0)
    conv_simt_kernel<float><<<grid, block, 0, s>>>(a);
  else if (dtype == 1)
    conv_simt_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(a);
  else
    return cudaErrorInvalidValue;   // fp16 storage exists for the tcgen05 kernel only
  return cudaGetLastError();
}
