// Small kernels around the convolution / activation hot ops: layout packing, weight repacking,
// the speaker-conditioning projection folded into per-segment biases, and conv_post + tanh.
#include <algorithm>
#include <cstdlib>

#include "bvg_common.cuh"
#include "bvg_misc.cuh"

namespace {

// latent [B, Tmax, C] row-major (any float type)  ->  packed c8 [C/8][R][8].
// src_row != nullptr: ragged source -- segment b's frames are rows src_row[b] .. src_row[b] + len of one [sum len, C] matrix
template <typename TI, typename TO>
__global__ void pack_latent_kernel(const TI* __restrict__ x, TO* __restrict__ y, const SegDesc* __restrict__ seg,
                                   const int* __restrict__ src_row, int B, int Tmax, int C, int R) {
  const int nch = C >> 3;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * Tmax * nch;
  if (idx >= total) return;
  int chunk = idx % nch;
  size_t bt = idx / nch;
  int t = bt % Tmax, b = bt / Tmax;
  SegDesc sd = seg[b];
  if (t >= sd.len) return;
  const TI* p = x + ((src_row ? (size_t)src_row[b] : (size_t)b * Tmax) + t) * C + chunk * 8;
  Vec8<TO> v;
#pragma unroll
  for (int c = 0; c < 8; ++c) v.v[c] = to_f32(p[c]);
  v.store(y + ((size_t)chunk * R + sd.off + t) * 8);
}

// latent -> split fp16 tensor [2 C/8][R][8] = [hi | lo] (fp32 tensor-core mode)
template <typename TI>
__global__ void pack_latent_split_kernel(const TI* __restrict__ x, __half* __restrict__ y, const SegDesc* __restrict__ seg,
                                         const int* __restrict__ src_row, int B, int Tmax, int C, int R) {
  const int nch = C >> 3;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * Tmax * nch;
  if (idx >= total) return;
  int chunk = idx % nch;
  size_t bt = idx / nch;
  int t = bt % Tmax, b = bt / Tmax;
  SegDesc sd = seg[b];
  if (t >= sd.len) return;
  const TI* p = x + ((src_row ? (size_t)src_row[b] : (size_t)b * Tmax) + t) * C + chunk * 8;
  Vec8<__half> hi, lo;
#pragma unroll
  for (int c = 0; c < 8; ++c) split_f32(to_f32(p[c]), hi.v[c], lo.v[c]);
  hi.store(y + ((size_t)chunk * R + sd.off + t) * 8);
  lo.store(y + ((size_t)(nch + chunk) * R + sd.off + t) * 8);
}

// fp32 [B, C, T]  <->  packed c8 (test entry points only)
template <typename TO>
__global__ void nct_to_c8_kernel(const float* __restrict__ x, TO* __restrict__ y, const SegDesc* __restrict__ seg,
                                 int B, int C, int T, int R) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * C * T;
  if (idx >= total) return;
  int t = idx % T;
  size_t bc = idx / T;
  int c = bc % C, b = bc / C;
  y[((size_t)(c >> 3) * R + seg[b].off + t) * 8 + (c & 7)] = from_f32<TO>(x[idx]);
}
template <typename TI>
__global__ void c8_to_nct_kernel(const TI* __restrict__ x, float* __restrict__ y, const SegDesc* __restrict__ seg,
                                 int B, int C, int T, int R) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * C * T;
  if (idx >= total) return;
  int t = idx % T;
  size_t bc = idx / T;
  int c = bc % C, b = bc / C;
  y[idx] = to_f32(x[((size_t)(c >> 3) * R + seg[b].off + t) * 8 + (c & 7)]);
}

// out[b][co] = bias[co] + cond_b[co] + sum_k cond_w[co][k] spk[b % spkB][k]     (one warp per output)
// reference: models.py:192-197 (1x1 convs on the [B',512,1] embedding), :226, :233-234 (broadcast add)
__global__ void cond_bias_kernel(const float* __restrict__ bias, const float* __restrict__ cw,
                                 const float* __restrict__ cb, const float* __restrict__ spk, float* __restrict__ out,
                                 int C, int D, int B, int spkB, int out_bstride) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * C) return;
  int b = warp / C, co = warp - b * C;
  float acc = 0.f;
  if (cw) {
    const float* w = cw + (size_t)co * D;
    const float* s = spk + (size_t)(spkB == 1 ? 0 : b) * D;
    for (int k = lane; k < D; k += 32) acc = fmaf(w[k], s[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (lane == 0) out[(size_t)b * out_bstride + co] = acc + bias[co] + (cb ? cb[co] : 0.f);
}

// conv_post (C -> 1, k = 7, pad 3) + tanh  (models.py:184, :247-248), C <= 64.  wav [B][Lmax] fp32.
template <typename T>
__global__ void conv_post_tanh_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ wav, short* __restrict__ pcm, const SegDesc* __restrict__ seg,
                                      const int* __restrict__ dst_row, int hop, int C, int R, int Lmax) {
  __shared__ float ws[7][64];
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) {
    int j = i / C, c = i - j * C;
    ws[j][c] = w[c * 7 + j];   // w [1][C][7]
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Lmax) return;
  SegDesc sd = seg[b];
  float out = 0.f;
  if (t < sd.len) {
    float acc = bias[0];
    const int nch = C >> 3;
    for (int ch = 0; ch < nch; ++ch) {
      const T* p = x + ((size_t)ch * R + sd.off + t - 3) * 8;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        Vec8<T> v;
        v.load(p + j * 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = fmaf(v.v[c], ws[j][ch * 8 + c], acc);
      }
    }
    out = tanhf(acc);
  }
  // dense output [B][Lmax] (zeros beyond a segment's length) or, with dst_row, ragged: segment b's samples start at
  // dst_row[b] * hop of one [sum len] vector and nothing is written beyond its length
  size_t o = (size_t)b * Lmax + t;
  if (dst_row) {
    if (t >= sd.len) return;
    o = (size_t)dst_row[b] * hop + t;
  }
  if (wav) wav[o] = out;
  // 16-bit PCM as the reference's callers produce it: clamp(32767 * wav, -32767, 32767) then a truncating
  // cast (infer.py:462, :627, :650)
  if (pcm) pcm[o] = (short)__float2int_rz(fminf(fmaxf(__fmul_rn(32767.f, out), -32767.f), 32767.f));
}

// Conv1d weight [Cout][Cin][k] -> [k][Cin][Cout]
__global__ void repack_conv_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int k) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)Cout * Cin * k;
  if (idx >= total) return;
  int co = idx % Cout;
  size_t r = idx / Cout;
  int ci = r % Cin, j = r / Cin;
  wp[idx] = w[((size_t)co * Cin + ci) * k + j];
}
// ConvTranspose1d weight [Cin][Cout][k] -> [k/u][Cin][u*Cout],  wp[m][ci][phi*Cout+co] = w[ci][co][phi+m*u]
__global__ void repack_convt_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int k,
                                    int u) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)Cin * Cout * k;
  if (idx >= total) return;
  int N = u * Cout;
  int n = idx % N;
  size_t r = idx / N;
  int ci = r % Cin, m = r / Cin;
  int phi = n / Cout, co = n - phi * Cout;
  wp[idx] = w[((size_t)ci * Cout + co) * k + phi + m * u];
}

__global__ void snake_params_kernel(const float* __restrict__ la, const float* __restrict__ lb,
                                    float* __restrict__ alpha, float* __restrict__ inv_beta, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  alpha[i] = expf(la[i]);                       // activations.py:115-118 (alpha_logscale)
  inv_beta[i] = 1.0f / (expf(lb[i]) + 1e-9f);   // activations.py:120 (no_div_by_zero)
}

// Zero every row of a packed c8 buffer that belongs to no segment (the guard gaps and the tail
// slack).  gap g lies between segment g-1 and segment g; block = (gap, chunk).
__global__ void zero_guards_kernel(uint4* __restrict__ buf, const SegDesc* __restrict__ seg, int B, int R,
                                   int vec_per_row) {
  const int g = blockIdx.x, chunk = blockIdx.y;
  const int lo = g == 0 ? 0 : seg[g - 1].off + seg[g - 1].len;
  const int hi = g == B ? R : seg[g].off;
  uint4* p = buf + ((size_t)chunk * R + lo) * vec_per_row;
  const int n = (hi - lo) * vec_per_row;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = z;
}

// All buffers of a plan in ONE launch: block = (gap, chunk counted across the buffers).
__global__ void zero_guards_all_kernel(const GuardJobs jobs, int B) {
  const int g = blockIdx.x;
  int chunk = blockIdx.y, j = 0;
  while (chunk >= jobs.job[j].chunks) { chunk -= jobs.job[j].chunks; ++j; }
  const GuardJob jb = jobs.job[j];
  const SegDesc* seg = jb.seg;
  const int lo = g == 0 ? 0 : seg[g - 1].off + seg[g - 1].len;
  const int hi = g == B ? jb.R : seg[g].off;
  uint4* p = reinterpret_cast<uint4*>(jb.buf) + ((size_t)chunk * jb.R + lo) * jb.vec_per_row;
  const int n = (hi - lo) * jb.vec_per_row;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = z;
}

// packed fp32 [C/8][R][8] -> split fp16 [2 C/8][R][8]: hi = fp16(x) in chunk c, 2^11 (x - hi) in chunk C/8 + c (split_f32)
__global__ void split_c8_kernel(const float* __restrict__ x, __half* __restrict__ y, const SegDesc* __restrict__ seg,
                                int nch, int R, int max_len) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const SegDesc sd = seg[b];
  if (t >= sd.len || t >= max_len) return;
  const size_t row = (size_t)chunk * R + sd.off + t;
  Vec8<float> v;
  v.load(x + row * 8);
  Vec8<__half> hi, lo;
#pragma unroll
  for (int c = 0; c < 8; ++c) split_f32(v.v[c], hi.v[c], lo.v[c]);
  hi.store(y + row * 8);
  lo.store(y + ((size_t)(nch + chunk) * R + sd.off + t) * 8);
}

__global__ void scale_vec_kernel(const float* __restrict__ x, const float* __restrict__ sc, float* __restrict__ y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * sc[i];
}

inline int nblk(size_t n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

cudaError_t launch_split_c8(const float* x, void* y_split, const SegDesc* seg, int B, int C, int R, int max_len, cudaStream_t s) {
  if (B <= 0 || max_len <= 0) return cudaSuccess;
  dim3 grid(nblk(max_len, 256), C >> 3, B);
  split_c8_kernel<<<grid, 256, 0, s>>>(x, (__half*)y_split, seg, C >> 3, R, max_len);
  return cudaGetLastError();
}

// max |x| over n floats into *out (which the caller zeroed): non-negative floats order like their bit patterns
__global__ void absmax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
cudaError_t launch_absmax(const float* x, size_t n, float* out, cudaStream_t s) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 1024);
  absmax_kernel<<<blocks, 256, 0, s>>>(x, n, reinterpret_cast<unsigned*>(out));
  return cudaGetLastError();
}

cudaError_t launch_scale_vec(const float* x, const float* scale, float* y, int n, cudaStream_t s) {
  scale_vec_kernel<<<nblk(n, 256), 256, 0, s>>>(x, scale, y, n);
  return cudaGetLastError();
}

bool bvg_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("BVG_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

cudaError_t launch_pack_latent(const void* x, int in_dtype, void* y, int out_dtype, const SegDesc* seg,
                               const int* src_row, int B, int Tmax, int C, int R, cudaStream_t s) {
  size_t total = (size_t)B * Tmax * (C >> 3);
  if (!total) return cudaSuccess;
  dim3 g(nblk(total, 256)), blk(256);
#define PL(TI, TO) pack_latent_kernel<TI, TO><<<g, blk, 0, s>>>((const TI*)x, (TO*)y, seg, src_row, B, Tmax, C, R)
#define PLS(TI) pack_latent_split_kernel<TI><<<g, blk, 0, s>>>((const TI*)x, (__half*)y, seg, src_row, B, Tmax, C, R)
  if (out_dtype == 3) {   // split [hi | lo] fp16
    if (in_dtype == 0) PLS(float); else if (in_dtype == 1) PLS(__nv_bfloat16); else PLS(__half);
  } else
  if (out_dtype == 0) {
    if (in_dtype == 0) PL(float, float); else if (in_dtype == 1) PL(__nv_bfloat16, float); else PL(__half, float);
  } else if (out_dtype == 1) {
    if (in_dtype == 0) PL(float, __nv_bfloat16); else if (in_dtype == 1) PL(__nv_bfloat16, __nv_bfloat16); else PL(__half, __nv_bfloat16);
  } else {
    if (in_dtype == 0) PL(float, __half); else if (in_dtype == 1) PL(__nv_bfloat16, __half); else PL(__half, __half);
  }
#undef PL
#undef PLS
  return cudaGetLastError();
}

cudaError_t launch_nct_to_c8(const float* x, void* y, int out_dtype, const SegDesc* seg, int B, int C, int T, int R,
                             cudaStream_t s) {
  size_t total = (size_t)B * C * T;
  if (!total) return cudaSuccess;
  if (out_dtype == 0) nct_to_c8_kernel<float><<<nblk(total, 256), 256, 0, s>>>(x, (float*)y, seg, B, C, T, R);
  else if (out_dtype == 1) nct_to_c8_kernel<__nv_bfloat16><<<nblk(total, 256), 256, 0, s>>>(x, (__nv_bfloat16*)y, seg, B, C, T, R);
  else nct_to_c8_kernel<__half><<<nblk(total, 256), 256, 0, s>>>(x, (__half*)y, seg, B, C, T, R);
  return cudaGetLastError();
}

cudaError_t launch_c8_to_nct(const void* x, int in_dtype, float* y, const SegDesc* seg, int B, int C, int T, int R,
                             cudaStream_t s) {
  size_t total = (size_t)B * C * T;
  if (!total) return cudaSuccess;
  if (in_dtype == 0) c8_to_nct_kernel<float><<<nblk(total, 256), 256, 0, s>>>((const float*)x, y, seg, B, C, T, R);
  else if (in_dtype == 1) c8_to_nct_kernel<__nv_bfloat16><<<nblk(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, y, seg, B, C, T, R);
  else c8_to_nct_kernel<__half><<<nblk(total, 256), 256, 0, s>>>((const __half*)x, y, seg, B, C, T, R);
  return cudaGetLastError();
}

cudaError_t launch_cond_bias(const float* bias, const float* cw, const float* cb, const float* spk, float* out, int C,
                             int D, int B, int spkB, int out_bstride, cudaStream_t s) {
  size_t threads = (size_t)B * C * 32;
  cond_bias_kernel<<<nblk(threads, 256), 256, 0, s>>>(bias, cw, cb, spk, out, C, D, B, spkB, out_bstride);
  return cudaGetLastError();
}

cudaError_t launch_conv_post_tanh(const void* x, int dtype, const float* w, const float* bias, float* wav, short* pcm,
                                  const SegDesc* seg, const int* dst_row, int hop, int B, int C, int R, int Lmax,
                                  cudaStream_t s) {
  if (B <= 0 || Lmax <= 0) return cudaSuccess;
  dim3 g(nblk(Lmax, 256), B), blk(256);
  if (dtype == 0) conv_post_tanh_kernel<float><<<g, blk, 0, s>>>((const float*)x, w, bias, wav, pcm, seg, dst_row, hop, C, R, Lmax);
  else if (dtype == 1) conv_post_tanh_kernel<__nv_bfloat16><<<g, blk, 0, s>>>((const __nv_bfloat16*)x, w, bias, wav, pcm, seg, dst_row, hop, C, R, Lmax);
  else conv_post_tanh_kernel<__half><<<g, blk, 0, s>>>((const __half*)x, w, bias, wav, pcm, seg, dst_row, hop, C, R, Lmax);
  return cudaGetLastError();
}

cudaError_t launch_zero_guards_all(const GuardJobs& jobs, int B, cudaStream_t s) {
  int chunks = 0;
  for (int j = 0; j < jobs.n; ++j) chunks += jobs.job[j].chunks;
  if (!chunks) return cudaSuccess;
  dim3 g(B + 1, chunks);
  zero_guards_all_kernel<<<g, 128, 0, s>>>(jobs, B);
  return cudaGetLastError();
}

cudaError_t launch_repack_conv(const float* w, float* wp, int Cout, int Cin, int k, cudaStream_t s) {
  size_t total = (size_t)Cout * Cin * k;
  repack_conv_kernel<<<nblk(total, 256), 256, 0, s>>>(w, wp, Cout, Cin, k);
  return cudaGetLastError();
}
cudaError_t launch_repack_convt(const float* w, float* wp, int Cin, int Cout, int k, int u, cudaStream_t s) {
  size_t total = (size_t)Cin * Cout * k;
  repack_convt_kernel<<<nblk(total, 256), 256, 0, s>>>(w, wp, Cin, Cout, k, u);
  return cudaGetLastError();
}
cudaError_t launch_snake_params(const float* la, const float* lb, float* alpha, float* inv_beta, int n,
                                cudaStream_t s) {
  snake_params_kernel<<<nblk(n, 256), 256, 0, s>>>(la, lb, alpha, inv_beta, n);
  return cudaGetLastError();
}
cudaError_t launch_zero_guards(void* buf, int esize, const SegDesc* seg, int B, int C, int R, cudaStream_t s) {
  dim3 g(B + 1, C >> 3);
  zero_guards_kernel<<<g, 128, 0, s>>>((uint4*)buf, seg, B, R, esize == 4 ? 2 : 1);
  return cudaGetLastError();
}
