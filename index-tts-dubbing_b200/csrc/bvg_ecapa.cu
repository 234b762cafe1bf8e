// ECAPA-TDNN speaker encoder, fp32 CUDA-core kernels + host orchestration.  See bvg_ecapa.cuh.
// Layout here is the reference's NCT ([B, C, T], T contiguous); tensors are tiny (<= 4608 x Tm).
#include "bvg_ecapa.cuh"

#include <cstdio>

namespace {

constexpr int BM = 64, BN = 32, BK = 16, MAXPAD = 8;   // econv tile: 64 output channels x 32 time steps, 16 reduction steps per stage

struct EConv {
  const float* x; long long xb;     // input slice: element (b, c, t) at x[b*xb + c*T + t]
  const float* x2; long long x2b;   // optional addend with the same indexing (Res2Net: x_i + y_{i-1})
  const float* w;                   // row co at w[co*wrs + wco ...]: Cin*K reduction weights, (ci, k) k fastest
  long long wrs; int wco;
  const float* bias;                // [Cout] or nullptr
  const float* bias_b;              // optional per-batch addend [B][Cout] (time-constant input channels folded into a bias)
  const float* scale; const float* shift;   // optional per-channel affine applied after the ReLU (folded BatchNorm)
  float* y; long long yb;
  int Cin, Cout, T, K, dil, relu, post;     // post: 0 none, 1 tanh, 2 sigmoid
};

// SpeechBrain "same" padding with padding_mode="reflect" (nnet/CNN.py:430-433, :519-545)
__device__ __forceinline__ int reflect(int t, int T) {
  if (t < 0) t = -t;
  if (t >= T) t = 2 * (T - 1) - t;
  return min(max(t, 0), T - 1);
}

// Implicit-GEMM Conv1d on the FP32 pipe: D[co, t] = sum_r W[co, r] * X[r, t], r = (ci, k).  128 threads, each a 4 x 4
// register tile; the next stage's operands are fetched into registers while the current one is multiplied (one barrier
// per stage).  Grid: (time tiles, channel tiles, batch) -- 128 blocks for the 512-channel layers of a 511-frame prompt.
__global__ void __launch_bounds__(128) econv_kernel(const EConv a) {
  __shared__ __align__(16) float Ws[2][BK][BM + 4];
  __shared__ __align__(16) float Xs[2][BK][BN + 4];
  const int t0 = blockIdx.x * BN, co0 = blockIdx.y * BM, b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int K = a.K, R = a.Cin * K, pad = a.dil * (K - 1) / 2;
  const float* xb = a.x + (size_t)b * a.xb;
  const float* x2b = a.x2 ? a.x2 + (size_t)b * a.x2b : nullptr;
  // operand fetch assignment: W -- row wr (of 64), 8 consecutive r from wh*8; X -- row xr (of 16), 4 consecutive t from xt
  const int wr = tid >> 1, wh = tid & 1, xr = tid >> 3, xt = (tid & 7) * 4;
  const bool wrow_ok = co0 + wr < a.Cout;
  const float* wrow = a.w + (size_t)(co0 + wr) * a.wrs + a.wco;
  const bool wvec = ((a.wrs | (long long)a.wco) & 3) == 0;   // 16-byte aligned weight rows
  float wreg[8], xreg[4];
  auto fetch = [&](int r0) {
    const int rw = r0 + wh * 8;
    if (wvec && wrow_ok && rw + 8 <= R) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(wrow + rw)), v1 = __ldg(reinterpret_cast<const float4*>(wrow + rw + 4));
      wreg[0] = v0.x; wreg[1] = v0.y; wreg[2] = v0.z; wreg[3] = v0.w; wreg[4] = v1.x; wreg[5] = v1.y; wreg[6] = v1.z; wreg[7] = v1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) wreg[j] = (wrow_ok && rw + j < R) ? __ldg(wrow + rw + j) : 0.f;
    }
    const int r = r0 + xr;
    if (r < R) {
      const int ci = r / K, kk = r - ci * K;
      const float* xc = xb + (size_t)ci * a.T;
      const float* x2c = x2b ? x2b + (size_t)ci * a.T : nullptr;
      const int tb = t0 + xt + kk * a.dil - pad;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = reflect(tb + j, a.T);
        float v = xc[t];
        if (x2c) v += x2c[t];
        xreg[j] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) xreg[j] = 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) Ws[buf][wh * 8 + j][wr] = wreg[j];
    *reinterpret_cast<float4*>(&Xs[buf][xr][xt]) = make_float4(xreg[0], xreg[1], xreg[2], xreg[3]);
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int r0 = 0; r0 < R; r0 += BK) {
    const bool more = r0 + BK < R;
    if (more) fetch(r0 + BK);
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[buf][r][ty * 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[buf][r][tx * 4]);
      const float wa[4] = {wv.x, wv.y, wv.z, wv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
    float bv = a.bias ? a.bias[co] : 0.f;
    if (a.bias_b) bv += a.bias_b[(size_t)b * a.Cout + co];
    const float sc = a.scale ? a.scale[co] : 1.f, sh = a.scale ? a.shift[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t >= a.T) continue;
      float v = acc[i][j] + bv;
      if (a.relu) v = fmaxf(v, 0.f);
      v = v * sc + sh;
      if (a.post == 1) v = tanhf(v);
      else if (a.post == 2) v = 1.f / (1.f + expf(-v));
      a.y[(size_t)b * a.yb + (size_t)co * a.T + t] = v;
    }
  }
}

// The same convolution for a SHORT reduction (R = Cin * K <= 192: the 21 sequential 64 -> 64, k = 3 Res2Net convolutions of a
// prompt).  The pipelined kernel above pays one exposed global-load latency per 16-step stage (12 stages, 17 us per launch);
// here a block loads its whole weight slab [R][32] and input slab [R][16] once -- one latency -- and then only computes.
// 128 threads, 32 x 16 output tile, 2 x 2 outputs per thread.
constexpr int SR = 192, SBM = 32, SBN = 16;
__global__ void __launch_bounds__(128) econv_small_kernel(const EConv a) {
  __shared__ float Ws[SR][SBM + 1];
  __shared__ float Xs[SR][SBN + 1];
  const int t0 = blockIdx.x * SBN, co0 = blockIdx.y * SBM, b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int K = a.K, R = a.Cin * K, pad = a.dil * (K - 1) / 2;
  const float* xb = a.x + (size_t)b * a.xb;
  const float* x2b = a.x2 ? a.x2 + (size_t)b * a.x2b : nullptr;
#pragma unroll 8
  for (int idx = tid; idx < R * SBM; idx += 128) {          // consecutive threads read consecutive r of one weight row
    const int row = idx / R, r = idx - row * R;
    Ws[r][row] = co0 + row < a.Cout ? __ldg(a.w + (size_t)(co0 + row) * a.wrs + a.wco + r) : 0.f;
  }
#pragma unroll 8
  for (int idx = tid; idx < R * SBN; idx += 128) {
    const int r = idx / SBN, j = idx - r * SBN;
    const int ci = r / K, kk = r - ci * K;
    const int t = reflect(t0 + j + kk * a.dil - pad, a.T);
    float v = xb[(size_t)ci * a.T + t];
    if (x2b) v += x2b[(size_t)ci * a.T + t];
    Xs[r][j] = v;
  }
  __syncthreads();
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 8
  for (int r = 0; r < R; ++r) {
    const float w0 = Ws[r][ty * 2], w1 = Ws[r][ty * 2 + 1], x0 = Xs[r][tx * 2], x1 = Xs[r][tx * 2 + 1];
    acc[0][0] = fmaf(w0, x0, acc[0][0]); acc[0][1] = fmaf(w0, x1, acc[0][1]);
    acc[1][0] = fmaf(w1, x0, acc[1][0]); acc[1][1] = fmaf(w1, x1, acc[1][1]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int co = co0 + ty * 2 + i;
    if (co >= a.Cout) continue;
    float bv = a.bias ? a.bias[co] : 0.f;
    if (a.bias_b) bv += a.bias_b[(size_t)b * a.Cout + co];
    const float sc = a.scale ? a.scale[co] : 1.f, sh = a.scale ? a.shift[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int t = t0 + tx * 2 + j;
      if (t >= a.T) continue;
      float v = acc[i][j] + bv;
      if (a.relu) v = fmaxf(v, 0.f);
      v = v * sc + sh;
      if (a.post == 1) v = tanhf(v);
      else if (a.post == 2) v = 1.f / (1.f + expf(-v));
      a.y[(size_t)b * a.yb + (size_t)co * a.T + t] = v;
    }
  }
}

// Length-1 "convolutions" (SE block, final fc) and time-constant input channels folded into a bias: one warp per output,
// y[b, co] = post(relu?(sum_ci w[co*wrs + wco + ci] * x[b*xb + ci] + bias[co]))
__global__ void __launch_bounds__(256) egemv_kernel(const float* __restrict__ x, long long xb, const float* __restrict__ w, long long wrs,
                                                    int wco, const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                                                    int relu, int post) {
  const int co = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
  if (co >= Cout) return;
  const float* wr = w + (size_t)co * wrs + wco;
  const float* xr = x + (size_t)b * xb;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int ci = lane;
  for (; ci + 96 < Cin; ci += 128) {
    s0 = fmaf(__ldg(wr + ci), xr[ci], s0); s1 = fmaf(__ldg(wr + ci + 32), xr[ci + 32], s1);
    s2 = fmaf(__ldg(wr + ci + 64), xr[ci + 64], s2); s3 = fmaf(__ldg(wr + ci + 96), xr[ci + 96], s3);
  }
  for (; ci < Cin; ci += 32) s0 = fmaf(__ldg(wr + ci), xr[ci], s0);
  float v = (s0 + s1) + (s2 + s3);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) {
    if (bias) v += bias[co];
    if (relu) v = fmaxf(v, 0.f);
    if (post == 2) v = 1.f / (1.f + expf(-v));
    y[(size_t)b * Cout + co] = v;
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}
// length_to_mask(lengths * L, max_len = L) (ECAPA_TDNN.py:16-61): the product is taken in fp32
__device__ __forceinline__ bool in_len(int t, const float* rel_lens, int b, int T) {
  return rel_lens == nullptr || (float)t < rel_lens[b] * (float)T;
}

// Weighted mean / std over time of one (b, c) row.  attn == nullptr: weights = mask / sum(mask)
// (SEBlock :228-242, AttentiveStatisticsPooling :300-316); else weights = attn[b, c, :] (:330-334).
// mean -> out[b*ob + c], std -> out[b*ob + C + c] (std skipped when want_std == 0).
__global__ void __launch_bounds__(128) estats_kernel(const float* x, long long xb, const float* attn, const float* rel_lens,
                                                     float* out, long long ob, int C, int T, int want_std) {
  __shared__ float red[4];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* xr = x + (size_t)b * xb + (size_t)c * T;
  const float* ar = attn ? attn + ((size_t)b * C + c) * T : nullptr;
  float wn = 1.f;
  if (!ar) {
    float cnt = 0.f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) cnt += in_len(t, rel_lens, b, T) ? 1.f : 0.f;
    wn = 1.f / block_sum(cnt, red);
  }
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float w = ar ? ar[t] : (in_len(t, rel_lens, b, T) ? wn : 0.f);
    s += w * xr[t];
  }
  const float mean = block_sum(s, red);
  if (threadIdx.x == 0) out[(size_t)b * ob + c] = mean;
  if (!want_std) return;
  float v = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float w = ar ? ar[t] : (in_len(t, rel_lens, b, T) ? wn : 0.f);
    const float d = xr[t] - mean;
    v += w * d * d;
  }
  const float var = block_sum(v, red);
  if (threadIdx.x == 0) out[(size_t)b * ob + C + c] = sqrtf(fmaxf(var, 1e-12f));
}

// masked softmax over time, in place (ECAPA_TDNN.py:327-329)
__global__ void __launch_bounds__(128) esoftmax_kernel(float* a, const float* rel_lens, int C, int T) {
  __shared__ float red[4];
  const int c = blockIdx.x, b = blockIdx.y;
  float* r = a + ((size_t)b * C + c) * T;
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (in_len(t, rel_lens, b, T)) mx = fmaxf(mx, r[t]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float e = in_len(t, rel_lens, b, T) ? expf(r[t] - mx) : 0.f;
    r[t] = e;
    s += e;
  }
  const float inv = 1.f / block_sum(s, red);
  for (int t = threadIdx.x; t < T; t += blockDim.x) r[t] *= inv;
}

// y[b,c,t] = s[b,c] * x[b,c,t] + res[b,c,t]   (SE scale + block residual, ECAPA_TDNN.py:242, :426);
// s == nullptr / res == nullptr degrade to a strided copy (Res2Net chunk 0, :185-186)
__global__ void escale_res_kernel(const float* x, long long xb, const float* s, const float* res, long long rb, float* y,
                                  long long yb, int C, int T, int B) {
  const size_t n = (size_t)B * C * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int c = (int)((i / T) % C);
    const int b = (int)(i / ((size_t)T * C));
    float v = x[(size_t)b * xb + (size_t)c * T + t];
    if (s) v *= s[(size_t)b * C + c];
    if (res) v += res[(size_t)b * rb + (size_t)c * T + t];
    y[(size_t)b * yb + (size_t)c * T + t] = v;
  }
}

// mel [B, T, M] -> [B, M, T]   (ECAPA_TDNN.py:556 x.transpose(1, 2))
__global__ void etranspose_kernel(const float* mel, float* y, int B, int T, int M) {
  const size_t n = (size_t)B * T * M;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int m = (int)((i / T) % M);
    const int b = (int)(i / ((size_t)T * M));
    y[i] = mel[((size_t)b * T + t) * M + m];
  }
}

__global__ void eaffine_kernel(const float* x, const float* scale, const float* shift, float* y, int C, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * C) y[i] = x[i] * scale[i % C] + shift[i % C];
}

__global__ void efold_bn_kernel(const float* w, const float* b, const float* m, const float* v, float* scale, float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float sc = w[i] / sqrtf(v[i] + 1e-5f);   // torch.nn.BatchNorm1d eps
    scale[i] = sc;
    shift[i] = b[i] - m[i] * sc;
  }
}

void reg_tdnn(const EcapaRegisterFn& reg, const std::string& name, EcapaTdnn& L, int cin, int cout, int k, int d) {
  L.cin = cin; L.cout = cout; L.k = k; L.d = d;
  reg(name + ".conv.conv.weight", &L.w, {cout, cin, k});
  reg(name + ".conv.conv.bias", &L.b, {cout});
  reg(name + ".norm.norm.weight", &L.bn_w, {cout});
  reg(name + ".norm.norm.bias", &L.bn_b, {cout});
  reg(name + ".norm.norm.running_mean", &L.bn_m, {cout});
  reg(name + ".norm.norm.running_var", &L.bn_v, {cout});
}
void reg_lin(const EcapaRegisterFn& reg, const std::string& name, EcapaLin& L, int cin, int cout) {
  L.cin = cin; L.cout = cout;
  reg(name + ".conv.weight", &L.w, {cout, cin, 1});
  reg(name + ".conv.bias", &L.b, {cout});
}

struct Runner {
  cudaStream_t s;
  int B, T;
  int launches = 0;
  cudaError_t err = cudaSuccess;
  void check() { if (err == cudaSuccess) err = cudaGetLastError(); ++launches; }
  void conv(const float* x, long long xb, const float* x2, long long x2b, const float* w, long long wrs, int wco, const float* bias,
            const float* bias_b, const float* scale, const float* shift, float* y, long long yb, int Cin, int Cout, int T_, int K,
            int dil, int relu, int post) {
    if (K > 5 || dil * (K - 1) / 2 > MAXPAD) { err = cudaErrorInvalidValue; return; }
    EConv a{x, xb, x2, x2b, w, wrs, wco, bias, bias_b, scale, shift, y, yb, Cin, Cout, T_, K, dil, relu, post};
    if (Cin * K <= SR && T_ > 1) {   // short reduction: everything preloaded (Res2Net convolutions)
      dim3 grid((T_ + SBN - 1) / SBN, (Cout + SBM - 1) / SBM, B);
      econv_small_kernel<<<grid, 128, 0, s>>>(a);
    } else {
      dim3 grid((T_ + BN - 1) / BN, (Cout + BM - 1) / BM, B);
      econv_kernel<<<grid, 128, 0, s>>>(a);
    }
    check();
  }
  // length-1 sequence: y[b, :] = post(relu?(W[:, wco : wco + Cin] x[b, :] + bias))
  void gemv(const float* x, long long xb, const float* w, long long wrs, int wco, const float* bias, float* y, int Cin, int Cout,
            int relu, int post) {
    egemv_kernel<<<dim3((Cout + 7) / 8, B), 256, 0, s>>>(x, xb, w, wrs, wco, bias, y, Cin, Cout, relu, post);
    check();
  }
  void tdnn(const EcapaTdnn& L, const float* x, long long xb, const float* x2, long long x2b, float* y, long long yb, int post = 0) {
    conv(x, xb, x2, x2b, L.w, (long long)L.cin * L.k, 0, L.b, nullptr, L.scale, L.shift, y, yb, L.cin, L.cout, T, L.k, L.d, 1, post);
  }
  int blocks(size_t n) const { size_t g = (n + 255) / 256; return (int)(g > 4096 ? 4096 : g); }
};

}  // namespace

void ecapa_register(EcapaModel& m, int n_mels, int lin_neurons, const EcapaRegisterFn& reg) {
  m.n_mels = n_mels; m.lin = lin_neurons;
  const int C = m.C, sub = C / m.scale;
  reg_tdnn(reg, "blocks.0", m.b0, n_mels, C, 5, 1);
  for (int i = 0; i < 3; ++i) {
    EcapaBlock& bl = m.blk[i];
    bl.d = i + 2;   // dilations 2, 3, 4 (ECAPA_TDNN.py:478)
    const std::string p = "blocks." + std::to_string(i + 1);
    reg_tdnn(reg, p + ".tdnn1", bl.tdnn1, C, C, 1, 1);
    for (int j = 0; j < m.scale - 1; ++j)
      reg_tdnn(reg, p + ".res2net_block.blocks." + std::to_string(j), bl.res[j], sub, sub, 3, bl.d);
    reg_tdnn(reg, p + ".tdnn2", bl.tdnn2, C, C, 1, 1);
    reg_lin(reg, p + ".se_block.conv1", bl.se1, C, m.se);
    reg_lin(reg, p + ".se_block.conv2", bl.se2, m.se, C);
  }
  reg_tdnn(reg, "mfa", m.mfa, 3 * C, 3 * C, 1, 1);
  reg_tdnn(reg, "asp.tdnn", m.asp_tdnn, 9 * C, m.att, 1, 1);
  reg_lin(reg, "asp.conv", m.asp_conv, m.att, 3 * C);
  reg("asp_bn.norm.weight", &m.abn_w, {6 * C});
  reg("asp_bn.norm.bias", &m.abn_b, {6 * C});
  reg("asp_bn.norm.running_mean", &m.abn_m, {6 * C});
  reg("asp_bn.norm.running_var", &m.abn_v, {6 * C});
  reg_lin(reg, "fc", m.fc, 6 * C, lin_neurons);
}

void ecapa_collect_bn(EcapaModel& m, std::vector<EcapaTdnn*>& t) {
  t.push_back(&m.b0);
  for (auto& bl : m.blk) {
    t.push_back(&bl.tdnn1);
    for (int j = 0; j < m.scale - 1; ++j) t.push_back(&bl.res[j]);
    t.push_back(&bl.tdnn2);
  }
  t.push_back(&m.mfa);
  t.push_back(&m.asp_tdnn);
}

cudaError_t ecapa_fold_bn(const float* w, const float* b, const float* mean, const float* var, float* scale, float* shift,
                          int C, cudaStream_t s) {
  efold_bn_kernel<<<(C + 255) / 256, 256, 0, s>>>(w, b, mean, var, scale, shift, C);
  return cudaGetLastError();
}

namespace {
struct Layout {
  size_t melT, x0, xcat, ta, tb, tc, sev, seh, ses, mfa, st, cat, a1, a2, pooled, pooled_bn, total;
};
Layout make_layout(const EcapaModel& m, int B, int T) {
  Layout L{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  const size_t C = (size_t)m.C, BT = (size_t)B * T;
  L.melT = take(BT * m.n_mels);
  L.x0 = take(BT * C);
  L.xcat = take(BT * 3 * C);
  L.ta = take(BT * C); L.tb = take(BT * C); L.tc = take(BT * C);
  L.sev = take((size_t)B * C); L.seh = take((size_t)B * m.se); L.ses = take((size_t)B * C);
  L.mfa = take(BT * 3 * C);
  L.st = take((size_t)B * 6 * C);
  L.cat = take((size_t)B * m.att);   // time-constant part of the attention TDNN's input, as a per-batch bias
  L.a1 = take(BT * m.att);
  L.a2 = take(BT * 3 * C);
  L.pooled = take((size_t)B * 6 * C);
  L.pooled_bn = take((size_t)B * 6 * C);
  L.total = o;
  return L;
}
}  // namespace

size_t ecapa_workspace_bytes(const EcapaModel& m, int B, int T) { return make_layout(m, B, T).total * sizeof(float); }

cudaError_t ecapa_forward(const EcapaModel& m, const float* mel, int B, int T, const float* rel_lens, float* emb, void* workspace,
                          cudaStream_t s, int* launches) {
  const Layout L = make_layout(m, B, T);
  float* ws = reinterpret_cast<float*>(workspace);
  const int C = m.C, sub = C / m.scale;
  const long long CT = (long long)C * T, C3T = 3LL * C * T;
  Runner r{s, B, T};
  float *melT = ws + L.melT, *x0 = ws + L.x0, *xcat = ws + L.xcat, *ta = ws + L.ta, *tb = ws + L.tb, *tc = ws + L.tc;
  // x.transpose(1, 2)  (:556)
  etranspose_kernel<<<r.blocks((size_t)B * T * m.n_mels), 256, 0, s>>>(mel, melT, B, T, m.n_mels);
  r.check();
  // blocks[0]: TDNN k5  (:559-566)
  r.tdnn(m.b0, melT, (long long)m.n_mels * T, nullptr, 0, x0, CT);
  const float* xin = x0;
  long long xin_b = CT;
  for (int i = 0; i < 3; ++i) {
    const EcapaBlock& bl = m.blk[i];
    float* xout = xcat + (size_t)i * C * T;   // block outputs land in their slice of the MFA input (:569 cat)
    // SERes2NetBlock.forward (:413-426)
    r.tdnn(bl.tdnn1, xin, xin_b, nullptr, 0, ta, CT);
    // Res2NetBlock.forward (:179-191): y_0 = x_0; y_1 = f_1(x_1); y_i = f_i(x_i + y_{i-1})
    escale_res_kernel<<<r.blocks((size_t)B * sub * T), 256, 0, s>>>(ta, CT, nullptr, nullptr, 0, tb, CT, sub, T, B);
    r.check();
    for (int j = 1; j < m.scale; ++j) {
      const float* xj = ta + (size_t)j * sub * T;
      const float* yprev = j >= 2 ? tb + (size_t)(j - 1) * sub * T : nullptr;
      r.tdnn(bl.res[j - 1], xj, CT, yprev, CT, tb + (size_t)j * sub * T, CT);
    }
    r.tdnn(bl.tdnn2, tb, CT, nullptr, 0, tc, CT);
    // SEBlock (:228-242): masked mean over time -> conv1 -> ReLU -> conv2 -> sigmoid -> scale
    estats_kernel<<<dim3(C, B), 128, 0, s>>>(tc, CT, nullptr, rel_lens, ws + L.sev, C, C, T, 0);
    r.check();
    r.gemv(ws + L.sev, C, bl.se1.w, C, 0, bl.se1.b, ws + L.seh, C, m.se, 1, 0);
    r.gemv(ws + L.seh, m.se, bl.se2.w, m.se, 0, bl.se2.b, ws + L.ses, m.se, C, 0, 2);
    escale_res_kernel<<<r.blocks((size_t)B * C * T), 256, 0, s>>>(tc, CT, ws + L.ses, xin, xin_b, xout, C3T, C, T, B);
    r.check();
    xin = xout;
    xin_b = C3T;
  }
  // multi-layer feature aggregation (:569-570)
  r.tdnn(m.mfa, xcat, C3T, nullptr, 0, ws + L.mfa, C3T);
  // AttentiveStatisticsPooling (:282-338), global_context=True
  const int C3 = 3 * C;
  estats_kernel<<<dim3(C3, B), 128, 0, s>>>(ws + L.mfa, C3T, nullptr, rel_lens, ws + L.st, 2 * C3, C3, T, 1);
  r.check();
  // attn input = cat([x, mean.expand, std.expand], dim=1) (:318-321): the 2*C3 time-constant channels contribute a
  // per-batch bias  W[:, C3:3*C3] [mean; std]  to the attention TDNN, so the concatenation is never materialised
  r.gemv(ws + L.st, 2 * C3, m.asp_tdnn.w, 3LL * C3, C3, nullptr, ws + L.cat, 2 * C3, m.att, 0, 0);
  r.conv(ws + L.mfa, C3T, nullptr, 0, m.asp_tdnn.w, 3LL * C3, 0, m.asp_tdnn.b, ws + L.cat, m.asp_tdnn.scale, m.asp_tdnn.shift, ws + L.a1,
         (long long)m.att * T, C3, m.att, T, 1, 1, 1, 1 /* tanh */);
  r.conv(ws + L.a1, (long long)m.att * T, nullptr, 0, m.asp_conv.w, m.att, 0, m.asp_conv.b, nullptr, nullptr, nullptr, ws + L.a2, C3T,
         m.att, C3, T, 1, 1, 0, 0);
  esoftmax_kernel<<<dim3(C3, B), 128, 0, s>>>(ws + L.a2, rel_lens, C3, T);
  r.check();
  estats_kernel<<<dim3(C3, B), 128, 0, s>>>(ws + L.mfa, C3T, ws + L.a2, rel_lens, ws + L.pooled, 2 * C3, C3, T, 1);
  r.check();
  // asp_bn (:575) and the final 1x1 conv (:578), on a length-1 sequence
  eaffine_kernel<<<(B * 2 * C3 + 255) / 256, 256, 0, s>>>(ws + L.pooled, m.abn_scale, m.abn_shift, ws + L.pooled_bn, 2 * C3, B);
  r.check();
  r.gemv(ws + L.pooled_bn, 2 * C3, m.fc.w, 2LL * C3, 0, m.fc.b, emb, 2 * C3, m.lin, 0, 0);
  if (launches) *launches = r.launches;
  return r.err;
}
