// ECAPA-TDNN speaker encoder, fp32 CUDA-core kernels + host orchestration.  See bvg_ecapa.cuh.
// Layout here is the reference's NCT ([B, C, T], T contiguous); tensors are tiny (<= 4608 x Tm).
#include "bvg_ecapa.cuh"

#include <cstdio>

namespace {

constexpr int TT = 64, TC = 32, CI = 8, MAXPAD = 8;   // econv tile: 64 time steps x 32 output channels, 8 input channels per step

struct EConv {
  const float* x; long long xb;     // input slice: element (b, c, t) at x[b*xb + c*T + t]
  const float* x2; long long x2b;   // optional addend with the same indexing (Res2Net: x_i + y_{i-1})
  const float* w;                   // [Cout, Cin, K]
  const float* bias;                // [Cout] or nullptr
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

__global__ void __launch_bounds__(256) econv_kernel(const EConv a) {
  __shared__ float xs[CI][TT + 2 * MAXPAD];
  __shared__ float ws[TC][CI * 5 + 1];
  const int t0 = blockIdx.x * TT, co0 = blockIdx.y * TC, b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pad = a.dil * (a.K - 1) / 2, width = TT + 2 * pad, K = a.K;
  const float* xb = a.x + (size_t)b * a.xb;
  const float* x2b = a.x2 ? a.x2 + (size_t)b * a.x2b : nullptr;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int ci0 = 0; ci0 < a.Cin; ci0 += CI) {
    for (int idx = tid; idx < CI * width; idx += 256) {
      const int c = idx / width, j = idx - c * width;
      float v = 0.f;
      if (ci0 + c < a.Cin) {
        const size_t off = (size_t)(ci0 + c) * a.T + reflect(t0 - pad + j, a.T);
        v = xb[off];
        if (x2b) v += x2b[off];
      }
      xs[c][j] = v;
    }
    for (int idx = tid; idx < TC * CI * K; idx += 256) {
      const int co = idx / (CI * K), r = idx - co * (CI * K);
      const int c = r / K, kk = r - c * K;
      float v = 0.f;
      if (co0 + co < a.Cout && ci0 + c < a.Cin) v = a.w[((size_t)(co0 + co) * a.Cin + ci0 + c) * K + kk];
      ws[co][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CI; ++c)
      for (int kk = 0; kk < K; ++kk) {
        const float w0 = ws[ty * 2][c * K + kk], w1 = ws[ty * 2 + 1][c * K + kk];
        const float* xp = &xs[c][tx * 4 + kk * a.dil];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[0][i] = fmaf(w0, xp[i], acc[0][i]); acc[1][i] = fmaf(w1, xp[i], acc[1][i]); }
      }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int co = co0 + ty * 2 + j;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
    const float sc = a.scale ? a.scale[co] : 1.f, sh = a.scale ? a.shift[co] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + tx * 4 + i;
      if (t >= a.T) continue;
      float v = acc[j][i] + bv;
      if (a.relu) v = fmaxf(v, 0.f);
      v = v * sc + sh;
      if (a.post == 1) v = tanhf(v);
      else if (a.post == 2) v = 1.f / (1.f + expf(-v));
      a.y[(size_t)b * a.yb + (size_t)co * a.T + t] = v;
    }
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

// attn input = cat([x, mean.expand, std.expand], dim=1)   (ECAPA_TDNN.py:318-321); st = [mean(C), std(C)] per batch
__global__ void ecat_kernel(const float* x, const float* st, float* y, int C, int T, int B) {
  const size_t n = (size_t)B * 3 * C * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int c = (int)((i / T) % (3 * C));
    const int b = (int)(i / ((size_t)T * 3 * C));
    y[i] = c < C ? x[((size_t)b * C + c) * T + t] : st[(size_t)b * 2 * C + (c - C)];
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
  void conv(const float* x, long long xb, const float* x2, long long x2b, const float* w, const float* bias, const float* scale,
            const float* shift, float* y, long long yb, int Cin, int Cout, int T_, int K, int dil, int relu, int post) {
    if (K > 5 || dil * (K - 1) / 2 > MAXPAD) { err = cudaErrorInvalidValue; return; }
    EConv a{x, xb, x2, x2b, w, bias, scale, shift, y, yb, Cin, Cout, T_, K, dil, relu, post};
    dim3 grid((T_ + TT - 1) / TT, (Cout + TC - 1) / TC, B);
    econv_kernel<<<grid, 256, 0, s>>>(a);
    check();
  }
  void tdnn(const EcapaTdnn& L, const float* x, long long xb, const float* x2, long long x2b, float* y, long long yb, int post = 0) {
    conv(x, xb, x2, x2b, L.w, L.b, L.scale, L.shift, y, yb, L.cin, L.cout, T, L.k, L.d, 1, post);
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
  L.cat = take(BT * 9 * C);
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
    r.conv(ws + L.sev, C, nullptr, 0, bl.se1.w, bl.se1.b, nullptr, nullptr, ws + L.seh, m.se, C, m.se, 1, 1, 1, 1, 0);
    r.conv(ws + L.seh, m.se, nullptr, 0, bl.se2.w, bl.se2.b, nullptr, nullptr, ws + L.ses, C, m.se, C, 1, 1, 1, 0, 2);
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
  ecat_kernel<<<r.blocks((size_t)B * 3 * C3 * T), 256, 0, s>>>(ws + L.mfa, ws + L.st, ws + L.cat, C3, T, B);
  r.check();
  r.tdnn(m.asp_tdnn, ws + L.cat, 3LL * C3 * T, nullptr, 0, ws + L.a1, (long long)m.att * T, 1 /* tanh */);
  r.conv(ws + L.a1, (long long)m.att * T, nullptr, 0, m.asp_conv.w, m.asp_conv.b, nullptr, nullptr, ws + L.a2, C3T, m.att, C3, T, 1, 1,
         0, 0);
  esoftmax_kernel<<<dim3(C3, B), 128, 0, s>>>(ws + L.a2, rel_lens, C3, T);
  r.check();
  estats_kernel<<<dim3(C3, B), 128, 0, s>>>(ws + L.mfa, C3T, ws + L.a2, rel_lens, ws + L.pooled, 2 * C3, C3, T, 1);
  r.check();
  // asp_bn (:575) and the final 1x1 conv (:578), on a length-1 sequence
  eaffine_kernel<<<(B * 2 * C3 + 255) / 256, 256, 0, s>>>(ws + L.pooled, m.abn_scale, m.abn_shift, ws + L.pooled_bn, 2 * C3, B);
  r.check();
  r.conv(ws + L.pooled_bn, 2 * C3, nullptr, 0, m.fc.w, m.fc.b, nullptr, nullptr, emb, m.lin, 2 * C3, m.lin, 1, 1, 1, 0, 0);
  if (launches) *launches = r.launches;
  return r.err;
}
