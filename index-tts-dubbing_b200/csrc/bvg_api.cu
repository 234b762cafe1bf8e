// C ABI of libb200vgan.so: generator handle, per-geometry plan, the forward launch sequence and
// the per-op entry points.  See include/b200vgan.h for the contract of every function.
//
// The launch sequence in bvg_forward mirrors the reference's BigVGAN.forward
// (indextts/BigVGAN/models.py:210-250) and AMPBlock1.forward (models.py:65-74) with
//   * the speaker conditioning  x + cond(spk)  (models.py:226, :233-234) folded into a per-segment
//     bias of conv_pre / the up-convolutions,
//   * the residual add  xt + x  (models.py:72) and the resblock mean  xs / 3  (models.py:237-243)
//     folded into the epilogue of each block's last convolution.
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200vgan.h"
#include "bvg_common.cuh"
#include "bvg_ecapa.cuh"
#include "bvg_misc.cuh"

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ConvLayer {
  bool transposed = false;
  int Cin = 0, Cout = 0, k = 0, d = 1, u = 1, p = 0;
  int ntaps = 0, N = 0, q_extra = 0;
  int tap_off[BVG_MAX_TAPS] = {0};
  float* w_raw = nullptr;   // reference layout, fp32 (device)
  float* bias = nullptr;    // [Cout]
  float* w_tap = nullptr;   // [ntaps][Cin][N] fp32
  void* w_umma = nullptr;   // bf16 UMMA shared-memory images
  void* w_umma16 = nullptr; // the same images in fp16 (BVG_MODE_F16)
  void* w_umma_s = nullptr; // 64-column n-tile images (small-batch variant; layers with N >= 256 only), bf16 / fp16
  void* w_umma16_s = nullptr;
  // 16-bit modes: per-channel factors folded into the UMMA images (see fold_activation_scales) and the matching bias
  const float* fold_out = nullptr;   // [Cout]
  const float* fold_in = nullptr;    // [Cin]
  float* bias_umma = nullptr;        // bias * fold_out (used instead of `bias` by the tcgen05 path)
  const float* fold_res = nullptr;   // [Cout] factor on the residual input (the residual stream is stored pre-scaled)
  float* fold_buf = nullptr;         // device storage owned by this layer for fold_out / fold_res products ([2 * Cout])
  bool k_packed = false;             // the UMMA images are K-packed (24-channel layers, bvg_conv_umma.cu)
  void* w_umma3 = nullptr;           // fp32 tensor-core mode: fp16 image of S [2^-11 W_hi; W_lo; W_hi] (3 Cin input channels), built on first use
  float tc32_acc_scale = 1.f;        // 1 / S
  void setup() {
    if (!transposed) {
      ntaps = k; N = Cout; u = 1; p = 0; q_extra = 0;
      for (int j = 0; j < k; ++j) tap_off[j] = (j - (k - 1) / 2) * d;
    } else {
      // output row q*u + phi - p: the last p rows of a segment need q up to len_in + ceil(p/u) - 1
      ntaps = k / u; N = u * Cout; p = (k - u) / 2; q_extra = (p + u - 1) / u;
      for (int m = 0; m < ntaps; ++m) tap_off[m] = -m;
    }
  }
};

struct ActLayer {
  int C = 0;
  float *la = nullptr, *lb = nullptr, *alpha = nullptr, *inv_beta = nullptr;
  bool prescaled = false;   // 16-bit modes: input / output carry the factor 2 alpha, folded into the neighbouring convolutions
};

struct CondLayer {
  int C = 0;
  float *w = nullptr, *b = nullptr;
};

struct ParamSlot {
  float** dst;
  std::vector<int64_t> shape;
  bool set = false;
};

enum { PROF_ACT = 0, PROF_CONV_TC = 1, PROF_CONV_CC = 2, PROF_OTHER = 3, PROF_NCLS = 4 };
struct ProfRec {
  cudaEvent_t e0, e1;
  int cls;
  double flops, bytes;
};

}  // namespace

struct bvg_handle {
  bvg_config cfg{};
  int device = 0;
  int hop = 1;
  int nups = 0, nk = 0, nd = 0;
  ConvLayer conv_pre, conv_post;
  std::vector<ConvLayer> ups;
  std::vector<ConvLayer> c1, c2;   // [(stage*nk + j)*nd + m]
  std::vector<ActLayer> acts;      // [(stage*nk + j)*2*nd + a]
  ActLayer act_post;
  CondLayer cond_pre;
  std::vector<CondLayer> conds;
  std::map<std::string, ParamSlot> params;
  // ECAPA-TDNN speaker encoder ("speaker_encoder.*" keys): optional -- a caller that supplies its own
  // embedding never uploads them; bvg_speaker_embedding needs all of them
  EcapaModel ecapa;
  std::map<std::string, ParamSlot> ecapa_params;
  bool ecapa_ready = false;
  std::vector<void*> owned;
  // Plan tables (segment descriptors, tile prefix sums) live in a handle-owned arena: device memory plus a pinned
  // host mirror, carved into slices that are recycled through per-size free lists.  Creating a plan therefore costs
  // no cudaMalloc and no synchronous copy once the arena is warm (ragged SRT batches make a new plan per batch).
  struct TableChunk { char* dev; char* host; size_t cap, used; };
  std::vector<TableChunk> tab_chunks;
  std::map<size_t, std::vector<std::pair<char*, char*>>> tab_free;   // slice bytes -> (device, host) slices
  int64_t plans_created = 0;
  int num_sms = 148;
  bool finalized = false;
  bool tc32_ready = false;  // the w_umma3 images exist (BVG_MODE_FP32_TC, built by the first forward that needs them)
  int launch_counter = 0;   // kernel launches issued by the forward in progress
  // optional per-launch CUDA-event timing (bench.py's roofline numbers)
  bool prof_on = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
};

namespace {
// Records a CUDA event pair around the launches issued inside its scope (only when profiling is on).
struct ProfScope {
  bvg_handle* h;
  cudaStream_t s;
  cudaEvent_t e1 = nullptr;
  ProfScope(bvg_handle* h_, cudaStream_t s_, int cls, double flops, double bytes) : h(h_), s(s_) {
    if (!h->prof_on) return;
    while (h->ev_pool.size() < h->ev_used + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      h->ev_pool.push_back(e);
    }
    cudaEvent_t e0 = h->ev_pool[h->ev_used++];
    e1 = h->ev_pool[h->ev_used++];
    cudaEventRecord(e0, s);
    h->prof.push_back(ProfRec{e0, e1, cls, flops, bytes});
  }
  ~ProfScope() {
    if (e1) cudaEventRecord(e1, s);
  }
};
}  // namespace

struct bvg_plan {
  bvg_handle* h = nullptr;
  int B = 0, mode = 0, dtype = 0, esize = 4;
  std::vector<int> frames;
  int max_frames = 0;
  // geometry: index 0 = latent rate (pre), 1.. = after up-conv i-1
  std::vector<int> R, maxlen, C;
  std::vector<long long> sumlen;   // valid rows over all segments, per geometry
  SegDesc* seg_dev = nullptr;   // [(nups+1)][B]
  int* prefix_dev = nullptr;    // m-tile prefix tables [g][msub in 1,2,4][q_extra][B+1]
  int* latrow_dev = nullptr;    // [B+1] prefix sum of frames (ragged latent / waveform I/O)
  char *tab_dev = nullptr, *tab_host = nullptr;   // one slice of the handle's table arena holds all three
  size_t tab_bytes = 0;
  bool tab_uploaded = false;    // the first forward copies the slice to the device on ITS stream
  std::vector<int> total_mt;    // [g][msub in 1,2,4][q_extra]
  size_t ws_bytes = 0;
  size_t off_lat = 0, off_pre = 0, off_bias = 0, off_split0 = 0;
  std::vector<size_t> off_U, off_X, off_A, off_Y, off_XS;
  // lockstep mode (16-bit modes): the nk AMP blocks of a stage advance together and share their Activation1d launches, so
  // blocks 1 .. nk-1 get their own X / A / Y buffers ([g][j-1]; block 0 uses off_X / off_A / off_Y)
  bool lockstep = false;
  std::vector<std::vector<size_t>> off_Xj, off_Aj, off_Yj;
  int bias_stride = 0;
  std::vector<int> bias_off;    // per cond layer (0 = pre, 1.. = ups)
  uint64_t uid = 0;
  int num_launches = 0;
};

namespace {

int dev_alloc(bvg_handle* h, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes ? bytes : 4));
  h->owned.push_back(*p);
  return 0;
}

int table_slice(bvg_handle* h, size_t bytes, char** dev, char** host) {
  auto& fl = h->tab_free[bytes];
  if (!fl.empty()) { *dev = fl.back().first; *host = fl.back().second; fl.pop_back(); return 0; }
  if (h->tab_chunks.empty() || h->tab_chunks.back().used + bytes > h->tab_chunks.back().cap) {
    bvg_handle::TableChunk c{nullptr, nullptr, bytes > ((size_t)1 << 20) ? bytes : ((size_t)1 << 20), 0};
    CK(cudaMalloc((void**)&c.dev, c.cap));
    if (cudaMallocHost((void**)&c.host, c.cap) != cudaSuccess) { cudaFree(c.dev); return fail("table arena: pinned allocation failed"); }
    h->tab_chunks.push_back(c);
  }
  auto& c = h->tab_chunks.back();
  *dev = c.dev + c.used; *host = c.host + c.used;
  c.used += bytes;
  return 0;
}

void add_param(bvg_handle* h, const std::string& name, float** dst, std::vector<int64_t> shape) {
  ParamSlot s;
  s.dst = dst;
  s.shape = std::move(shape);
  h->params[name] = s;
}

void add_conv_params(bvg_handle* h, const std::string& name, ConvLayer& L) {
  if (L.transposed) add_param(h, name + ".weight", &L.w_raw, {L.Cin, L.Cout, L.k});
  else add_param(h, name + ".weight", &L.w_raw, {L.Cout, L.Cin, L.k});
  add_param(h, name + ".bias", &L.bias, {L.Cout});
}

ConvArgs make_conv_args(const ConvLayer& L, const bvg_plan* p, int gin, int gout, const void* x, void* y,
                        const void* res, const float* bias, int bias_bstride, float scale, int accumulate,
                        bool umma, const ActLayer* act = nullptr) {
  ConvArgs a{};
  a.x = x; a.y = y; a.res = res;
  a.res_scale = (umma && res) ? L.fold_res : nullptr;
  a.k_packed = (umma && L.k_packed) ? 1 : 0;
  a.dtype = p->dtype;
  a.w = umma ? (const void*)(p->dtype == 2 ? L.w_umma16 : L.w_umma) : (const void*)L.w_tap;
  if (umma && L.w_umma_s && !act) {
    // small-batch variant: with 256-column n-tiles this layer would occupy fewer than half of the SMs
    const int ti1 = (gin * 3 + 0) * 2 + (L.q_extra ? 1 : 0);
    const int nt_big = (L.N + 255) / 256;
    static const int allow = [] { const char* e = getenv("BVG_CONV_SMALL"); return e ? atoi(e) : 1; }();
    if (allow && (long long)p->total_mt[ti1] * nt_big * 2 <= p->h->num_sms) {
      a.bn_small = 1;
      a.w = p->dtype == 2 ? L.w_umma16_s : L.w_umma_s;
    }
  }
  a.bias = bias; a.bias_bstride = bias_bstride;
  a.acc_img_scale = (float)p->h->nk;
  a.seg_in = p->seg_dev + (size_t)gin * p->B;
  a.seg_out = p->seg_dev + (size_t)gout * p->B;
  a.Rx = p->R[gin]; a.Ry = p->R[gout];
  a.Cin = L.Cin; a.Cout = L.Cout; a.ntaps = L.ntaps;
  for (int j = 0; j < L.ntaps; ++j) a.tap_off[j] = L.tap_off[j];
  a.u = L.u; a.p = L.p; a.q_extra = L.q_extra;
  a.B = p->B; a.max_q = p->maxlen[gin] + L.q_extra;
  a.out_scale = scale; a.accumulate = accumulate;
  a.msub = 0;
  if (act && umma) {   // try the fused Activation1d -> conv kernel
    a.act_alpha = act->alpha; a.act_inv_beta = act->inv_beta;
    a.msub = conv_umma_fused_msub(a);
    if (!a.msub) { a.act_alpha = nullptr; a.act_inv_beta = nullptr; }
  }
  if (!a.msub) a.msub = conv_umma_default_msub(a);
  const int ti = (gin * 3 + (a.msub == 4 ? 2 : a.msub - 1)) * 2 + (L.q_extra ? 1 : 0);
  a.tile_prefix = p->prefix_dev + (size_t)ti * (p->B + 1);
  a.total_mt = p->total_mt[ti];
  return a;
}

int run_conv(const ConvLayer& L, const bvg_plan* p, int gin, int gout, const void* x, void* y, const void* res,
             const float* bias, int bias_bstride, float scale, int accumulate, cudaStream_t s, int pdl_mode = 0) {
  // algorithmic work of this launch: 2 * Cin * taps * N MACs per input row (valid rows only)
  const double flops = 2.0 * L.Cin * L.ntaps * L.N * (double)p->sumlen[gin];
  const double bytes = ((double)L.Cin * p->sumlen[gin] + (double)L.Cout * p->sumlen[gout] * (res ? 2 : 1)) * p->esize;
  if (p->mode == BVG_MODE_FP32_TC) {
    // x is the split [hi | lo] fp16 view of the fp32 input; the GEMM runs over 3 Cin channels (lo, hi, hi again)
    // against S [2^-11 W_hi; W_lo; W_hi]; bias / residual / old output / store in fp32 (three tensor-core passes per MAC)
    if (!L.w_umma3) return fail("fp32 tensor-core mode: layer without a split weight image (Cin %d, N %d)", L.Cin, L.N);
    ConvArgs a = make_conv_args(L, p, gin, gout, x, y, res, bias, bias_bstride, scale, accumulate, false);
    a.w = L.w_umma3; a.dtype = 2; a.f32io = 1; a.split3_chunks = L.Cin / 8; a.Cin = 3 * L.Cin;
    a.acc_scale = L.tc32_acc_scale;
    a.msub = conv_umma_default_msub(a);
    const int ti = (gin * 3 + (a.msub == 4 ? 2 : a.msub - 1)) * 2 + (L.q_extra ? 1 : 0);
    a.tile_prefix = p->prefix_dev + (size_t)ti * (p->B + 1);
    a.total_mt = p->total_mt[ti];
    if (!conv_umma_supported(a)) return fail("fp32 tensor-core mode: layer shape not supported (Cin %d, N %d)", L.Cin, L.N);
    ProfScope ps(p->h, s, PROF_CONV_TC, flops, bytes);   // algorithmic flops (the tensor cores do 3x that)
    CK(launch_conv_umma(a, s));
    ++p->h->launch_counter;
    return 0;
  }
  if (p->mode != BVG_MODE_FP32 && L.w_umma) {   // (BVG_MODE_FP32_TC returned above)
    const float* ubias = (L.bias_umma && bias == L.bias) ? L.bias_umma : bias;   // the image's output channels are pre-scaled
    ConvArgs a = make_conv_args(L, p, gin, gout, x, y, res, ubias, bias_bstride, scale, accumulate, true);
    a.pdl_mode = pdl_mode;
    if (conv_umma_supported(a)) {
      ProfScope ps(p->h, s, PROF_CONV_TC, flops, bytes);
      CK(launch_conv_umma(a, s));
      ++p->h->launch_counter;
      return 0;
    }
  }
  ConvArgs a = make_conv_args(L, p, gin, gout, x, y, res, bias, bias_bstride, scale, accumulate, false);
  ProfScope ps(p->h, s, PROF_CONV_CC, flops, bytes);
  CK(launch_conv_simt(a, p->dtype, s));
  ++p->h->launch_counter;
  return 0;
}

int run_act(const ActLayer& A, const bvg_plan* p, int g, const void* x, void* y, cudaStream_t s, bool split_out = false);

// Activation1d followed by a convolution (one step of AMPBlock1.forward, models.py:69-72).  In the bf16
// mode the activation is fused into the conv kernel's producer stage when the layer allows it (the whole
// N fits one CTA tile); otherwise it runs as its own pass through `actbuf`.
int run_act_conv(const ActLayer& A, const ConvLayer& L, const bvg_plan* p, int g, const void* raw, void* actbuf,
                 void* y, const void* res, float scale, int accumulate, cudaStream_t s) {
  if (p->mode == BVG_MODE_BF16 && L.w_umma) {
    ConvArgs a = make_conv_args(L, p, g, g, raw, y, res, L.bias, 0, scale, accumulate, true, &A);
    if (a.act_alpha && conv_umma_supported(a)) {
      const double flops = 2.0 * L.Cin * L.ntaps * L.N * (double)p->sumlen[g];
      const double bytes = ((double)L.Cin * p->sumlen[g] + (double)L.Cout * p->sumlen[g] * (res ? 2 : 1)) * p->esize;
      ProfScope ps(p->h, s, PROF_CONV_TC, flops, bytes);
      CK(launch_conv_umma(a, s));
      ++p->h->launch_counter;
      return 0;
    }
  }
  if (run_act(A, p, g, raw, actbuf, s, p->mode == BVG_MODE_FP32_TC)) return 1;
  return run_conv(L, p, g, g, actbuf, y, res, L.bias, 0, scale, accumulate, s);
}

int run_act(const ActLayer& A, const bvg_plan* p, int g, const void* x, void* y, cudaStream_t s, bool split_out) {
  ActArgs aa{x, y, A.alpha, A.inv_beta, p->seg_dev + (size_t)g * p->B, p->R[g], p->C[g], p->B, p->maxlen[g]};
  const bool f32 = p->mode == BVG_MODE_FP32 || p->mode == BVG_MODE_FP32_TC;
  aa.prescaled = (A.prescaled && !f32) ? 1 : 0;
  aa.split_out = split_out ? 1 : 0;
  aa.fast_fp32 = p->mode == BVG_MODE_FP32_TC ? 1 : 0;
  // algorithmic bytes of a standalone Activation1d launch: read + write of every valid element
  const double bytes = 2.0 * p->C[g] * (double)p->sumlen[g] * p->esize;
  ProfScope ps(p->h, s, PROF_ACT, 0.0, bytes);
  CK(launch_act_c8(aa, p->dtype, f32, s));
  ++p->h->launch_counter;
  return 0;
}

// The same-geometry activations of the nk AMP blocks in one launch (tensor-core kernel); layers whose pre-scaled flags differ
// fall back to one launch each.
int run_act_group(const ActLayer* const* A, int n, const bvg_plan* p, int g, const void* const* x, void* const* y, cudaStream_t s) {
  bool same = n >= 2 && n <= 3;
  for (int j = 1; j < n; ++j) same = same && A[j]->prescaled == A[0]->prescaled;
  if (!same) {
    for (int j = 0; j < n; ++j)
      if (run_act(*A[j], p, g, x[j], y[j], s)) return 1;
    return 0;
  }
  ActArgs aa{x[0], y[0], A[0]->alpha, A[0]->inv_beta, p->seg_dev + (size_t)g * p->B, p->R[g], p->C[g], p->B, p->maxlen[g]};
  aa.prescaled = A[0]->prescaled ? 1 : 0;
  aa.extra_jobs = n - 1;
  for (int j = 1; j < n; ++j) {
    aa.xj[j - 1] = x[j]; aa.yj[j - 1] = y[j]; aa.alphaj[j - 1] = A[j]->alpha; aa.inv_betaj[j - 1] = A[j]->inv_beta;
  }
  const double bytes = 2.0 * n * p->C[g] * (double)p->sumlen[g] * p->esize;
  ProfScope ps(p->h, s, PROF_ACT, 0.0, bytes);
  CK(launch_act_c8(aa, p->dtype, false, s));
  ++p->h->launch_counter;
  return 0;
}

// BVG_MODE_FP32_TC: the S [2^-11 W_hi; W_lo; W_hi] images of every convolution on the path, built once per handle on the
// stream of the first forward that needs them (the 16-bit modes never pay for them: + 3x the fp16 weights = 0.68 GB).
int build_tc32_image(bvg_handle* h, ConvLayer& L, float* w3, float absmax, cudaStream_t s) {
  const size_t bytes = umma_weight_image_bytes(L.ntaps, 3 * L.Cin, L.N);
  if (!bytes) return fail("fp32 tensor-core mode: no tiling for Cin %d, N %d", L.Cin, L.N);
  if (!L.w_umma3 && dev_alloc(h, &L.w_umma3, bytes)) return 1;
  const float S = split3_weight_scale(absmax);
  L.tc32_acc_scale = 1.f / S;
  CK(launch_split3_weights(L.w_tap, w3, L.ntaps, L.Cin, L.N, S, s));
  CK(launch_repack_umma(w3, L.w_umma3, 2, L.ntaps, 3 * L.Cin, L.N, 1.f, false, s));
  return 0;
}

int ensure_tc32_images(bvg_handle* h, cudaStream_t s) {
  if (h->tc32_ready) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
    return fail("fp32 tensor-core mode: the first forward of a handle builds the split weight images (allocation + synchronisation) "
                "and cannot run inside a stream capture; run one eager forward first");
  std::vector<ConvLayer*> layers{&h->conv_pre};
  for (auto& L : h->ups) layers.push_back(&L);
  for (auto& L : h->c1) layers.push_back(&L);
  for (auto& L : h->c2) layers.push_back(&L);
  size_t mx = 0;
  for (ConvLayer* L : layers) mx = std::max(mx, (size_t)L->Cin * L->Cout * L->k);
  float *w3 = nullptr, *amax = nullptr;
  std::vector<float> amax_h(layers.size());
  CK(cudaMalloc((void**)&amax, layers.size() * sizeof(float)));
  if (cudaMalloc((void**)&w3, 3 * mx * sizeof(float)) != cudaSuccess) { cudaFree(amax); return fail("fp32 tensor-core mode: scratch allocation failed"); }
  int rc = 0;
  auto ck = [&](cudaError_t e) { if (e != cudaSuccess && !rc) rc = fail("fp32 tensor-core images: %s", cudaGetErrorString(e)); };
  ck(cudaMemsetAsync(amax, 0, layers.size() * sizeof(float), s));
  for (size_t i = 0; i < layers.size() && !rc; ++i)
    ck(launch_absmax(layers[i]->w_tap, (size_t)layers[i]->Cin * layers[i]->Cout * layers[i]->k, amax + i, s));
  ck(cudaMemcpyAsync(amax_h.data(), amax, layers.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
  ck(cudaStreamSynchronize(s));
  for (size_t i = 0; i < layers.size() && !rc; ++i) rc = build_tc32_image(h, *layers[i], w3, amax_h[i], s);
  cudaStreamSynchronize(s);
  cudaFree(w3);
  cudaFree(amax);
  if (rc) return 1;
  h->tc32_ready = true;
  return 0;
}

}  // namespace

extern "C" {

const char* bvg_last_error(void) { return g_err.c_str(); }
int bvg_set_error(const char* msg) { return fail("%s", msg ? msg : "error"); }   // for the other translation units
int bvg_version(void) { return 100; }

int bvg_device_check(void) {
  int dev = 0, n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail("no CUDA device: b200vgan has no CPU fallback (%s)", cudaGetErrorString(e));
  CK(cudaGetDevice(&dev));
  // cached per device: this runs at the top of every per-op entry point (cudaGetDeviceProperties costs milliseconds)
  static int major_of_dev[64] = {0};
  if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
  if (!major_of_dev[dev]) {
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    major_of_dev[dev] = major ? major : -1;
  }
  if (major_of_dev[dev] != 10) return fail("device %d is sm_%dx; b200vgan is built for sm_100a only", dev, major_of_dev[dev]);
  return 0;
}

int bvg_create(const bvg_config* cfg, bvg_handle** out) {
  if (!cfg || !out) return fail("bvg_create: null argument");
  if (bvg_device_check()) return 1;
  const int nups = cfg->num_upsamples, nk = cfg->num_kernels, nd = cfg->num_dilations;
  if (nups < 1 || nups > BVG_MAX_UPS || nk < 1 || nk > BVG_MAX_KERNELS || nd < 1 || nd > BVG_MAX_DILATIONS)
    return fail("bvg_create: unsupported architecture sizes");
  if (cfg->gpt_dim % 8 || cfg->upsample_initial_channel % (8 << nups))
    return fail("bvg_create: channel counts must stay multiples of 8 at every stage");
  bvg_handle* h = new bvg_handle();
  h->cfg = *cfg;
  CK(cudaGetDevice(&h->device));
  CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  h->nups = nups; h->nk = nk; h->nd = nd;
  const int C0 = cfg->upsample_initial_channel, D = cfg->speaker_embedding_dim;

  h->conv_pre.Cin = cfg->gpt_dim; h->conv_pre.Cout = C0; h->conv_pre.k = 7; h->conv_pre.setup();
  add_conv_params(h, "conv_pre", h->conv_pre);
  h->cond_pre.C = C0;
  add_param(h, "cond_layer.weight", &h->cond_pre.w, {C0, D, 1});
  add_param(h, "cond_layer.bias", &h->cond_pre.b, {C0});

  h->ups.resize(nups); h->conds.resize(nups);
  h->c1.resize((size_t)nups * nk * nd); h->c2.resize((size_t)nups * nk * nd);
  h->acts.resize((size_t)nups * nk * 2 * nd);
  h->hop = 1;
  int ch = C0;
  for (int i = 0; i < nups; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u < 1 || k % u || (k - u) % 2 || k / u > BVG_MAX_TAPS || (k - u) / 2 > 8 * u) {
      delete h;
      return fail("bvg_create: upsample (k=%d,u=%d) unsupported (need k %% u == 0, k - u even, k/u <= %d taps)", k, u, BVG_MAX_TAPS);
    }
    h->hop *= u;
    ConvLayer& U = h->ups[i];
    U.transposed = true; U.Cin = ch; U.Cout = ch / 2; U.k = k; U.u = u; U.setup();
    ch /= 2;
    char nm[128];
    snprintf(nm, sizeof nm, "ups.%d.0", i);
    add_conv_params(h, nm, U);
    if (cfg->cond_in_each_up_layer) {
      h->conds[i].C = ch;
      snprintf(nm, sizeof nm, "conds.%d.weight", i);
      add_param(h, nm, &h->conds[i].w, {ch, D, 1});
      snprintf(nm, sizeof nm, "conds.%d.bias", i);
      add_param(h, nm, &h->conds[i].b, {ch});
    }
    for (int j = 0; j < nk; ++j) {
      const int rb = i * nk + j, ks = cfg->resblock_kernel_sizes[j];
      if (ks > BVG_MAX_TAPS || !(ks & 1)) { delete h; return fail("bvg_create: resblock kernel %d unsupported", ks); }
      for (int m = 0; m < nd; ++m) {
        ConvLayer& A = h->c1[(size_t)rb * nd + m];
        A.Cin = A.Cout = ch; A.k = ks; A.d = cfg->resblock_dilation_sizes[j][m]; A.setup();
        ConvLayer& Bc = h->c2[(size_t)rb * nd + m];
        Bc.Cin = Bc.Cout = ch; Bc.k = ks; Bc.d = 1; Bc.setup();
        if ((ks - 1) / 2 * A.d > BVG_GUARD - 6) { delete h; return fail("bvg_create: conv halo exceeds guard rows"); }
        snprintf(nm, sizeof nm, "resblocks.%d.convs1.%d", rb, m);
        add_conv_params(h, nm, A);
        snprintf(nm, sizeof nm, "resblocks.%d.convs2.%d", rb, m);
        add_conv_params(h, nm, Bc);
      }
      for (int a = 0; a < 2 * nd; ++a) {
        ActLayer& AL = h->acts[(size_t)rb * 2 * nd + a];
        AL.C = ch;
        snprintf(nm, sizeof nm, "resblocks.%d.activations.%d.act.alpha", rb, a);
        add_param(h, nm, &AL.la, {ch});
        snprintf(nm, sizeof nm, "resblocks.%d.activations.%d.act.beta", rb, a);
        add_param(h, nm, &AL.lb, {ch});
      }
    }
  }
  h->act_post.C = ch;
  add_param(h, "activation_post.act.alpha", &h->act_post.la, {ch});
  add_param(h, "activation_post.act.beta", &h->act_post.lb, {ch});
  h->conv_post.Cin = ch; h->conv_post.Cout = 1; h->conv_post.k = 7; h->conv_post.setup();
  add_conv_params(h, "conv_post", h->conv_post);
  if (ch > 64) { delete h; return fail("bvg_create: conv_post supports at most 64 input channels"); }
  ecapa_register(h->ecapa, cfg->num_mels > 0 ? cfg->num_mels : 100, D,
                 [h](const std::string& name, float** dst, std::vector<long long> shape) {
                   ParamSlot sl;
                   sl.dst = dst;
                   sl.shape.assign(shape.begin(), shape.end());
                   h->ecapa_params["speaker_encoder." + name] = sl;
                 });
  *out = h;
  return 0;
}

void bvg_destroy(bvg_handle* h) {
  if (!h) return;
  for (void* p : h->owned) cudaFree(p);
  for (auto& c : h->tab_chunks) { cudaFree(c.dev); cudaFreeHost(c.host); }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  delete h;
}

int bvg_set_weight(bvg_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                   int32_t is_device, void* stream) {
  if (!h || !name || !data) return fail("bvg_set_weight: null argument");
  auto it = h->params.find(name);
  if (it == h->params.end()) {
    it = h->ecapa_params.find(name);
    if (it == h->ecapa_params.end()) return fail("bvg_set_weight: unknown parameter '%s'", name);
  }
  ParamSlot& s = it->second;
  size_t n = 1;
  bool ok = (size_t)ndim == s.shape.size();
  for (int i = 0; ok && i < ndim; ++i) ok = shape[i] == s.shape[i];
  if (!ok) {
    std::string want, got;
    for (auto v : s.shape) want += std::to_string(v) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    return fail("bvg_set_weight: '%s' expects shape [%s] got [%s]", name, want.c_str(), got.c_str());
  }
  for (auto v : s.shape) n *= (size_t)v;
  if (!*s.dst) {
    void* p = nullptr;
    if (dev_alloc(h, &p, n * sizeof(float))) return 1;
    *s.dst = (float*)p;
  }
  CK(cudaMemcpyAsync(*s.dst, data, n * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                     (cudaStream_t)stream));
  if (!is_device) CK(cudaStreamSynchronize((cudaStream_t)stream));
  s.set = true;
  h->finalized = false;
  h->tc32_ready = false;   // the fp32 tensor-core images are rebuilt from the new weights by the next forward that needs them
  return 0;
}

static int finalize_conv(bvg_handle* h, ConvLayer& L, cudaStream_t s, bool want_umma) {
  size_t n = (size_t)L.Cin * L.Cout * L.k;
  if (!L.w_tap) {
    void* p = nullptr;
    if (dev_alloc(h, &p, n * sizeof(float))) return 1;
    L.w_tap = (float*)p;
  }
  if (L.transposed) CK(launch_repack_convt(L.w_raw, L.w_tap, L.Cin, L.Cout, L.k, L.u, s));
  else CK(launch_repack_conv(L.w_raw, L.w_tap, L.Cout, L.Cin, L.k, s));
  if (want_umma) {
    L.k_packed = !L.transposed && umma_k_packed_default(L.Cin, L.N);
    size_t bytes = umma_weight_image_bytes(L.ntaps, L.Cin, L.N, false, L.k_packed);
    if (bytes) {
      if (!L.w_umma) {
        if (dev_alloc(h, &L.w_umma, bytes)) return 1;
      }
      if (!L.w_umma16) {
        if (dev_alloc(h, &L.w_umma16, bytes)) return 1;
      }
      CK(launch_repack_umma(L.w_tap, L.w_umma, 1, L.ntaps, L.Cin, L.N, (float)h->nk, false, s, L.fold_out, L.fold_in, L.fold_res, L.k_packed));
      CK(launch_repack_umma(L.w_tap, L.w_umma16, 2, L.ntaps, L.Cin, L.N, (float)h->nk, false, s, L.fold_out, L.fold_in, L.fold_res, L.k_packed));
      if (L.fold_out) {
        if (!L.bias_umma) {
          void* pb = nullptr;
          if (dev_alloc(h, &pb, (size_t)L.Cout * sizeof(float))) return 1;
          L.bias_umma = (float*)pb;
        }
        CK(launch_scale_vec(L.bias, L.fold_out, L.bias_umma, L.Cout, s));
      }
      const size_t sbytes = L.N >= 256 ? umma_weight_image_bytes(L.ntaps, L.Cin, L.N, true) : 0;
      if (sbytes) {
        if (!L.w_umma_s && (dev_alloc(h, &L.w_umma_s, sbytes) || dev_alloc(h, &L.w_umma16_s, sbytes))) return 1;
        CK(launch_repack_umma(L.w_tap, L.w_umma_s, 1, L.ntaps, L.Cin, L.N, (float)h->nk, true, s, L.fold_out, L.fold_in, L.fold_res));
        CK(launch_repack_umma(L.w_tap, L.w_umma16_s, 2, L.ntaps, L.Cin, L.N, (float)h->nk, true, s, L.fold_out, L.fold_in, L.fold_res));
      }
    }
  }
  return 0;
}

static int finalize_act(bvg_handle* h, ActLayer& A, cudaStream_t s) {
  if (!A.alpha) {
    void* p = nullptr;
    if (dev_alloc(h, &p, 2 * (size_t)A.C * sizeof(float))) return 1;
    A.alpha = (float*)p;
    A.inv_beta = A.alpha + A.C;
  }
  CK(launch_snake_params(A.la, A.lb, A.alpha, A.inv_beta, A.C, s));
  return 0;
}

// 16-bit modes: pre-scaled activations.  An AMPBlock1 (models.py:65-74) runs, for m = 0 .. nd-1,
//     x_{m+1} = c2_m(act2_m(c1_m(act1_m(x_m)))) + x_m.
// The tensor-core activation kernel has a variant (bvg_act3.cu, PRE) whose input already carries the factor 2 alpha per
// channel -- the up-FIR then delivers the cosine's argument directly, one multiply per activated sample less -- and whose
// output keeps that factor.  All the factors are folded into the neighbouring convolutions here, in fp32, before the UMMA
// images are rounded once:
//   * act2_m: reads only c1_m's output, feeds only c2_m  ->  2 alpha into c1_m's output channels (weights and bias),
//     1 / (2 alpha) into c2_m's input channels;
//   * act1_m, m >= 1: reads the residual stream x_m, which c2_{m-1} writes and c2_m adds back  ->  x_m is STORED as
//     s_m x_m (s_m = 2 alpha of act1_m): s_m into c2_{m-1}'s output channels and bias, 1 / s_m into c1_m's input channels,
//     and a per-channel factor on each c2's residual input (s_{m+1} / s_m, with s = 1 where nothing is folded: x_0 is the
//     stage input shared by the nk resblocks, and the last c2 accumulates into the unscaled stage output) -- the diagonal of
//     the residual identity image.  Done for the narrow layers (C <= 96) only: in the wide ones the residual is added in the
//     epilogue and the extra per-channel multiply there costs more than the activation saves.
// A layer whose alpha leaves [1/64, 64] keeps the plain kernel (fp16 range).  The fp32 parity mode, the per-op entry points
// and the experimental fused kernel use the unfolded parameters.
static int fold_activation_scales(bvg_handle* h, cudaStream_t s) {
  static const int enabled = [] {
    const char* e = getenv("BVG_ACT_FOLD");   // 0 = off, 1 = act2 only, 2 (default) = act2 and the act1 of m >= 1
    const char* f = getenv("BVG_FUSE_ACT");   // the experimental fused kernel and the register-streamed kernel
    const char* m = getenv("BVG_ACT_MMA");    // (BVG_ACT_MMA=0) have no pre-scaled variant
    if ((f && atoi(f)) || (m && !atoi(m))) return 0;
    return e ? atoi(e) : 2;
  }();
  const int nd = h->nd;
  auto alpha_host = [&](const ActLayer& A, std::vector<float>& out) -> int {
    out.resize(A.C);
    CK(cudaMemcpyAsync(out.data(), A.alpha, (size_t)A.C * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
  };
  auto in_range = [](const std::vector<float>& a) {
    bool ok = true;
    for (float v : a) ok = ok && v >= 1.f / 64.f && v <= 64.f;
    return ok;
  };
  auto upload = [&](float** dst, size_t n, const std::vector<float>& v) -> int {
    if (!*dst) {
      void* pv = nullptr;
      if (dev_alloc(h, &pv, n * sizeof(float))) return 1;
      *dst = (float*)pv;
    }
    CK(cudaMemcpyAsync(*dst, v.data(), n * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return 0;
  };
  for (size_t rb = 0; rb < (size_t)h->nups * h->nk; ++rb) {
    const int C = h->c1[rb * nd].Cin;
    // host copies of 2 alpha for the 2 nd activations of this resblock (empty = not pre-scaled)
    std::vector<std::vector<float>> sc(2 * nd);
    for (int a = 0; a < 2 * nd; ++a) {
      ActLayer& A = h->acts[rb * 2 * nd + a];
      A.prescaled = false;
      const bool is_act2 = a & 1;
      // act1 of m = 0 reads the shared stage input.  act1 of m >= 1: only where the residual goes through the identity MMA
      // (C <= 96): in the wide layers the per-channel residual factor is an epilogue multiply that costs more than the
      // activation saves (measured: +15 us per c2 launch at C = 192); BVG_ACT_FOLD=3 folds them anyway.
      const bool want = enabled >= 1 && (is_act2 || (a >= 2 && (enabled >= 3 || (enabled >= 2 && C <= 96))));
      if (!want) continue;
      std::vector<float> al;
      if (alpha_host(A, al)) return 1;
      if (!in_range(al)) continue;
      for (float& v : al) v *= 2.f;
      sc[a] = al;
      A.prescaled = true;
    }
    for (int m = 0; m < nd; ++m) {
      ConvLayer& c1 = h->c1[rb * nd + m];
      ConvLayer& c2 = h->c2[rb * nd + m];
      const std::vector<float>& s1 = sc[2 * m];                                  // act1_m: scale of the stored x_m
      const std::vector<float>& s2 = sc[2 * m + 1];                              // act2_m
      static const std::vector<float> none;
      const std::vector<float>& snext = m + 1 < nd ? sc[2 * (m + 1)] : none;     // scale of the stored x_{m+1}
      // c1_m: input channels / s1, output channels * s2          c2_m: input / s2, output * snext, residual * snext / s1
      std::vector<float> buf(2 * (size_t)C);
      c1.fold_in = c1.fold_out = c2.fold_in = c2.fold_out = c2.fold_res = nullptr; c1.fold_res = nullptr;
      if (!s1.empty() || !s2.empty()) {
        for (int c = 0; c < C; ++c) { buf[c] = s1.empty() ? 1.f : 1.f / s1[c]; buf[C + c] = s2.empty() ? 1.f : s2[c]; }
        if (upload(&c1.fold_buf, buf.size(), buf)) return 1;
        if (!s1.empty()) c1.fold_in = c1.fold_buf;
        if (!s2.empty()) c1.fold_out = c1.fold_buf + C;
      }
      if (!s2.empty() || !snext.empty() || !s1.empty()) {
        std::vector<float> b3(3 * (size_t)C);
        for (int c = 0; c < C; ++c) {
          b3[c] = s2.empty() ? 1.f : 1.f / s2[c];
          b3[C + c] = snext.empty() ? 1.f : snext[c];
          b3[2 * C + c] = (snext.empty() ? 1.f : snext[c]) / (s1.empty() ? 1.f : s1[c]);
        }
        if (upload(&c2.fold_buf, b3.size(), b3)) return 1;
        if (!s2.empty()) c2.fold_in = c2.fold_buf;
        if (!snext.empty()) c2.fold_out = c2.fold_buf + C;
        if (!snext.empty() || !s1.empty()) c2.fold_res = c2.fold_buf + 2 * C;
      }
    }
  }
  return 0;
}

int bvg_finalize(bvg_handle* h, void* stream) {
  if (!h) return fail("bvg_finalize: null handle");
  cudaStream_t s = (cudaStream_t)stream;
  for (auto& kv : h->params)
    if (!kv.second.set) return fail("bvg_finalize: parameter '%s' was never set", kv.first.c_str());
  for (auto& A : h->acts) if (finalize_act(h, A, s)) return 1;
  if (finalize_act(h, h->act_post, s)) return 1;
  if (fold_activation_scales(h, s)) return 1;
  if (finalize_conv(h, h->conv_pre, s, true)) return 1;
  for (auto& L : h->ups) if (finalize_conv(h, L, s, true)) return 1;
  for (auto& L : h->c1) if (finalize_conv(h, L, s, true)) return 1;
  for (auto& L : h->c2) if (finalize_conv(h, L, s, true)) return 1;
  // speaker encoder: all-or-nothing
  size_t nset = 0;
  for (auto& kv : h->ecapa_params) nset += kv.second.set ? 1 : 0;
  h->ecapa_ready = false;
  if (nset) {
    for (auto& kv : h->ecapa_params)
      if (!kv.second.set) return fail("bvg_finalize: speaker-encoder parameter '%s' was never set", kv.first.c_str());
    std::vector<EcapaTdnn*> tdnns;
    ecapa_collect_bn(h->ecapa, tdnns);
    for (EcapaTdnn* t : tdnns) {
      if (!t->scale) {
        void* p = nullptr;
        if (dev_alloc(h, &p, 2 * (size_t)t->cout * sizeof(float))) return 1;
        t->scale = (float*)p; t->shift = t->scale + t->cout;
      }
      CK(ecapa_fold_bn(t->bn_w, t->bn_b, t->bn_m, t->bn_v, t->scale, t->shift, t->cout, s));
    }
    const int c6 = 6 * h->ecapa.C;
    if (!h->ecapa.abn_scale) {
      void* p = nullptr;
      if (dev_alloc(h, &p, 2 * (size_t)c6 * sizeof(float))) return 1;
      h->ecapa.abn_scale = (float*)p; h->ecapa.abn_shift = h->ecapa.abn_scale + c6;
    }
    CK(ecapa_fold_bn(h->ecapa.abn_w, h->ecapa.abn_b, h->ecapa.abn_m, h->ecapa.abn_v, h->ecapa.abn_scale, h->ecapa.abn_shift, c6, s));
    h->ecapa_ready = true;
  }
  CK(cudaStreamSynchronize(s));
  h->finalized = true;
  return 0;
}

size_t bvg_ecapa_workspace_bytes(const bvg_handle* h, int32_t B, int32_t Tm) {
  if (!h || B < 1 || Tm < 1) return 0;
  return ecapa_workspace_bytes(h->ecapa, B, Tm);
}

int bvg_speaker_embedding(bvg_handle* h, const float* mel, int32_t B, int32_t Tm, const float* rel_lens, float* emb,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !mel || !emb || !workspace) return fail("bvg_speaker_embedding: null argument");
  if (bvg_device_check()) return 1;
  if (!h->finalized || !h->ecapa_ready)
    return fail("bvg_speaker_embedding: the speaker_encoder.* parameters were not uploaded (bvg_set_weight) before bvg_finalize");
  if (B < 1 || Tm < 5) return fail("bvg_speaker_embedding: need B >= 1 and at least 5 mel frames (reflect padding), got B=%d Tm=%d", B, Tm);
  if (workspace_bytes < ecapa_workspace_bytes(h->ecapa, B, Tm)) return fail("bvg_speaker_embedding: workspace too small");
  int launches = 0;
  CK(ecapa_forward(h->ecapa, mel, B, Tm, rel_lens, emb, workspace, (cudaStream_t)stream, &launches));
  return 0;
}

int bvg_plan_create(bvg_handle* h, int32_t B, const int32_t* frames, int32_t mode, bvg_plan** out) {
  if (!h || !frames || !out || B < 1) return fail("bvg_plan_create: bad argument");
  if (mode != BVG_MODE_FP32 && mode != BVG_MODE_BF16 && mode != BVG_MODE_F16 && mode != BVG_MODE_FP32_TC)
    return fail("bvg_plan_create: unknown mode %d", mode);
  static uint64_t next_uid = 1;
  bvg_plan* p = new bvg_plan();
  p->uid = next_uid++;
  p->h = h; p->B = B; p->mode = mode;
  p->dtype = (mode == BVG_MODE_FP32 || mode == BVG_MODE_FP32_TC) ? 0 : (mode == BVG_MODE_BF16 ? 1 : 2);   // storage type of the packed tensors
  p->esize = p->dtype == 0 ? 4 : 2;
  p->frames.assign(frames, frames + B);
  const int ng = h->nups + 1;
  std::vector<SegDesc> seg((size_t)ng * B);
  p->R.resize(ng); p->maxlen.resize(ng); p->C.resize(ng); p->sumlen.assign(ng, 0);
  int mult = 1;
  for (int g = 0; g < ng; ++g) {
    if (g > 0) mult *= h->cfg.upsample_rates[g - 1];
    p->C[g] = g == 0 ? h->cfg.upsample_initial_channel : h->cfg.upsample_initial_channel >> g;
    long long row = BVG_GUARD;
    int mx = 0;
    for (int b = 0; b < B; ++b) {
      if (frames[b] < 1) { delete p; return fail("bvg_plan_create: segment %d has %d frames", b, frames[b]); }
      int len = frames[b] * mult;
      seg[(size_t)g * B + b] = SegDesc{(int)row, len};
      row += len + BVG_GUARD;
      mx = len > mx ? len : mx;
      p->sumlen[g] += len;
    }
    row += BVG_TAIL_SLACK;
    if (row > 0x3fffffff) { delete p; return fail("bvg_plan_create: batch too long"); }
    p->R[g] = (int)row;
    p->maxlen[g] = mx;
  }
  p->max_frames = p->maxlen[0];
  {
    // one table slice: [SegDesc ng*B][tile prefix ng*6*(B+1)][latent row prefix B+1]
    const size_t seg_bytes = align_up(seg.size() * sizeof(SegDesc), 16);
    const size_t pref_n = (size_t)ng * 6 * (B + 1);
    const size_t pref_bytes = align_up(pref_n * sizeof(int), 16);
    p->tab_bytes = align_up(seg_bytes + pref_bytes + (size_t)(B + 1) * sizeof(int), 1024);
    if (table_slice(h, p->tab_bytes, &p->tab_dev, &p->tab_host)) { delete p; return 1; }
    memcpy(p->tab_host, seg.data(), seg.size() * sizeof(SegDesc));
    int* pref = reinterpret_cast<int*>(p->tab_host + seg_bytes);
    p->total_mt.assign((size_t)ng * 6, 0);
    for (int g = 0; g < ng; ++g)
      for (int mi = 0; mi < 3; ++mi)
        for (int qe = 0; qe < 2; ++qe) {
          const int ms = 1 << mi;   // 128-row sub-tiles per tile: 1, 2, 4
          const int ti = (g * 3 + mi) * 2 + qe;
          // slot 1: extra q rows of the up-convolution that reads this geometry (the only transposed layer that does)
          const int qx = (qe && g < h->nups) ? h->ups[g].q_extra : 0;
          int* pf = pref + (size_t)ti * (B + 1);
          pf[0] = 0;
          for (int b = 0; b < B; ++b) pf[b + 1] = pf[b] + (seg[(size_t)g * B + b].len + qx + 128 * ms - 1) / (128 * ms);
          p->total_mt[ti] = pf[B];
        }
    int* lr = reinterpret_cast<int*>(p->tab_host + seg_bytes + pref_bytes);
    lr[0] = 0;
    for (int b = 0; b < B; ++b) lr[b + 1] = lr[b] + frames[b];
    p->seg_dev = reinterpret_cast<SegDesc*>(p->tab_dev);
    p->prefix_dev = reinterpret_cast<int*>(p->tab_dev + seg_bytes);
    p->latrow_dev = reinterpret_cast<int*>(p->tab_dev + seg_bytes + pref_bytes);
    p->tab_uploaded = false;
    ++h->plans_created;
  }
  // workspace carve-up
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t es = p->esize;
  p->off_lat = take((size_t)h->cfg.gpt_dim * p->R[0] * es);
  p->off_pre = take((size_t)p->C[0] * p->R[0] * es);
  // fp32 tensor-core mode: the split [hi | lo] copy of conv_pre's output, input of the first up-convolution (the later
  // stages split their input into the previous stage's activation buffer, which is free by then)
  p->off_split0 = mode == BVG_MODE_FP32_TC ? take((size_t)p->C[0] * p->R[0] * es) : 0;
  p->off_U.resize(ng); p->off_X.resize(ng); p->off_A.resize(ng); p->off_Y.resize(ng); p->off_XS.resize(ng);
  for (int g = 1; g < ng; ++g) {
    size_t bytes = (size_t)p->C[g] * p->R[g] * es;
    p->off_U[g] = take(bytes); p->off_X[g] = take(bytes); p->off_A[g] = take(bytes);
    p->off_Y[g] = take(bytes); p->off_XS[g] = take(bytes);
  }
  {
    static const int group = [] {
      const char* e = getenv("BVG_ACT_GROUP");
      const char* f = getenv("BVG_FUSE_ACT");   // the fused kernel and the register-streamed kernel keep one block at a time
      const char* m = getenv("BVG_ACT_MMA");
      if ((f && atoi(f)) || (m && !atoi(m))) return 0;
      return e ? atoi(e) : 1;
    }();
    p->lockstep = group && (mode == BVG_MODE_BF16 || mode == BVG_MODE_F16) && h->nk >= 2 && h->nk <= 3;
    p->off_Xj.resize(ng); p->off_Aj.resize(ng); p->off_Yj.resize(ng);
    if (p->lockstep)
      for (int g = 1; g < ng; ++g) {
        const size_t bytes = (size_t)p->C[g] * p->R[g] * es;
        for (int j = 1; j < h->nk; ++j) {
          p->off_Xj[g].push_back(take(bytes)); p->off_Aj[g].push_back(take(bytes)); p->off_Yj[g].push_back(take(bytes));
        }
      }
  }
  p->bias_off.resize(ng);
  int bo = 0;
  for (int g = 0; g < ng; ++g) { p->bias_off[g] = bo; bo += p->C[g]; }
  p->bias_stride = bo;
  p->off_bias = take((size_t)B * bo * sizeof(float));
  p->ws_bytes = off;
  p->num_launches = 1 + 1 + ng + 1 + h->nups * (1 + h->nk * h->nd * 4) + 2;
  *out = p;
  return 0;
}

void bvg_plan_destroy(bvg_plan* p) {
  if (!p) return;
  if (p->tab_dev) p->h->tab_free[p->tab_bytes].push_back({p->tab_dev, p->tab_host});   // recycle the table slice
  delete p;
}
int64_t bvg_plans_created(const bvg_handle* h) { return h ? h->plans_created : 0; }
int64_t bvg_plan_total_frames(const bvg_plan* p) { return p ? (int64_t)p->sumlen[0] : 0; }
size_t bvg_plan_workspace_bytes(const bvg_plan* p) { return p ? p->ws_bytes : 0; }
int32_t bvg_plan_max_frames(const bvg_plan* p) { return p ? p->max_frames : 0; }
int32_t bvg_plan_num_launches(const bvg_plan* p) { return p ? p->num_launches : 0; }

static int forward_impl(bvg_handle* h, bvg_plan* p, const void* latent, int32_t latent_dtype, const float* spk_emb,
                        int32_t spk_batch, float* wav, int16_t* pcm, void* workspace, size_t workspace_bytes, void* stream,
                        bool ragged = false);

int bvg_forward_ragged(bvg_handle* h, bvg_plan* p, const void* latent_rows, int32_t latent_dtype, const float* spk_emb,
                       int32_t spk_batch, float* wav_rows_or_null, int16_t* pcm_rows_or_null, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!wav_rows_or_null && !pcm_rows_or_null) return fail("bvg_forward_ragged: no output buffer");
  return forward_impl(h, p, latent_rows, latent_dtype, spk_emb, spk_batch, wav_rows_or_null, pcm_rows_or_null, workspace,
                      workspace_bytes, stream, true);
}

int bvg_forward(bvg_handle* h, bvg_plan* p, const void* latent, int32_t latent_dtype, const float* spk_emb,
                int32_t spk_batch, float* wav, void* workspace, size_t workspace_bytes, void* stream) {
  if (!wav) return fail("bvg_forward: null argument");
  return forward_impl(h, p, latent, latent_dtype, spk_emb, spk_batch, wav, nullptr, workspace, workspace_bytes, stream);
}

int bvg_forward_pcm16(bvg_handle* h, bvg_plan* p, const void* latent, int32_t latent_dtype, const float* spk_emb,
                      int32_t spk_batch, int16_t* pcm, float* wav_or_null, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!pcm) return fail("bvg_forward_pcm16: null argument");
  return forward_impl(h, p, latent, latent_dtype, spk_emb, spk_batch, wav_or_null, pcm, workspace, workspace_bytes, stream);
}

static int forward_impl(bvg_handle* h, bvg_plan* p, const void* latent, int32_t latent_dtype, const float* spk_emb,
                        int32_t spk_batch, float* wav, int16_t* pcm, void* workspace, size_t workspace_bytes, void* stream,
                        bool ragged) {
  if (!h || !p || !latent || !spk_emb || !workspace) return fail("bvg_forward: null argument");
  if (!h->finalized) return fail("bvg_forward: call bvg_finalize first");
  if (p->h != h) return fail("bvg_forward: plan belongs to another handle");
  if (workspace_bytes < p->ws_bytes) return fail("bvg_forward: workspace too small (%zu < %zu)", workspace_bytes, p->ws_bytes);
  if (spk_batch != 1 && spk_batch != p->B) return fail("bvg_forward: spk_batch must be 1 or B");
  if (latent_dtype < 0 || latent_dtype > 2) return fail("bvg_forward: bad latent dtype");
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int B = p->B, ng = h->nups + 1, dt = p->dtype;
  const bool tc32 = p->mode == BVG_MODE_FP32_TC;
  if (tc32 && ensure_tc32_images(h, s)) return 1;
  // plan tables: pinned host mirror -> device slice, once, ordered on this stream (no allocation, no host sync)
  if (!p->tab_uploaded) {
    CK(cudaMemcpyAsync(p->tab_dev, p->tab_host, p->tab_bytes, cudaMemcpyHostToDevice, s));
    p->tab_uploaded = true;
  }
  // Guard rows (between / around segments) must read as zero.  Kernels never write them, but the workspace may have
  // been laid out for another plan (or be a recycled allocation) since: one launch re-clears the guard rows of all
  // 2 + 5 * nups buffers on every forward (a few hundred KB of stores).
  {
    GuardJobs jobs{};
    auto add = [&](size_t off, int g, int C, bool split = false) {
      // (a split buffer is a bf16 tensor of 2 C channels in the same bytes)
      if (split) jobs.job[jobs.n++] = GuardJob{ws + off, p->seg_dev + (size_t)g * B, 2 * (C >> 3), p->R[g], 1};
      else jobs.job[jobs.n++] = GuardJob{ws + off, p->seg_dev + (size_t)g * B, C >> 3, p->R[g], p->esize == 4 ? 2 : 1};
    };
    add(p->off_lat, 0, h->cfg.gpt_dim, tc32);
    add(p->off_pre, 0, p->C[0]);
    if (tc32) add(p->off_split0, 0, p->C[0], true);
    for (int g = 1; g < ng; ++g) {
      for (size_t o : {p->off_U[g], p->off_X[g], p->off_Y[g], p->off_XS[g]}) add(o, g, p->C[g]);
      add(p->off_A[g], g, p->C[g], tc32);
      for (size_t o : p->off_Xj[g]) add(o, g, p->C[g]);
      for (size_t o : p->off_Aj[g]) add(o, g, p->C[g]);
      for (size_t o : p->off_Yj[g]) add(o, g, p->C[g]);
    }
    ProfScope ps(h, s, PROF_OTHER, 0.0, 0.0);
    CK(launch_zero_guards_all(jobs, B, s));
  }
  const SegDesc* seg0 = p->seg_dev;
  h->launch_counter = 1 + 1 + ng + 1;   // guards + pack + cond biases + conv_post (the helpers below count their own)
  float* biasb = (float*)(ws + p->off_bias);
  const int D = h->cfg.speaker_embedding_dim;
  {
  ProfScope ps(h, s, PROF_OTHER, 0.0, (double)h->cfg.gpt_dim * p->sumlen[0] * (4.0 + p->esize));
  CK(launch_pack_latent(latent, latent_dtype, ws + p->off_lat, tc32 ? 3 : dt, seg0, ragged ? p->latrow_dev : nullptr, B, p->max_frames,
                        h->cfg.gpt_dim, p->R[0], s));
  // speaker conditioning folded into per-segment biases
  CK(launch_cond_bias(h->conv_pre.bias, h->cond_pre.w, h->cond_pre.b, spk_emb, biasb + p->bias_off[0], p->C[0], D, B,
                      spk_batch, p->bias_stride, s));
  for (int g = 1; g < ng; ++g) {
    const CondLayer& c = h->conds[g - 1];
    CK(launch_cond_bias(h->ups[g - 1].bias, c.w, c.b, spk_emb, biasb + p->bias_off[g], p->C[g], D, B, spk_batch,
                        p->bias_stride, s));
  }
  }
  // conv_pre (models.py:224-226)
  if (run_conv(h->conv_pre, p, 0, 0, ws + p->off_lat, ws + p->off_pre, nullptr, biasb + p->bias_off[0], p->bias_stride,
               1.f, 0, s)) return 1;
  const void* stage_in = ws + p->off_pre;
  for (int g = 1; g < ng; ++g) {
    const int i = g - 1;
    char *U = ws + p->off_U[g], *X = ws + p->off_X[g], *A = ws + p->off_A[g], *Y = ws + p->off_Y[g],
         *XS = ws + p->off_XS[g];
    // up-convolution + conditioning (models.py:230-234)
    if (tc32) {   // split the fp32 stage input for the tensor cores
      char* sp = g == 1 ? ws + p->off_split0 : ws + p->off_A[g - 1];
      ProfScope ps(h, s, PROF_OTHER, 0.0, 2.0 * p->C[g - 1] * (double)p->sumlen[g - 1] * 4.0);
      CK(launch_split_c8((const float*)stage_in, sp, p->seg_dev + (size_t)(g - 1) * B, B, p->C[g - 1], p->R[g - 1], p->maxlen[g - 1], s));
      ++h->launch_counter;
      stage_in = sp;
    }
    if (run_conv(h->ups[i], p, g - 1, g, stage_in, U, nullptr, biasb + p->bias_off[g], p->bias_stride, 1.f, 0, s))
      return 1;
    if (p->lockstep) {
      // The nk AMP blocks (models.py:237-243) are independent until their outputs are averaged: step them together, one
      // Activation1d launch per step for all of them (the kernel's fixed cost -- a warp's latency through two tiles plus
      // the tail of the last wave, ~16 us -- is paid once instead of nk times), their convolutions back to back.
      const int nk = h->nk, nd = h->nd;
      char *Xj[3] = {X, nullptr, nullptr}, *Aj[3] = {A, nullptr, nullptr}, *Yj[3] = {Y, nullptr, nullptr};
      for (int j = 1; j < nk; ++j) {
        Xj[j] = ws + p->off_Xj[g][j - 1]; Aj[j] = ws + p->off_Aj[g][j - 1]; Yj[j] = ws + p->off_Yj[g][j - 1];
      }
      for (int m = 0; m < nd; ++m) {
        const bool last = m == nd - 1;
        const ActLayer* a1[3]; const ActLayer* a2[3]; const void* in[3]; void* outA[3]; const void* inY[3];
        for (int j = 0; j < nk; ++j) {
          const int rb = i * nk + j;
          a1[j] = &h->acts[(size_t)rb * 2 * nd + 2 * m];
          a2[j] = &h->acts[(size_t)rb * 2 * nd + 2 * m + 1];
          in[j] = m == 0 ? U : Xj[j];
          outA[j] = Aj[j];
          inY[j] = Yj[j];
        }
        if (run_act_group(a1, nk, p, g, in, outA, s)) return 1;
        // the nk convolutions of a step do not depend on each other: the later ones start under their predecessor's tail
        // (ConvArgs::pdl_mode); the last step's second convolutions accumulate into XS one after the other
        static const int indep = [] { const char* e = getenv("BVG_PDL_INDEP"); return e ? atoi(e) : 1; }();
        for (int j = 0; j < nk; ++j) {
          const ConvLayer& c1 = h->c1[(size_t)(i * nk + j) * nd + m];
          if (run_conv(c1, p, g, g, Aj[j], Yj[j], nullptr, c1.bias, 0, 1.f, 0, s, indep ? (j == 0 ? 1 : 2) : 0)) return 1;
        }
        if (run_act_group(a2, nk, p, g, inY, outA, s)) return 1;
        for (int j = 0; j < nk; ++j) {
          const ConvLayer& c2 = h->c2[(size_t)(i * nk + j) * nd + m];
          if (run_conv(c2, p, g, g, Aj[j], last ? XS : Xj[j], in[j], c2.bias, 0, last ? 1.f / nk : 1.f, last && j > 0, s,
                       (indep && !last) ? (j == 0 ? 1 : 2) : 0)) return 1;
        }
      }
    } else
    for (int j = 0; j < h->nk; ++j) {          // AMP blocks (models.py:237-243)
      const int rb = i * h->nk + j;
      for (int m = 0; m < h->nd; ++m) {        // AMPBlock1.forward (models.py:65-74)
        const char* xin = m == 0 ? U : X;
        const ActLayer& a1 = h->acts[(size_t)rb * 2 * h->nd + 2 * m];
        const ActLayer& a2 = h->acts[(size_t)rb * 2 * h->nd + 2 * m + 1];
        const ConvLayer& c1 = h->c1[(size_t)rb * h->nd + m];
        if (run_act_conv(a1, c1, p, g, xin, A, Y, nullptr, 1.f, 0, s)) return 1;
        const ConvLayer& c2 = h->c2[(size_t)rb * h->nd + m];
        const bool last = m == h->nd - 1;
        if (run_act_conv(a2, c2, p, g, Y, A, last ? XS : X, xin, last ? 1.f / h->nk : 1.f, last && j > 0, s))
          return 1;
      }
    }
    stage_in = XS;
  }
  // activation_post -> conv_post -> tanh (models.py:246-248)
  {
    const int g = ng - 1;
    const SegDesc* seg = p->seg_dev + (size_t)g * B;
    // (fp32 tensor-core mode: A[g] holds split data and split-view guard rows; the plain fp32 result goes to Y[g])
    char* post = tc32 ? ws + p->off_Y[g] : ws + p->off_A[g];
    if (run_act(h->act_post, p, g, stage_in, post, s)) return 1;
    ProfScope ps(h, s, PROF_OTHER, 0.0, ((double)p->C[g] * p->esize + 4.0) * (double)p->sumlen[g]);
    CK(launch_conv_post_tanh(post, dt, h->conv_post.w_raw, h->conv_post.bias, wav, (short*)pcm, seg,
                             ragged ? p->latrow_dev : nullptr, h->hop, B, p->C[g], p->R[g], p->max_frames * h->hop, s));
  }
  p->num_launches = h->launch_counter;   // what this forward actually issued (fused layers launch once)
  return 0;
}

int bvg_profile_enable(bvg_handle* h, int32_t on) {
  if (!h) return fail("bvg_profile_enable: null handle");
  h->prof_on = on != 0;
  return 0;
}

int bvg_profile_read(bvg_handle* h, double* ms, double* flops, double* bytes, int64_t* launches) {
  if (!h || !ms || !flops || !bytes || !launches) return fail("bvg_profile_read: null argument");
  for (int c = 0; c < PROF_NCLS; ++c) { ms[c] = flops[c] = bytes[c] = 0.0; launches[c] = 0; }
  const char* dump = getenv("BVG_PROF_DUMP");   // optional per-launch dump: "cls ms flops bytes" per line
  FILE* df = dump ? fopen(dump, "w") : nullptr;
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{df};
  for (const ProfRec& r : h->prof) {
    CK(cudaEventSynchronize(r.e1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (df) fprintf(df, "%d %.6f %.6e %.6e\n", r.cls, t, r.flops, r.bytes);
    ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1;
  }
  h->prof.clear();
  h->ev_used = 0;
  return 0;
}

int bvg_forward_host(bvg_handle* h, bvg_plan* p, const void* latent_host, int32_t latent_dtype, void* latent_dev,
                     const float* spk_emb, int32_t spk_batch, float* wav_host, float* wav_dev, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (!h || !p || !latent_host || !latent_dev || !wav_host || !wav_dev) return fail("bvg_forward_host: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = latent_dtype == BVG_F32 ? 4 : 2;
  const size_t in_bytes = (size_t)p->B * p->max_frames * h->cfg.gpt_dim * esz;
  const size_t out_bytes = (size_t)p->B * p->max_frames * h->hop * sizeof(float);
  CK(cudaMemcpyAsync(latent_dev, latent_host, in_bytes, cudaMemcpyHostToDevice, s));
  if (bvg_forward(h, p, latent_dev, latent_dtype, spk_emb, spk_batch, wav_dev, workspace, workspace_bytes, stream))
    return 1;
  CK(cudaMemcpyAsync(wav_host, wav_dev, out_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ------------------------------------------------------------------------------------------
// per-op entry points
// ------------------------------------------------------------------------------------------
int bvg_activation1d(const void* x, void* y, const float* log_alpha, const float* log_beta, int32_t B, int32_t C,
                     int32_t T, int32_t dtype, void* stream) {
  if (bvg_device_check()) return 1;
  if (!x || !y || !log_alpha || !log_beta) return fail("bvg_activation1d: null argument");
  if (dtype < 0 || dtype > 2) return fail("bvg_activation1d: dtype must be F32, BF16 or F16");
  if (B <= 0 || C <= 0 || T <= 0) return 0;   // the reference returns early for seq_len == 0 (.cu:194-197)
  cudaStream_t s = (cudaStream_t)stream;
  float* prm = nullptr;
  CK(cudaMallocAsync((void**)&prm, 2 * (size_t)C * sizeof(float), s));
  cudaError_t e = launch_snake_params(log_alpha, log_beta, prm, prm + C, C, s);
  if (e == cudaSuccess) e = launch_act_nct(x, y, prm, prm + C, B, C, T, dtype, s);
  cudaFreeAsync(prm, s);
  CK(e);
  return 0;
}

namespace {
struct OpTemps {
  std::vector<void*> ptrs;
  ~OpTemps() { for (void* p : ptrs) cudaFree(p); }
  int alloc(void** p, size_t bytes) {
    CK(cudaMalloc(p, bytes ? bytes : 4));
    ptrs.push_back(*p);
    return 0;
  }
};

int conv_op(bool transposed, const float* x, const float* w, const float* bias, const float* residual, float* y,
            int B, int Cin, int Cout, int T, int k, int d, int u, int mode, cudaStream_t s,
            const float* log_alpha = nullptr, const float* log_beta = nullptr, int* fused_out = nullptr) {
  if (bvg_device_check()) return 1;
  if (!x || !w || !y) return fail("conv op: null argument");
  if (Cin % 8 || Cout % 8) return fail("conv op: Cin and Cout must be multiples of 8");
  if (k < 1 || u < 1 || (!transposed && (k > BVG_MAX_TAPS || !(k & 1))) ||
      (transposed && (k % u || (k - u) % 2 || k / u > BVG_MAX_TAPS || (k - u) / 2 > 8 * u)))
    return fail("conv op: unsupported kernel size");
  ConvLayer L;
  L.transposed = transposed; L.Cin = Cin; L.Cout = Cout; L.k = k; L.d = d; L.u = u; L.setup();
  if (!transposed && (k - 1) / 2 * d > BVG_GUARD - 6) return fail("conv op: halo exceeds guard rows");
  if (mode != BVG_MODE_FP32 && mode != BVG_MODE_BF16 && mode != BVG_MODE_F16 && mode != BVG_MODE_FP32_TC)
    return fail("conv op: unknown mode %d", mode);
  const bool tc32 = mode == BVG_MODE_FP32_TC;   // fp32 storage, split-bf16 operands on the tensor cores
  const int dt = (mode == BVG_MODE_FP32 || tc32) ? 0 : (mode == BVG_MODE_BF16 ? 1 : 2);
  const size_t es = dt == 0 ? 4 : 2;
  const int Tout = T * u;
  OpTemps tmp;
  std::vector<SegDesc> seg(2 * (size_t)B);
  int Rin = BVG_GUARD, Rout = BVG_GUARD;
  for (int b = 0; b < B; ++b) {
    seg[b] = SegDesc{Rin, T}; Rin += T + BVG_GUARD;
    seg[B + b] = SegDesc{Rout, Tout}; Rout += Tout + BVG_GUARD;
  }
  Rin += BVG_TAIL_SLACK; Rout += BVG_TAIL_SLACK;
  SegDesc* seg_dev; void *xc, *yc, *rc = nullptr; float* wt; void* wu = nullptr;
  if (tmp.alloc((void**)&seg_dev, seg.size() * sizeof(SegDesc))) return 1;
  CK(cudaMemcpyAsync(seg_dev, seg.data(), seg.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, s));
  if (tmp.alloc(&xc, (size_t)Cin * Rin * es) || tmp.alloc(&yc, (size_t)Cout * Rout * es)) return 1;
  CK(cudaMemsetAsync(xc, 0, (size_t)Cin * Rin * es, s));
  CK(cudaMemsetAsync(yc, 0, (size_t)Cout * Rout * es, s));
  CK(launch_nct_to_c8(x, xc, dt, seg_dev, B, Cin, T, Rin, s));
  if (residual) {
    if (tmp.alloc(&rc, (size_t)Cout * Rout * es)) return 1;
    CK(cudaMemsetAsync(rc, 0, (size_t)Cout * Rout * es, s));
    CK(launch_nct_to_c8(residual, rc, dt, seg_dev + B, B, Cout, Tout, Rout, s));
  }
  if (tmp.alloc((void**)&wt, (size_t)Cin * Cout * k * sizeof(float))) return 1;
  if (transposed) CK(launch_repack_convt(w, wt, Cin, Cout, k, u, s));
  else CK(launch_repack_conv(w, wt, Cout, Cin, k, s));
  ConvArgs a{};
  a.dtype = dt;
  a.x = xc; a.y = yc; a.res = rc; a.w = wt; a.bias = bias; a.bias_bstride = 0; a.acc_img_scale = 1.f;
  a.seg_in = seg_dev; a.seg_out = seg_dev + B; a.Rx = Rin; a.Ry = Rout;
  a.Cin = Cin; a.Cout = Cout; a.ntaps = L.ntaps;
  for (int j = 0; j < L.ntaps; ++j) a.tap_off[j] = L.tap_off[j];
  a.u = L.u; a.p = L.p; a.q_extra = L.q_extra; a.B = B; a.max_q = T + L.q_extra;
  a.out_scale = 1.f; a.accumulate = 0;
  a.tile_prefix = nullptr; a.total_mt = 0; a.msub = 1;
  if (log_alpha && log_beta) {   // Activation1d in front of the convolution
    float* prm;
    if (tmp.alloc((void**)&prm, 2 * (size_t)Cin * sizeof(float))) return 1;
    CK(launch_snake_params(log_alpha, log_beta, prm, prm + Cin, Cin, s));
    a.act_alpha = prm; a.act_inv_beta = prm + Cin;
    const int fm = mode == BVG_MODE_BF16 ? conv_umma_fused_msub(a, fused_out && *fused_out) : 0;   // (never in the fp32 modes)
    if (fused_out) *fused_out = fm ? 1 : 0;
    if (fm) {
      a.msub = fm;
    } else {         // separate pass, as bvg_forward does for layers the fused kernel does not cover
      void* ac;
      if (tmp.alloc(&ac, (size_t)Cin * Rin * es)) return 1;
      CK(cudaMemsetAsync(ac, 0, (size_t)Cin * Rin * es, s));
      ActArgs aa{xc, ac, prm, prm + Cin, seg_dev, Rin, Cin, B, T};
      CK(launch_act_c8(aa, dt, mode == BVG_MODE_FP32 || tc32, s));
      a.x = ac; a.act_alpha = nullptr; a.act_inv_beta = nullptr;
    }
  }
  if (tc32) {
    // [hi | lo] fp16 split of the fp32 input, S [2^-11 W_hi; W_lo; W_hi] weight image, fp32 epilogue
    void* xs; float *w3, *amax, amax_h = 0.f;
    if (tmp.alloc(&xs, (size_t)Cin * Rin * 4) || tmp.alloc((void**)&w3, 3 * (size_t)Cin * Cout * k * sizeof(float)) ||
        tmp.alloc((void**)&amax, sizeof(float))) return 1;
    CK(cudaMemsetAsync(xs, 0, (size_t)Cin * Rin * 4, s));
    CK(cudaMemsetAsync(amax, 0, sizeof(float), s));
    CK(launch_split_c8((const float*)a.x, xs, seg_dev, B, Cin, Rin, T, s));
    CK(launch_absmax(wt, (size_t)Cin * Cout * k, amax, s));
    CK(cudaMemcpyAsync(&amax_h, amax, sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const float S = split3_weight_scale(amax_h);
    CK(launch_split3_weights(wt, w3, L.ntaps, Cin, L.N, S, s));
    a.x = xs; a.dtype = 2; a.f32io = 1; a.split3_chunks = Cin / 8; a.Cin = 3 * Cin; a.acc_scale = 1.f / S;
    a.msub = conv_umma_default_msub(a);
    std::vector<int> pf(B + 1, 0);
    for (int b = 0; b < B; ++b) pf[b + 1] = pf[b] + (T + L.q_extra + 128 * a.msub - 1) / (128 * a.msub);
    int* pf_dev;
    if (tmp.alloc((void**)&pf_dev, pf.size() * sizeof(int))) return 1;
    CK(cudaMemcpyAsync(pf_dev, pf.data(), pf.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    a.tile_prefix = pf_dev; a.total_mt = pf[B];
    size_t bytes = umma_weight_image_bytes(L.ntaps, 3 * Cin, L.N);
    if (!bytes || !conv_umma_supported(a)) return fail("conv op: shape not supported by the tcgen05 kernel (fp32 tensor-core mode)");
    if (tmp.alloc(&wu, bytes)) return 1;
    CK(launch_repack_umma(w3, wu, 2, L.ntaps, 3 * Cin, L.N, 1.f, false, s));
    a.w = wu;
    CK(launch_conv_umma(a, s));
  } else if (mode != BVG_MODE_FP32) {
    if (!a.act_alpha && !transposed && umma_k_packed_default(Cin, L.N)) a.k_packed = 1;   // as bvg_forward does (not with the fused kernel)
    if (!a.act_alpha && L.N >= 256) {   // same rule as bvg_forward: the small-batch variant when few CTAs would run
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      static const int allow = [] { const char* e = getenv("BVG_CONV_SMALL"); return e ? atoi(e) : 1; }();
      const long long mt1 = (long long)B * ((T + L.q_extra + 127) / 128);
      if (allow && mt1 * ((L.N + 255) / 256) * 2 <= sms) a.bn_small = 1;
    }
    if (!a.act_alpha) a.msub = conv_umma_default_msub(a);
    std::vector<int> pf(B + 1, 0);
    for (int b = 0; b < B; ++b) pf[b + 1] = pf[b] + (T + L.q_extra + 128 * a.msub - 1) / (128 * a.msub);
    int* pf_dev;
    if (tmp.alloc((void**)&pf_dev, pf.size() * sizeof(int))) return 1;
    CK(cudaMemcpyAsync(pf_dev, pf.data(), pf.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    a.tile_prefix = pf_dev; a.total_mt = pf[B];
    size_t bytes = umma_weight_image_bytes(L.ntaps, Cin, L.N, a.bn_small != 0, a.k_packed != 0);
    if (!bytes || !conv_umma_supported(a)) return fail("conv op: shape not supported by the tcgen05 kernel");
    if (tmp.alloc(&wu, bytes)) return 1;
    CK(launch_repack_umma(wt, wu, dt, L.ntaps, Cin, L.N, 1.f, a.bn_small != 0, s, nullptr, nullptr, nullptr, a.k_packed != 0));
    a.w = wu;
    CK(launch_conv_umma(a, s));
  } else {
    CK(launch_conv_simt(a, dt, s));
  }
  CK(launch_c8_to_nct(yc, dt, y, seg_dev + B, B, Cout, Tout, Rout, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}
}  // namespace

int bvg_activation1d_packed(const float* x, float* y, const float* log_alpha, const float* log_beta, int32_t B,
                            int32_t C, int32_t T, int32_t mode, void* stream) {
  if (bvg_device_check()) return 1;
  if (!x || !y || !log_alpha || !log_beta) return fail("bvg_activation1d_packed: null argument");
  if (C % 8 || B < 1 || T < 1) return fail("bvg_activation1d_packed: C must be a multiple of 8, B,T >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (mode != BVG_MODE_FP32 && mode != BVG_MODE_BF16 && mode != BVG_MODE_F16) return fail("bvg_activation1d_packed: unknown mode %d", mode);
  const int dt = mode == BVG_MODE_FP32 ? 0 : (mode == BVG_MODE_BF16 ? 1 : 2);
  const size_t es = dt == 0 ? 4 : 2;
  OpTemps tmp;
  std::vector<SegDesc> seg(B);
  int R = BVG_GUARD;
  for (int b = 0; b < B; ++b) { seg[b] = SegDesc{R, T}; R += T + BVG_GUARD; }
  R += BVG_TAIL_SLACK;
  SegDesc* seg_dev; void *xc, *yc; float* prm;
  if (tmp.alloc((void**)&seg_dev, seg.size() * sizeof(SegDesc)) || tmp.alloc(&xc, (size_t)C * R * es) ||
      tmp.alloc(&yc, (size_t)C * R * es) || tmp.alloc((void**)&prm, 2 * (size_t)C * sizeof(float))) return 1;
  CK(cudaMemcpyAsync(seg_dev, seg.data(), seg.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(xc, 0, (size_t)C * R * es, s));
  CK(cudaMemsetAsync(yc, 0, (size_t)C * R * es, s));
  CK(launch_nct_to_c8(x, xc, dt, seg_dev, B, C, T, R, s));
  CK(launch_snake_params(log_alpha, log_beta, prm, prm + C, C, s));
  ActArgs aa{xc, yc, prm, prm + C, seg_dev, R, C, B, T};
  CK(launch_act_c8(aa, dt, mode == BVG_MODE_FP32, s));
  CK(launch_c8_to_nct(yc, dt, y, seg_dev, B, C, T, R, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int bvg_conv1d(const float* x, const float* w, const float* bias, const float* residual, float* y, int32_t B,
               int32_t Cin, int32_t Cout, int32_t T, int32_t k, int32_t dilation, int32_t mode, void* stream) {
  return conv_op(false, x, w, bias, residual, y, B, Cin, Cout, T, k, dilation, 1, mode, (cudaStream_t)stream);
}

int bvg_act_conv1d(const float* x, const float* log_alpha, const float* log_beta, const float* w, const float* bias,
                   const float* residual, float* y, int32_t B, int32_t Cin, int32_t Cout, int32_t T, int32_t k,
                   int32_t dilation, int32_t mode, int32_t* fused, void* stream) {
  if (!log_alpha || !log_beta) return fail("bvg_act_conv1d: null activation parameters");
  int f = fused ? *fused : 0;   // in: request the fused kernel even if BVG_FUSE_ACT is off
  int rc = conv_op(false, x, w, bias, residual, y, B, Cin, Cout, T, k, dilation, 1, mode, (cudaStream_t)stream, log_alpha,
                   log_beta, &f);
  if (fused) *fused = f;
  return rc;
}

int bvg_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin,
                         int32_t Cout, int32_t T, int32_t k, int32_t u, int32_t mode, void* stream) {
  return conv_op(true, x, w, bias, nullptr, y, B, Cin, Cout, T, k, 1, u, mode, (cudaStream_t)stream);
}

}  // extern "C"
