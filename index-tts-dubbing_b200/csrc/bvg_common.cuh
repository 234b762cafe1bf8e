// Internal declarations shared by the b200vgan kernels (sm_100a only).
//
// Data layout in HBM ("packed c8"): an activation tensor with C channels (C % 8 == 0) over a
// batch of variable-length segments is stored as  [C/8][R][8]  where R is ONE packed time axis:
//
//     | G zero rows | segment 0 (len0 rows) | G zero rows | segment 1 | ... | G + slack zero rows |
//
// element (b, t, c)  ->  ((c/8) * R + off[b] + t) * 8 + (c % 8).
// The guard rows are zeroed once (workspace init) and never written afterwards, so the "same"
// zero padding of every convolution, and the halo of every tile, is a plain in-bounds read.
// An 8-channel group of one time step is one 16-byte (bf16) / 32-byte (fp32) vector: that is
// the K-chunk ("core matrix row") of the tcgen05 no-swizzle K-major operand layout, so an
// activation tile [rows][64 ch] maps to UMMA shared memory with a row stride of 16 B and any
// dilated tap is a descriptor start-address shift (see bvg_conv_umma.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define BVG_GUARD 32          // zero rows between segments (>= max conv halo 25 + act halo 5)
#define BVG_TAIL_SLACK 640    // extra zero rows after the last segment (tile overrun of loads: up to 512 + halo)
#define BVG_MAX_TAPS 11

struct SegDesc {   // one per (stage, segment)
  int off;         // first row of the segment on the packed time axis
  int len;         // valid rows
};

struct ConvArgs {
  const void* x;        // [Cin/8][Rx][8]
  void* y;              // [Cout/8][Ry][8]
  const void* res;      // residual, geometry of y, or nullptr
  const float* res_scale;   // optional per-output-channel factor on the residual (tcgen05 kernel; nullptr = 1)
  const void* w;        // kernel-specific weight image
  const float* bias;    // bias[b * bias_bstride + co]
  const SegDesc* seg_in;
  const SegDesc* seg_out;
  int bias_bstride;
  float acc_img_scale;  // value on the diagonal of the weight image's second identity set (0 if unknown)
  int Rx, Ry;
  int Cin, Cout;        // Cout = channels of y; GEMM N = u * Cout
  int ntaps;
  int tap_off[BVG_MAX_TAPS];  // input row = q + tap_off[j]
  int u, p;             // output row = q*u + phase - p  (u = 1, p = 0: plain conv)
  int q_extra;          // tiles cover q in [0, len_in + q_extra)
  int B;
  int max_q;            // max over segments of (len_in + q_extra)
  float out_scale;      // y = (acc + bias + res) * out_scale (+ y_old if accumulate)
  int accumulate;
  int dtype;            // storage type of x / y / res: 0 = fp32, 1 = bf16, 2 = fp16 (tcgen05 kernel: 1 or 2)
  // tcgen05 kernel only: m-tile table (tile = 128*msub rows of q per segment)
  const int* tile_prefix;   // [B+1] prefix sum of tiles per segment (device)
  int total_mt;             // tile_prefix[B]
  int msub;                 // 1, 2 or 4
  int bn_small;             // 1: `w` is the 64-column n-tile image (small-batch variant of a wide layer), msub must be 1
  int k_packed;             // 1: `w` is the K-packed image of a 24-channel layer (taps share K-steps, see make_tiling)
  // fp32 tensor-core mode (bvg_conv_umma.cu, F32IO): x is a SPLIT tensor -- split3_chunks 8-channel chunks of fp16 "hi"
  // values followed by as many chunks of fp16 remainders 2^11 (x - hi) (22 significand bits together, split_f32 below) --
  // and Cin counts 3 x the real input channels: the GEMM multiplies [lo | hi | hi] with the fp16 weight image
  // S [2^-11 W_hi; W_lo; W_hi] (S = the layer's power-of-two scale, acc_scale = 1 / S); y / res are fp32 packed tensors
  // programmatic dependent launch between INDEPENDENT convolutions (the same step of the nk AMP blocks, bvg_api.cu):
  //   0  trigger, then wait for the previous kernel before touching memory (the default chain);
  //   1  wait, THEN trigger: a successor of kind 2 may assume everything before this kernel has completed;
  //   2  independent of its predecessor (which is of kind 1 or 2): no wait before the work, so its first tiles run under
  //      the predecessor's tail; it waits just before it exits, which keeps completion order transitive for whoever
  //      waits on it.
  int pdl_mode;
  float acc_scale;   // f32io: factor on the accumulator (undoes the power-of-two scale of the fp16 weight image)
  int f32io;
  int split3_chunks;
  // fused Activation1d (tcgen05 kernel, bf16): when set, x is the RAW input and the kernel applies the
  // activation with these per-input-channel parameters while staging its A operand
  const float* act_alpha;
  const float* act_inv_beta;
};

struct ActArgs {
  const void* x;
  void* y;
  const float* alpha;     // exp(log_alpha)  [C]
  const float* inv_beta;  // 1 / (exp(log_beta) + 1e-9)  [C]
  const SegDesc* seg;
  int R, C, B, max_len;
  // 16-bit tensor-core kernel only: x holds 2*alpha*x_true per channel (folded into the producing convolution's weights)
  // and y receives 2*alpha*y_true (unfolded by the consuming convolution's weights): saves the per-sample argument multiply
  int prescaled;
  // tensor-core kernel only: up to two more activations of the same geometry (other tensors, other parameters) in the same
  // launch -- the three AMP blocks of a stage run in lockstep (bvg_api.cu) and share their Activation1d launches
  int extra_jobs;
  const void* xj[2];
  void* yj[2];
  const float* alphaj[2];
  const float* inv_betaj[2];
  int fast_fp32;   // fp32 kernel only: fp32 tensor-core mode's activation arithmetic (see launch_act_c8_v2)
  int split_out;   // fp32 kernel only: y is a split bf16 tensor [hi chunks | lo chunks] (fp32 tensor-core mode)
};

// -------------------------------------------------------------------------------------------
template <typename T> struct Vec8;   // 8 consecutive channels of one time step
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

template <> struct Vec8<__half> {
  float v[8];
  __device__ __forceinline__ void load(const __half* p) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ void store(__half* p) const {   // saturating: a value beyond fp16's range becomes +-65504, not inf
    uint4 r;
    uint32_t* h = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// fp32 tensor-core mode: x = hi + lo with hi = fp16(x) (saturating) and the remainder kept as fp16(2^11 (x - hi)) -- 22
// significand bits in two fp16 numbers; the 2^11 keeps the remainder out of fp16's subnormal range and is undone by the
// weight image's second block (bvg_conv_umma.cu, split3_weights_kernel).
constexpr float BVG_SPLIT_LO_SCALE = 2048.f;
__device__ __forceinline__ float f16_round_sat(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ void split_f32(float v, float& hi, float& lo_scaled) {
  hi = f16_round_sat(v);
  lo_scaled = (v - hi) * BVG_SPLIT_LO_SCALE;   // (rounded to fp16 by the Vec8<__half> store)
}

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) {
  return __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
}

// kaiser_sinc_filter1d(0.25, 0.3, 12) -- reference alias_free_torch/filter.py:29-58; the taps are
// symmetric, sum to 1 and are shared by UpSample1d and DownSample1d.
#define BVG_F0 0.00202896469f
#define BVG_F1 0.00938946567f
#define BVG_F2 (-0.0255434588f)
#define BVG_F3 (-0.0576573834f)
#define BVG_F4 0.128572583f
#define BVG_F5 0.443209797f

// Programmatic dependent launch: a kernel launched through launch_pdl may start (block scheduling, prologue) while the
// previous kernel of the stream drains; it must execute pdl_wait() before it touches anything an earlier kernel wrote
// or still reads, and calls pdl_trigger() to let ITS successor do the same.  Both are no-ops in a plain launch.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool bvg_pdl_enabled();   // BVG_PDL (default 1)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = bvg_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// host-side launchers (each returns cudaGetLastError())
cudaError_t launch_act_c8(const ActArgs& a, int dtype, bool precise, cudaStream_t s);
cudaError_t launch_act_c8_mma(const ActArgs& a, int dtype, cudaStream_t s);   // dtype 1 = bf16, 2 = fp16
cudaError_t launch_act_c8_v2(const ActArgs& a, int dtype, bool precise, int rt, cudaStream_t s);
cudaError_t launch_act_nct(const void* x, void* y, const float* alpha, const float* inv_beta, int B, int C,
                           int T, int dtype, cudaStream_t s);
cudaError_t launch_conv_simt(const ConvArgs& a, int dtype, cudaStream_t s);
cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s);
bool conv_umma_supported(const ConvArgs& a);
int conv_umma_default_msub(const ConvArgs& a);
int conv_umma_fused_msub(const ConvArgs& a, bool force = false);   // 0 = this conv cannot take the fused activation
