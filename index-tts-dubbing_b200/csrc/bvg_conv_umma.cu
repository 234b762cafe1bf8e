// tcgen05 (5th-gen tensor core) implicit-GEMM 1-D convolution for sm_100a, bf16 in / fp32 accumulate.
// Version 3: persistent CTAs, 128/256/512-row tiles, multi-stage TMEM accumulators (the epilogue of
// tile i overlaps the MMAs of tile i+1), weight-stationary B operand for the narrow stages, a
// single-thread uniform-datapath MMA issuer and a 32-column-at-a-time epilogue.
//
// GEMM view (same as bvg_conv_simt.cu):  D[q, n] = sum_tap sum_ci X[q + tap_off[tap], ci] * W[tap][ci][n]
//   M tile = MSUB x 128 time rows (UMMA_M = 128, cta_group::1, MSUB accumulators per tile),
//   N tile = BN <= 256 columns, K = taps x Cin walked as k-blocks of KC 8-channel chunks x taps.
//
// Operand staging (no tensor maps needed -- the HBM layouts ARE the shared-memory images):
//   A  activations, packed c8 layout [Cin/8][R][8] bf16.  For one k-block the producer issues KC bulk
//      copies (cp.async.bulk + mbarrier complete_tx) of AROWS = 128*MSUB + span contiguous rows, one
//      per 8-channel chunk, giving the UMMA no-swizzle K-major layout [chunk][row][16 B]: core matrix
//      = 8 rows x 16 B contiguous, SBO (next 8 rows) = 128 B, LBO (next K chunk) = ASTRIDE*16 B.
//      Rows are 16 B apart, so a dilated tap is a descriptor start-address shift of tap_off*16 B and
//      the s-th 128-row sub-tile a shift of s*128*16 B: the tile (+halo) is loaded ONCE per k-block
//      and reused by all taps.  Zero padding / halo come from the zero guard rows of the layout.
//   B  weights, pre-packed by launch_repack_umma into per-(n-tile, k-block, tap) images
//      [chunk][n][16 B] (LBO = BN*16 B): one bulk copy per pipeline stage; all sub-tiles reuse it.
//      When a layer's whole weight slice fits (narrow stages) it is loaded once per CTA and kept.
//   D  fp32 accumulators in TMEM: ACC stages x MSUB accumulators x BNC columns (<= 512 columns).
//
// Warp roles (320 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4; the two warps of a quarter alternate over
// the (sub-tile, 32-column group) work items).
//
// Residual / accumulate through the tensor core: for square layers with Cin <= 96 the residual tile
// (and, for accumulating layers, the old output tile) is loaded as extra k-blocks and multiplied by
// identity (resp. 1/out_scale * identity) weight images appended to the layer's image, so the epilogue
// of those layers issues no global loads (see use_res_mma / use_acc_mma).
//
// conv_umma_kernel<true> (opt-in, BVG_FUSE_ACT=1) additionally computes Activation1d in 10 extra warps
// between the producer (raw rows) and the MMA issuer (activated A operand); measured slower than the
// separate activation pass in round 1, kept for the next round.
//
// Round 2 additions: fp16 storage (template F16: kind::f16 on fp16 operands, saturating stores); K-packed images for the
// 24-channel layers and 64-column n-tiles for small batches (make_tiling); and the fp32 tensor-core mode (template F32IO):
// the input is a split fp16 tensor [hi | 2^11 lo] walked as 3 Cin channels against S [2^-11 W_hi; W_lo; W_hi], so that
// x_lo W_hi + x_hi W_lo + x_hi W_hi accumulate with every product exact, correction blocks first (the tensor core truncates
// its fp32 accumulator once per MMA), and bias / residual / old output / store are fp32.
// Launch structure: programmatic dependent launch, with ConvArgs::pdl_mode letting the independent convolutions of a
// lockstep AMP step start under their predecessor's tail.
//
// What the measurements behind this structure were (tools/umma_bench.cu, profiles/r1_umma_*):
//   * one M=128 MMA costs max(N/2, ~40..50) cycles when issued from uniform registers, but ~124 when
//     its descriptors are built in a divergent single-lane region (R2UR per operand) -> the issuer is
//     ONE thread chosen with elect.sync (ptxas then emits uniform-datapath code for the whole loop);
//   * consecutive MMAs into the same accumulator serialise -> the sub-tile loop is innermost;
//   * narrow stages are bound by per-tile fixed costs (barrier round trips, tile decode, epilogue
//     latency), not by MMA or HBM -> big tiles (MSUB = 4), x32 TMEM loads, decode prefetch;
//   * the tensor pipe queues only a few MMAs: every cycle the issuing thread spends between taps is
//     idle pipe time (BVG_CONV_TRACE showed ~400 cycles per tap) -> incremental 32-bit descriptors and
//     straight-line K-step sequences; tools/umma_bench2.cu gives the per-shape floor (46 cycles for
//     M128 x N32 x K16: the A operand's shared-memory reads, not the math).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include <cuda_fp16.h>

#include "bvg_act_core.cuh"
#include "bvg_common.cuh"
#include "bvg_misc.cuh"

namespace {

constexpr int MAXSPAN = 50;
constexpr int SMEM_BUDGET = 200 * 1024;   // operand stages; barriers etc. come on top
constexpr int SMEM_MAX = 224 * 1024;
constexpr int MAX_STAGES = 12;
constexpr int SBIAS_MAX = 512;             // columns of batch-independent bias kept in shared memory
constexpr int EPIW = 8;                   // epilogue warps
constexpr int NTHREADS = 64 + 32 * EPIW;
// fused mode (Activation1d computed in the conv's producer stage): 4 epilogue warps + 10 activation warps
constexpr int EPIW_FUSED = 4;
constexpr int NACT = 10;
constexpr int NTHREADS_FUSED = 64 + 32 * (EPIW_FUSED + NACT);
constexpr int ACT_RT = 25;                // rows per activation thread (odd: conflict-free 4-byte columns)

struct UmmaTiling {
  int KC, NKB;      // 8-channel chunks per k-block, k-blocks
  int BN, NT;       // columns per CTA tile (multiple of 16, <= 256), n tiles
  int BNC;          // TMEM columns per accumulator (BN rounded up to 32)
  // K-packed mode (Cin = 24): a tap holds 3 real 8-channel chunks, i.e. 1.5 K-steps; instead of padding every tap to 2
  // K-steps with a zero chunk, the (tap, chunk) items are laid out back to back: K-step s multiplies items 2s and 2s+1,
  // which may belong to different taps (the A descriptor's start address and K-chunk stride are per K-step then).
  int packed, nks, nstg;   // nstg: weight stages per k-block (= ntaps when not packed; a stage = KC chunks = KC/2 K-steps)
  bool ok;
};

inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }
inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Tiling that only depends on the layer (the weight image layout depends on it).  `small`: 64-column n-tiles -- the
// small-batch variant of the wide layers: with only a few m-tiles (one short utterance) a 256-column n-tile leaves
// 3-12 CTAs streaming 4-13 MB of weights each; four times as many CTAs stream a quarter each.
inline UmmaTiling make_tiling(int ntaps, int Cin, int N, bool small = false, bool packed = false) {
  UmmaTiling t{};
  t.ok = false;
  t.packed = 0; t.nks = 0; t.nstg = ntaps;
  if (Cin % 8 || N % 8 || ntaps < 1 || ntaps > BVG_MAX_TAPS) return t;
  const int cin_pad = round_up_i(Cin, 16);
  if (cin_pad % 64 == 0) { t.KC = 8; t.NKB = cin_pad / 64; }
  else if (cin_pad <= 128) { t.KC = cin_pad / 8; t.NKB = 1; }
  else {
    // 3 x C of the split mode (144, 288 channels): 6-chunk k-blocks when they divide the channels exactly (no padded
    // K-steps), else 8-chunk k-blocks with a partly padded last one (zero weight rows)
    const int chunks = cin_pad / 8;
    if (chunks % 6 == 0) { t.KC = 6; t.NKB = chunks / 6; }
    else { t.KC = 8; t.NKB = round_up_i(cin_pad, 64) / 64; }
  }
  // N tile: 256 columns by default -- a 128 x 256 x 16 MMA is the only cta_group::1 shape that runs at
  // the tensor-pipe floor (tools/umma_bench.cu) and it halves the A-tile re-reads per output column.
  static const int bnmax_env = [] { int v = env_int("BVG_CONV_BNMAX", 256); return (v == 128 || v == 192) ? v : 256; }();
  const int bnmax = small ? 64 : bnmax_env;
  const int n_pad = round_up_i(N, 16);
  t.NT = (n_pad + bnmax - 1) / bnmax;
  t.BN = round_up_i((n_pad + t.NT - 1) / t.NT, 16);
  t.BNC = round_up_i(t.BN, 32);
  if (packed) {
    if (Cin != 24 || t.NT != 1) return t;   // only the 24-channel layers (KC = 4: three real chunks + one zero chunk)
    t.packed = 1;
    t.nks = (3 * ntaps + 1) / 2;
    t.nstg = (t.nks + 1) / 2;
  }
  t.ok = true;
  return t;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// elect.sync: one lane of the (converged) warp; unlike `lane == 0` it tells ptxas that the guarded
// region runs with a single active thread, so uniform-datapath code is generated for it.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread = lane (time row), r[j] = column j
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: start addr, LBO (K-chunk stride), SBO (8-row
// group stride), version 1 (Blackwell).  All byte quantities are encoded >> 4.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// F16: the tensors are stored in fp16 instead of bf16 (stores saturate at +-65504 instead of producing inf)
template <bool F16>
__device__ __forceinline__ void unpack_add(const uint4& p, float (&v)[8]) {
  if constexpr (F16) {
    const __half2* h = reinterpret_cast<const __half2*>(&p);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&p);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
  }
}
template <bool F16>
__device__ __forceinline__ void unpack_fma(const uint4& p, float (&v)[8], const float4& s0, const float4& s1) {
  float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  unpack_add<F16>(p, r);
  v[0] = fmaf(r[0], s0.x, v[0]); v[1] = fmaf(r[1], s0.y, v[1]); v[2] = fmaf(r[2], s0.z, v[2]); v[3] = fmaf(r[3], s0.w, v[3]);
  v[4] = fmaf(r[4], s1.x, v[4]); v[5] = fmaf(r[5], s1.y, v[5]); v[6] = fmaf(r[6], s1.z, v[6]); v[7] = fmaf(r[7], s1.w, v[7]);
}
template <bool F16>
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  if constexpr (F16) {
    uint32_t* oh = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(oh[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
  } else {
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  }
  return o;
}

// MMAs of one (k-block, tap): nk16 K-steps x MS sub-tile accumulators, sub-tile loop innermost so that
// consecutive MMAs go to different accumulators.  MS is a compile-time constant to keep the single
// issuing thread's loop branch-free.
template <int MS>
__device__ __forceinline__ void issue_tap(uint32_t d0, uint32_t bnc, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accum, int nk16, uint32_t a_kstep, uint32_t b_kstep) {
#pragma unroll 4
  for (int k16 = 0; k16 < nk16; ++k16) {
    const uint32_t af = accum | (uint32_t)k16;
    umma_bf16(d0, adesc, bdesc, idesc, af);
    if (MS >= 2) umma_bf16(d0 + bnc, adesc + 128u, bdesc, idesc, af);   // +128 rows
    if (MS == 4) {
      umma_bf16(d0 + 2 * bnc, adesc + 256u, bdesc, idesc, af);
      umma_bf16(d0 + 3 * bnc, adesc + 384u, bdesc, idesc, af);
    }
    adesc += a_kstep; bdesc += b_kstep;
  }
}

// ---- warp-level MMA pieces of the fused activation (same scheme as csrc/bvg_act3.cu) ----
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stsm_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void wmma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void wmma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk_f16(float lo, float hi) {   // saturating, see bvg_act3.cu
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pk_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float fir_tap(int k) {   // f[k], 0 outside 0..11
  const float f[6] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5};
  if (k < 0 || k > 11) return 0.f;
  return f[k < 6 ? k : 11 - k];
}

struct UmmaKernelArgs {
  ConvArgs c;
  const int* tile_prefix;   // [B+1] prefix sum of m-tiles per segment (tile = 128*MSUB rows)
  int total_mt;             // tile_prefix[B]
  int KC, NKB, BN, BNC, NT;
  int MSUB;                 // 128-row sub-tiles per tile (1, 2 or 4)
  int ACC;                  // TMEM accumulator stages
  int tmem_cols;
  int NA, NB;               // smem pipeline depths
  int b_resident;           // weights loaded once per CTA (NB == NKB*ntaps, NT == 1)
  int astride;              // rows between K chunks of an A stage
  int a_stage_bytes, b_stage_bytes;
  int kc_last_load;         // real (non-padding) chunks of the last k-block
  int minoff, span;
  // fused Activation1d (FUSE kernel only): the producer loads RAW rows into R stages, activation warps
  // write the activated tile into the A stages
  const float* act_alpha;   // exp(log_alpha) per input channel
  const float* act_inv_beta;
  int NR;                   // raw stages
  int rstride, r_stage_bytes;
  int nblk;                 // 8-row output column tiles per chunk (astride = 8 * nblk)
  int act_ngc, act_jr;      // fused activation work items: act_ngc row ranges per chunk of act_jr column tiles each
  // residual folded into the accumulator: after the conv k-blocks, NKB more k-blocks take the residual
  // tile as A operand against identity weight images (appended to the layer's image), so the epilogue
  // issues no global loads (narrow stages were bound by that latency chain)
  int res_mma;
  // same trick for `accumulate` (y = old + ...): the old output tile x (1/out_scale) * identity (second image set)
  int acc_mma;
  int n_extra;              // res_mma + acc_mma: extra k-block groups after the convolution's
  int packed, nks, nstg;    // K-packed mode (see UmmaTiling); nstg = weight stages of the convolution part per k-block
  unsigned long long* trace;   // optional event trace of CTA 0 (BVG_CONV_TRACE), nullptr normally
  int debug;                // tuning aid (BVG_CONV_DEBUG): 1 = epilogue skips global memory, 2 = no MMAs, 4 = no A loads
};

struct TileRef { int nt, b, q0; };

// F32IO: the fp32 tensor-core mode -- split [hi | lo] bf16 input, fp32 output / residual / accumulate (see ConvArgs::f32io)
template <bool FUSE, int EPW = EPIW, bool F16 = false, bool F32IO = false>
__global__ void __launch_bounds__(FUSE ? NTHREADS_FUSED : 64 + 32 * EPW, 1) conv_umma_kernel(const UmmaKernelArgs ka) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = ka.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = ka.total_mt * ka.NT;
  const int TM = 128 * ka.MSUB;

  // smem carve-up: [R stages (fused only)][A stages][B stages][barriers][tmem slot]
  uint8_t* r_smem = smem;
  uint8_t* a_smem = smem + (FUSE ? (size_t)ka.NR * ka.r_stage_bytes : 0);
  constexpr int epiw = FUSE ? EPIW_FUSED : EPW;
  uint8_t* b_smem = a_smem + (size_t)ka.NA * ka.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)ka.NB * ka.b_stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto A_FULL = [&](int s) { return bar0 + 8u * s; };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (ka.NA + s); };
  auto B_FULL = [&](int s) { return bar0 + 8u * (2 * ka.NA + s); };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 * ka.NA + ka.NB + s); };
  auto T_FULL = [&](int s) { return bar0 + 8u * (2 * ka.NA + 2 * ka.NB + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (2 * ka.NA + 2 * ka.NB + ka.ACC + s); };
  auto R_FULL = [&](int s) { return bar0 + 8u * (2 * ka.NA + 2 * ka.NB + 2 * ka.ACC + s); };
  auto R_EMPTY = [&](int s) { return bar0 + 8u * (2 * ka.NA + 2 * ka.NB + 2 * ka.ACC + ka.NR + s); };
  // fused mode: TMA-filled (residual) uses of an A stage complete on their own barrier, because A_FULL
  // counts the activation warps' arrivals there
  auto AR_FULL = [&](int s) { return FUSE ? bar0 + 8u * (2 * ka.NA + 2 * ka.NB + 2 * ka.ACC + 2 * ka.NR + s) : A_FULL(s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ka.NA + 2 * ka.NB + 2 * ka.ACC + 2 * ka.NR + (FUSE ? ka.NA : 0));

  // padding chunks (Cin not a multiple of 16) must read as zero: clear the A stages once
  if (ka.kc_last_load < ka.KC) {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* p = reinterpret_cast<uint4*>(a_smem);
    const int n16 = ka.NA * ka.a_stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) p[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // batch-independent bias (the AMP-block convolutions): staged once in shared memory, zero padded to
  // whole 32-column groups, so the epilogue reads it with broadcast LDS instead of dependent L1 loads
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 28) + 15) & ~(uintptr_t)15);   // [2..27]: the issuer's tap / K-step table
  const bool use_sbias = a.bias != nullptr && a.bias_bstride == 0 && ka.NT * ka.BN <= SBIAS_MAX;
  if (use_sbias) {
    const int N_ = a.u * a.Cout;
    for (int i = threadIdx.x; i < SBIAS_MAX; i += blockDim.x) sbias[i] = i < N_ ? a.bias[i] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < ka.NA; ++s) { mbar_init(A_FULL(s), FUSE ? (uint32_t)NACT : 1u); mbar_init(A_EMPTY(s), 1); }
    if (FUSE) {
      for (int s = 0; s < ka.NR; ++s) { mbar_init(R_FULL(s), 1); mbar_init(R_EMPTY(s), (uint32_t)NACT); }
      for (int s = 0; s < ka.NA; ++s) mbar_init(AR_FULL(s), 1);
    }
    for (int s = 0; s < ka.NB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
    for (int s = 0; s < ka.ACC; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), (uint32_t)epiw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)ka.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: the set-up above (barriers, TMEM, bias staging -- launch constants only) may run
  // while the previous kernel drains; activations are read and outputs written only after it has completed
  if (a.pdl_mode == 1) {
    pdl_wait();
    pdl_trigger();
  } else {
    pdl_trigger();
    if (a.pdl_mode != 2) pdl_wait();
  }

  // low-overhead event trace (CTA 0 only): per-role ring in shared memory, flushed at kernel end
  unsigned long long* tr_smem = reinterpret_cast<unsigned long long*>(smem + (SMEM_MAX - 4 * 1024 * 8));
  int tr_n = 0;
  auto TRACE = [&](int role, int ev, int tile) {
    if (ka.trace && blockIdx.x == 0 && tr_n < 1024) {
      tr_smem[role * 1024 + tr_n] = ((unsigned long long)ev << 56) | ((unsigned long long)(tile & 0xffff) << 40) | (clock64() & 0xffffffffffULL);
      ++tr_n;
    }
  };
  const int arows = TM + ka.span;          // rows loaded per chunk
  const int acc_cols = ka.MSUB * ka.BNC;   // TMEM columns per accumulator stage

  // tile id -> (n tile, segment, first row).  Tiles are n-tile-major so that CTAs running at the same
  // time stream the same slice of the weights (L2 reuse).  Each role walks its tiles in increasing
  // order, so the segment lookup is incremental; roles decode the NEXT tile before they block on a
  // barrier so the table loads are off the critical path.
  int dec_nt = -1, dec_b = 0;
  auto decode = [&](int t) {
    TileRef r;
    r.nt = t / ka.total_mt;
    const int mt = t - r.nt * ka.total_mt;
    if (r.nt != dec_nt) { dec_nt = r.nt; dec_b = 0; }
    while (mt >= __ldg(ka.tile_prefix + dec_b + 1)) ++dec_b;
    r.b = dec_b;
    r.q0 = (mt - __ldg(ka.tile_prefix + dec_b)) * TM;
    return r;
  };

  if (warp == 0) {
    // ===================== producer (one thread) =====================
    if (elect_one()) {
      const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(a.x);
      const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.w);
      int sa = 0, pa = 0, sb = 0, pb = 0, sr = 0, pr = 0;
      bool first = true;
      int t = blockIdx.x;
      TileRef cur = t < total_tiles ? decode(t) : TileRef{0, 0, 0};
      const __nv_bfloat16* resg = reinterpret_cast<const __nv_bfloat16*>(a.res);
      long long row0 = t < total_tiles ? (long long)a.seg_in[cur.b].off + cur.q0 + ka.minoff : 0;
      long long rrow0 = t < total_tiles ? (long long)a.seg_out[cur.b].off + cur.q0 : 0;
      int seglen = (FUSE && t < total_tiles) ? a.seg_in[cur.b].len : 0;
      while (t < total_tiles) {
        TRACE(0, 0, t);
        const int tn = t + gridDim.x;
        const int nt = cur.nt;
        const long long row0c = row0, rrow0c = rrow0;
        const int Lc = seglen, q0c = cur.q0;
        if (tn < total_tiles) {   // prefetch the next tile's coordinates
          cur = decode(tn);
          row0 = (long long)a.seg_in[cur.b].off + cur.q0 + ka.minoff;
          rrow0 = (long long)a.seg_out[cur.b].off + cur.q0;
          if (FUSE) seglen = a.seg_in[cur.b].len;
        }
        for (int kb = 0; kb < ka.NKB; ++kb) {
          const int kcl = (kb == ka.NKB - 1) ? ka.kc_last_load : ka.KC;
          if constexpr (FUSE) {
            // raw rows [row0 - 8, row0 + 8*nblk + 24) of every chunk: the activation warps consume them
            const int rrows = ka.rstride;
            mbar_wait(R_EMPTY(sr), pr ^ 1);
            TRACE(0, 2, t);
            mbar_expect_tx(R_FULL(sr), (uint32_t)(kcl * rrows * 16));
            const uint32_t rdst = smem_u32(r_smem + (size_t)sr * ka.r_stage_bytes);
            // Slab row j holds segment time tR0 + j.  Every row before the segment and the five rows after it
            // carry the replicate padding of Activation1d's input (x[0] / x[L-1]) instead of the layout's zero
            // guard rows, so the activation warps can run every column tile with the interior formulas (and
            // never multiply an out-of-segment bit pattern by a zero tap); the pieces are disjoint (no
            // ordering exists between bulk copies).  Interior tiles: one piece.
            const int tR0 = q0c + ka.minoff - 8;
            auto clampj = [&](int v) { return v < 0 ? 0 : (v > rrows ? rrows : v); };
            const int jl0 = 0, jl1 = clampj(-tR0), jr0 = clampj(Lc - tR0), jr1 = clampj(Lc + 5 - tR0);
            for (int c = 0; c < kcl; ++c) {
              const __nv_bfloat16* src = xg + ((size_t)(kb * ka.KC + c) * a.Rx + row0c - 8) * 8;   // slab row 0
              const uint32_t dst = rdst + (uint32_t)(c * ka.rstride) * 16;
              if (jl0 > 0) bulk_g2s(dst, src, (uint32_t)(jl0 * 16), R_FULL(sr));
              for (int j = jl0; j < jl1; ++j) bulk_g2s(dst + j * 16, src + (size_t)(-tR0) * 8, 16u, R_FULL(sr));          // x[0]
              if (jr0 > jl1) bulk_g2s(dst + jl1 * 16, src + (size_t)jl1 * 8, (uint32_t)((jr0 - jl1) * 16), R_FULL(sr));
              for (int j = jr0; j < jr1; ++j) bulk_g2s(dst + j * 16, src + (size_t)(Lc - 1 - tR0) * 8, 16u, R_FULL(sr));  // x[L-1]
              if (rrows > jr1) bulk_g2s(dst + jr1 * 16, src + (size_t)jr1 * 8, (uint32_t)((rrows - jr1) * 16), R_FULL(sr));
            }
            if (++sr == ka.NR) { sr = 0; pr ^= 1; }
            if (++sa == ka.NA) { sa = 0; pa ^= 1; }   // the activation warps fill this A stage
          } else {
          mbar_wait(A_EMPTY(sa), pa ^ 1);
          TRACE(0, 2, t);
          if (ka.debug & 4) {
            mbar_arrive(A_FULL(sa));
          } else {
            mbar_expect_tx(A_FULL(sa), (uint32_t)(kcl * arows * 16));
            const uint32_t adst = smem_u32(a_smem + (size_t)sa * ka.a_stage_bytes);
            for (int c = 0; c < kcl; ++c) {
              int cc = kb * ka.KC + c;
              if (F32IO) {
                // GEMM channel blocks [lo | hi | hi] over the tensor's [hi | lo] chunks (the small terms accumulate first)
                const int c8 = a.split3_chunks;
                cc = cc < c8 ? cc + c8 : (cc < 2 * c8 ? cc - c8 : cc - 2 * c8);
              }
              const __nv_bfloat16* src = xg + ((size_t)cc * a.Rx + row0c) * 8;
              bulk_g2s(adst + (uint32_t)(c * ka.astride) * 16, src, (uint32_t)(arows * 16), A_FULL(sa));
            }
          }
          if (++sa == ka.NA) { sa = 0; pa ^= 1; }
          }
          if (!ka.b_resident || first) {
            for (int tap = 0; tap < ka.nstg; ++tap) {
              if (!ka.b_resident) mbar_wait(B_EMPTY(sb), pb ^ 1);
              mbar_expect_tx(B_FULL(sb), (uint32_t)ka.b_stage_bytes);
              const uint8_t* src = wimg + ((size_t)(nt * ka.NKB + kb) * ka.nstg + tap) * ka.b_stage_bytes;
              bulk_g2s(smem_u32(b_smem + (size_t)sb * ka.b_stage_bytes), src, (uint32_t)ka.b_stage_bytes, B_FULL(sb));
              if (++sb == ka.NB) { sb = 0; pb ^= 1; }
            }
          }
        }
        for (int e = 0; e < ka.n_extra; ++e) {
          // residual / old-output k-blocks: the tile's TM output rows of the source, placed at slab row -minoff
          const bool is_old = ka.acc_mma && e == ka.n_extra - 1;
          const __nv_bfloat16* srcg = is_old ? reinterpret_cast<const __nv_bfloat16*>(a.y) : resg;
          const int img_set = is_old ? 1 : 0;
          for (int kb = 0; kb < ka.NKB; ++kb) {
            const int kcl = (kb == ka.NKB - 1) ? ka.kc_last_load : ka.KC;
            mbar_wait(A_EMPTY(sa), pa ^ 1);
            mbar_expect_tx(AR_FULL(sa), (uint32_t)(kcl * TM * 16));
            const uint32_t adst = smem_u32(a_smem + (size_t)sa * ka.a_stage_bytes) - (uint32_t)(ka.minoff * 16);
            for (int c = 0; c < kcl; ++c) {
              const __nv_bfloat16* src = srcg + ((size_t)(kb * ka.KC + c) * a.Ry + rrow0c) * 8;
              bulk_g2s(adst + (uint32_t)(c * ka.astride) * 16, src, (uint32_t)(TM * 16), AR_FULL(sa));
            }
            if (++sa == ka.NA) { sa = 0; pa ^= 1; }
            if (!ka.b_resident || first) {
              if (!ka.b_resident) mbar_wait(B_EMPTY(sb), pb ^ 1);
              mbar_expect_tx(B_FULL(sb), (uint32_t)ka.b_stage_bytes);
              const uint8_t* src = wimg + ((size_t)ka.NKB * (ka.nstg + img_set) + kb) * ka.b_stage_bytes;   // identity images (NT == 1)
              bulk_g2s(smem_u32(b_smem + (size_t)sb * ka.b_stage_bytes), src, (uint32_t)ka.b_stage_bytes, B_FULL(sb));
              if (++sb == ka.NB) { sb = 0; pb ^= 1; }
            }
          }
        }
        first = false;
        t = tn;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    // Single-thread region selected with elect.sync: every address / descriptor computation is then
    // uniform and ptxas keeps it in uniform registers (tcgen05.mma takes its descriptors from URs).
    // The tensor pipe queues only a couple of MMAs, so every cycle this thread spends between two taps
    // is a cycle the pipe idles (measured: ~400 cycles per tap with a naive loop, r1 traces): the tap
    // loop is kept to a handful of instructions -- 32-bit descriptor words advanced incrementally, the
    // per-tap row shifts read from a small shared-memory table one tap ahead, MSUB a compile-time constant.
    if (elect_one()) {
      // tap offsets are an arithmetic progression (dilated conv: step d; transposed conv: step -1; checked in configure):
      // the per-tap row shift advances by a uniform add -- no table, no vector-to-uniform register moves in the issue loop
      const uint32_t sh0 = (uint32_t)(a.tap_off[0] - ka.minoff);
      const uint32_t sh_step = a.ntaps > 1 ? (uint32_t)(a.tap_off[1] - a.tap_off[0]) : 0u;
      auto run = [&](auto ms_tag, auto nk_tag, auto pk_tag) {
        constexpr int MS = decltype(ms_tag)::value;
        constexpr int NK = decltype(nk_tag)::value;   // K-steps per tap known at compile time (0 = runtime loop)
        constexpr bool PK = decltype(pk_tag)::value;  // K-packed 24-channel layer (NK == 2)
        // instruction descriptor: D=f32, A=B=bf16 (format 1) or fp16 (format 0), both K-major, N=BN, M=128
        const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(ka.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lbo_a = (uint32_t)ka.astride * 16, lbo_b = (uint32_t)ka.BN * 16;
        // Descriptors differ only in the 14-bit start-address field (bits 0-13 of the low word).
        const uint64_t adesc0 = make_desc(0, lbo_a, 128), bdesc0 = make_desc(0, lbo_b, 128);
        const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
        const uint32_t a_kstep = (2u * lbo_a) >> 4, b_kstep = (2u * lbo_b) >> 4;
        const int nk16 = (ka.debug & 2) ? 0 : ka.KC / 2;
        const uint32_t a_lo0 = (uint32_t)adesc0 + ((smem_u32(a_smem) & 0x3FFFFu) >> 4);
        const uint32_t b_lo0 = (uint32_t)bdesc0 + ((smem_u32(b_smem) & 0x3FFFFu) >> 4);
        const uint32_t a_stage_lo = (uint32_t)ka.a_stage_bytes >> 4, b_stage_lo = (uint32_t)ka.b_stage_bytes >> 4;
        const uint32_t bnc = (uint32_t)ka.BNC;
        const uint32_t res_shift = (uint32_t)(-ka.minoff);
        const uint32_t a_lbo_word = (uint32_t)adesc0;   // the K-chunk stride field of the low word (K-packed steps carry their own)
        const int ntaps = a.ntaps, NKB = ka.NKB, NA = ka.NA, NB = ka.NB, NACC = ka.ACC;
        const bool resident = ka.b_resident != 0;
        int sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
        uint32_t a_lo_stage = a_lo0, b_lo = b_lo0;
        uint32_t ph_act = 0, ph_res = 0;   // fused mode: per-stage phase bits of A_FULL / AR_FULL
        const int nkbt = (1 + ka.n_extra) * NKB;
        bool first = true;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          TRACE(1, 0, t);
          mbar_wait(T_EMPTY(acc), pacc ^ 1);   // epilogue has drained this accumulator stage
          TRACE(1, 1, t);
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(acc * acc_cols);
          uint32_t accum = 0;
          for (int kb = 0; kb < nkbt; ++kb) {
            const bool is_res = kb >= NKB;   // residual tile x identity weights
            uint32_t sh = is_res ? res_shift : sh0;
            if constexpr (FUSE) {
              if (is_res) { mbar_wait(AR_FULL(sa), (ph_res >> sa) & 1u); ph_res ^= 1u << sa; }
              else { mbar_wait(A_FULL(sa), (ph_act >> sa) & 1u); ph_act ^= 1u << sa; }
            } else {
              mbar_wait(A_FULL(sa), pa);
            }
            TRACE(1, 2, t);
            tc_fence_after();
            int ntap_kb = is_res ? 1 : ntaps;
            if constexpr (PK) {
              if (!is_res) {
                // K-packed 24-channel convolution: the (tap, chunk) items back to back, two per K-step.  Per pair of taps
                // (t, t+1) three K-steps:  (t: chunks 0,1)  ((t+1: chunk 0), (t: chunk 2))  (t+1: chunks 1,2)  -- the middle
                // one straddles the taps, stored in that order so its K-chunk stride 2*astride - d stays positive; an odd
                // last tap adds (chunks 0,1) and (chunk 2, zero chunk).  Start address and stride are uniform arithmetic
                // on (tap, dilation): no table, everything stays in uniform registers.  The weights are resident.
                if (first) {
                  int s2 = sb, p2 = pb;
                  for (int i = 0; i < ka.nstg; ++i) { mbar_wait(B_FULL(s2), p2); if (++s2 == NB) { s2 = 0; p2 ^= 1; } }
                  tc_fence_after();
                }
                const uint32_t ast = (uint32_t)ka.astride;
                const uint32_t dil = (uint32_t)(a.tap_off[1] - a.tap_off[0]);
                const uint32_t L1 = ast << 16, L2 = (2u * ast - dil) << 16;
                uint32_t base = a_lo_stage - a_lbo_word;   // start field (stage base + tap shift), stride field empty
                uint32_t blo = b_lo, af = accum;
                auto ks = [&](uint32_t alo) {
                  const uint64_t bd = ((uint64_t)b_hi << 32) | blo;
                  umma_bf16(d0, ((uint64_t)a_hi << 32) | alo, bd, idesc, af);
                  if (MS >= 2) umma_bf16(d0 + bnc, ((uint64_t)a_hi << 32) | (alo + 128u), bd, idesc, af);
                  if (MS == 4) {
                    umma_bf16(d0 + 2 * bnc, ((uint64_t)a_hi << 32) | (alo + 256u), bd, idesc, af);
                    umma_bf16(d0 + 3 * bnc, ((uint64_t)a_hi << 32) | (alo + 384u), bd, idesc, af);
                  }
                  blo += b_kstep;
                  af = 1;
                };
                int tp = 0;
                for (; tp + 1 < a.ntaps; tp += 2) {
                  ks(base + L1);
                  ks(base + dil + L2);
                  ks(base + ast + dil + L1);
                  base += 2u * dil;
                }
                if (tp < a.ntaps) {
                  ks(base + L1);
                  ks(base + 2u * ast + L1);
                }
                accum = 1;
                sb += ka.nstg;                                   // resident: NB == stages per tile (wraps here when no identity stage follows)
                b_lo += (uint32_t)ka.nstg * b_stage_lo;
                if (sb >= NB) { sb -= NB; pb ^= 1; b_lo = b_lo0 + (uint32_t)sb * b_stage_lo; }
                ntap_kb = 0;
              }
            }
            for (int tap = 0; tap < ntap_kb; ++tap) {
              if (!resident || first) {
                mbar_wait(B_FULL(sb), pb);
                tc_fence_after();
              }
              uint32_t alo = a_lo_stage + sh, blo = b_lo;
              uint32_t af = accum;
              auto kstep = [&]() {
                const uint64_t bd = ((uint64_t)b_hi << 32) | blo;
                umma_bf16(d0, ((uint64_t)a_hi << 32) | alo, bd, idesc, af);
                if (MS >= 2) umma_bf16(d0 + bnc, ((uint64_t)a_hi << 32) | (alo + 128u), bd, idesc, af);   // +128 rows
                if (MS == 4) {
                  umma_bf16(d0 + 2 * bnc, ((uint64_t)a_hi << 32) | (alo + 256u), bd, idesc, af);
                  umma_bf16(d0 + 3 * bnc, ((uint64_t)a_hi << 32) | (alo + 384u), bd, idesc, af);
                }
                alo += a_kstep; blo += b_kstep;
                af = 1;
              };
              if constexpr (NK > 0) {   // straight-line: no loop control between the MMAs of a tap
#pragma unroll
                for (int k16 = 0; k16 < NK; ++k16) kstep();
              } else {
#pragma unroll 2
                for (int k16 = 0; k16 < nk16; ++k16) kstep();
              }
              accum = 1;
              if (!resident) umma_commit(B_EMPTY(sb));
              b_lo += b_stage_lo;
              if (++sb == NB) { sb = 0; pb ^= 1; b_lo = b_lo0; }
              sh += sh_step;
            }
            umma_commit(A_EMPTY(sa));
            a_lo_stage += a_stage_lo;
            if (++sa == NA) { sa = 0; pa ^= 1; a_lo_stage = a_lo0; }
          }
          umma_commit(T_FULL(acc));
          TRACE(1, 3, t);
          if (++acc == NACC) { acc = 0; pacc ^= 1; }
          first = false;
        }
      };
      // specialised (MSUB, K-steps) pairs = the narrow AMP-block layers, where the issue loop is the limit
      using std::integral_constant;
      const int nk = (ka.debug & 2) ? -1 : ka.KC / 2;
      using no_pk = std::false_type;
      if (ka.packed && !FUSE) {   // K-packed 24-channel layer (nk == 2)
        if (ka.MSUB == 4) run(integral_constant<int, 4>{}, integral_constant<int, 2>{}, std::true_type{});
        else if (ka.MSUB == 2) run(integral_constant<int, 2>{}, integral_constant<int, 2>{}, std::true_type{});
        else run(integral_constant<int, 1>{}, integral_constant<int, 2>{}, std::true_type{});
      }
      else if (ka.MSUB == 4 && nk == 2) run(integral_constant<int, 4>{}, integral_constant<int, 2>{}, no_pk{});        // C = 24
      else if (ka.MSUB == 4 && nk == 3) run(integral_constant<int, 4>{}, integral_constant<int, 3>{}, no_pk{});   // C = 48
      else if (ka.MSUB == 2 && nk == 6) run(integral_constant<int, 2>{}, integral_constant<int, 6>{}, no_pk{});   // C = 96
      else if (ka.MSUB == 1 && nk == 4) run(integral_constant<int, 1>{}, integral_constant<int, 4>{}, no_pk{});   // C >= 192 (KC = 8)
      else if (ka.MSUB == 2 && nk == 4) run(integral_constant<int, 2>{}, integral_constant<int, 4>{}, no_pk{});
      else if (ka.MSUB == 4) run(integral_constant<int, 4>{}, integral_constant<int, 0>{}, no_pk{});
      else if (ka.MSUB == 2) run(integral_constant<int, 2>{}, integral_constant<int, 0>{}, no_pk{});
      else run(integral_constant<int, 1>{}, integral_constant<int, 0>{}, no_pk{});
    }
    __syncwarp();
  } else if (warp < 2 + epiw) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32*quarter ..
    const int half = (warp - 2) >> 2;        // the epiw/4 warps of a quarter alternate over work items
    constexpr int nsplit = epiw / 4;
    const int ngroups = ka.BNC >> 5;         // 32-column groups per accumulator
    const int nitems = ka.MSUB * ngroups;
    const int N = a.u * a.Cout;
    const bool tr = warp == 2 && lane == 0;
    const bool plain = a.u == 1;
    __nv_bfloat16* yg = reinterpret_cast<__nv_bfloat16*>(a.y);
    const __nv_bfloat16* rg = ka.res_mma ? nullptr : reinterpret_cast<const __nv_bfloat16*>(a.res);
    const bool accum_epi = a.accumulate && !ka.acc_mma;   // old output still to be added by the epilogue
    const bool simple = plain && use_sbias && rg == nullptr && !accum_epi;
    const size_t cs = (size_t)a.Ry * 8;      // elements between consecutive 8-channel chunks
    int acc = 0, pacc = 0;
    int t = blockIdx.x;
    TileRef cur = t < total_tiles ? decode(t) : TileRef{0, 0, 0};
    SegDesc so = a.seg_out[cur.b];
    int Lq = a.seg_in[cur.b].len + a.q_extra;
    while (t < total_tiles) {
      if (tr) TRACE(2, 0, t);
      const TileRef tl = cur;
      const SegDesc soc = so;
      const int Lqc = Lq;
      const int tn = t + gridDim.x;
      if (tn < total_tiles) {   // prefetch the next tile's coordinates before blocking
        cur = decode(tn);
        so = a.seg_out[cur.b];
        Lq = a.seg_in[cur.b].len + a.q_extra;
      }
      const int n0 = tl.nt * ka.BN;
      const float* biasp = a.bias ? a.bias + (size_t)tl.b * a.bias_bstride : nullptr;
      if (tr) TRACE(2, 1, t);
      mbar_wait(T_FULL(acc), pacc);
      if (tr) TRACE(2, 2, t);
      tc_fence_after();
      for (int item = half; item < nitems; item += nsplit) {
        const int sub = item / ngroups, grp = item - sub * ngroups;
        const int q = tl.q0 + sub * 128 + quarter * 32 + lane;
        const bool qok = q < Lqc && !(ka.debug & 1);
        uint32_t r[32];
        __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
        tmem_ld32_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + sub * ka.BNC + grp * 32), r);
        const int nbase = n0 + grp * 32;     // first GEMM column of this group
        if constexpr (F32IO) {
          // fp32 tensor-core mode: fp32 bias / residual / old output / store, plain and transposed layers in one path
          float* yf = reinterpret_cast<float*>(a.y);
          const float* rf = reinterpret_cast<const float*>(a.res);
          tmem_ld_wait();
          int phase = plain ? 0 : nbase / a.Cout, co = plain ? nbase : nbase - phase * a.Cout;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int orow = plain ? q : q * a.u + phase - a.p;
            const bool okk = qok && nbase + 8 * u < N && orow >= 0 && orow < soc.len;
            if (okk) {
              const size_t off = ((size_t)(co >> 3) * a.Ry + soc.off + orow) * 8;
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[8 * u + j]), a.acc_scale, biasp ? __ldg(biasp + co + j) : 0.f);
              if (rf) {
                const float4 r0 = *reinterpret_cast<const float4*>(rf + off), r1 = *reinterpret_cast<const float4*>(rf + off + 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= a.out_scale;
              if (a.accumulate) {
                const float4 o0 = *reinterpret_cast<const float4*>(yf + off), o1 = *reinterpret_cast<const float4*>(yf + off + 4);
                v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w; v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
              }
              *reinterpret_cast<float4*>(yf + off) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(yf + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            co += 8;
            if (!plain && co >= a.Cout) { co -= a.Cout; ++phase; }
          }
        } else
        if (simple) {
          // fast path (no residual / accumulate reads, batch-independent bias in shared memory): the
          // common case of the AMP-block convolutions -- bias, scale, convert, four 16-byte stores
          const bool valid = qok && q < soc.len;
          __nv_bfloat16* yp = yg + ((size_t)(nbase >> 3) * a.Ry + soc.off + q) * 8;
          const float4* sb4 = reinterpret_cast<const float4*>(sbias + nbase);
          float4 bv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = sb4[j];
          const float osc = a.out_scale;
          tmem_ld_wait();
          if (osc == 1.f) {   // the common case: no scaling pass
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float v[8];
              v[0] = __uint_as_float(r[8 * u + 0]) + bv[2 * u].x; v[1] = __uint_as_float(r[8 * u + 1]) + bv[2 * u].y;
              v[2] = __uint_as_float(r[8 * u + 2]) + bv[2 * u].z; v[3] = __uint_as_float(r[8 * u + 3]) + bv[2 * u].w;
              v[4] = __uint_as_float(r[8 * u + 4]) + bv[2 * u + 1].x; v[5] = __uint_as_float(r[8 * u + 5]) + bv[2 * u + 1].y;
              v[6] = __uint_as_float(r[8 * u + 6]) + bv[2 * u + 1].z; v[7] = __uint_as_float(r[8 * u + 7]) + bv[2 * u + 1].w;
              if (valid && nbase + 8 * u < N) *reinterpret_cast<uint4*>(yp + u * cs) = pack8<F16>(v);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float v[8];
              v[0] = (__uint_as_float(r[8 * u + 0]) + bv[2 * u].x) * osc; v[1] = (__uint_as_float(r[8 * u + 1]) + bv[2 * u].y) * osc;
              v[2] = (__uint_as_float(r[8 * u + 2]) + bv[2 * u].z) * osc; v[3] = (__uint_as_float(r[8 * u + 3]) + bv[2 * u].w) * osc;
              v[4] = (__uint_as_float(r[8 * u + 4]) + bv[2 * u + 1].x) * osc; v[5] = (__uint_as_float(r[8 * u + 5]) + bv[2 * u + 1].y) * osc;
              v[6] = (__uint_as_float(r[8 * u + 6]) + bv[2 * u + 1].z) * osc; v[7] = (__uint_as_float(r[8 * u + 7]) + bv[2 * u + 1].w) * osc;
              if (valid && nbase + 8 * u < N) *reinterpret_cast<uint4*>(yp + u * cs) = pack8<F16>(v);
            }
          }
        } else if (plain) {
          // plain convolution: output row == q, column == channel; chunk c lives cs elements further.
          // Everything that does not depend on the accumulator (bias, residual, old output) is issued
          // before waiting for the TMEM load, and the four 8-channel chunks are handled branch-free.
          const bool valid = qok && q < soc.len;
          const size_t base = ((size_t)(nbase >> 3) * a.Ry + soc.off + q) * 8;
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) ok[u] = valid && nbase + 8 * u < N;
          float4 bv[8];
          if (use_sbias) {
            const float4* sb4 = reinterpret_cast<const float4*>(sbias + nbase);
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = sb4[j];
          } else if (biasp) {   // per-segment bias (speaker conditioning) or a very wide layer: same address for the whole warp
            const float4* gb4 = reinterpret_cast<const float4*>(biasp + nbase);
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = nbase + 4 * j < N ? __ldg(gb4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // per-channel residual factors (pre-scaled residual stream, see fold_activation_scales): read per 8-channel unit
          // where they are used (the same addresses for the whole warp, L1 hits) -- holding all 32 in registers made the
          // kernel spill
          const bool rscaled = rg != nullptr && a.res_scale != nullptr;
          const float4* rs4 = reinterpret_cast<const float4*>(a.res_scale + nbase);
          uint4 resv[4], oldv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (rg && ok[u]) resv[u] = *reinterpret_cast<const uint4*>(rg + base + u * cs);
            if (accum_epi && ok[u]) oldv[u] = *reinterpret_cast<const uint4*>(yg + base + u * cs);
          }
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8];
            v[0] = __uint_as_float(r[8 * u + 0]) + bv[2 * u].x; v[1] = __uint_as_float(r[8 * u + 1]) + bv[2 * u].y;
            v[2] = __uint_as_float(r[8 * u + 2]) + bv[2 * u].z; v[3] = __uint_as_float(r[8 * u + 3]) + bv[2 * u].w;
            v[4] = __uint_as_float(r[8 * u + 4]) + bv[2 * u + 1].x; v[5] = __uint_as_float(r[8 * u + 5]) + bv[2 * u + 1].y;
            v[6] = __uint_as_float(r[8 * u + 6]) + bv[2 * u + 1].z; v[7] = __uint_as_float(r[8 * u + 7]) + bv[2 * u + 1].w;
            if (rg && ok[u]) {
              if (rscaled) unpack_fma<F16>(resv[u], v, __ldg(rs4 + 2 * u), __ldg(rs4 + 2 * u + 1));
              else unpack_add<F16>(resv[u], v);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= a.out_scale;
            if (accum_epi && ok[u]) unpack_add<F16>(oldv[u], v);
            if (ok[u]) *reinterpret_cast<uint4*>(yg + base + u * cs) = pack8<F16>(v);
          }
        } else {
          // transposed convolution: column n = (phase, channel), output row = q*u + phase - p.
          // (phase, channel) of the four 8-column chunks advance incrementally (Cout is a multiple of 8);
          // bias / residual / old rows are requested before waiting for the TMEM load.
          int phase = nbase / a.Cout, co = nbase - phase * a.Cout;
          bool ok[4];
          size_t off[4];
          float4 b0[4], b1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int orow = q * a.u + phase - a.p;
            ok[u] = qok && nbase + 8 * u < N && orow >= 0 && orow < soc.len;
            off[u] = ((size_t)(co >> 3) * a.Ry + soc.off + orow) * 8;
            b0[u] = b1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (biasp && nbase + 8 * u < N) {
              b0[u] = __ldg(reinterpret_cast<const float4*>(biasp + co));
              b1[u] = __ldg(reinterpret_cast<const float4*>(biasp + co + 4));
            }
            co += 8;
            if (co >= a.Cout) { co -= a.Cout; ++phase; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8];
            v[0] = __uint_as_float(r[8 * u + 0]) + b0[u].x; v[1] = __uint_as_float(r[8 * u + 1]) + b0[u].y;
            v[2] = __uint_as_float(r[8 * u + 2]) + b0[u].z; v[3] = __uint_as_float(r[8 * u + 3]) + b0[u].w;
            v[4] = __uint_as_float(r[8 * u + 4]) + b1[u].x; v[5] = __uint_as_float(r[8 * u + 5]) + b1[u].y;
            v[6] = __uint_as_float(r[8 * u + 6]) + b1[u].z; v[7] = __uint_as_float(r[8 * u + 7]) + b1[u].w;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= a.out_scale;
            if (ok[u]) *reinterpret_cast<uint4*>(yg + off[u]) = pack8<F16>(v);   // (no residual / accumulate for transposed layers)
          }
        }
      }
      // all of this warp's tcgen05.ld for the stage have completed (wait::ld above): release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(T_EMPTY(acc));
      if (tr) TRACE(2, 3, t);
      if (++acc == ka.ACC) { acc = 0; pacc ^= 1; }
      t = tn;
    }
  }
  else if constexpr (FUSE) {
    // ===================== activation warps (fused mode) =====================
    // Raw tile (R stage) -> Activation1d -> activated tile (A stage, UMMA layout), with both FIR filters as
    // warp-level MMAs (the scheme of csrc/bvg_act3.cu): the 16 MMA rows are two work items of 8 channels
    // (a chunk x a range of act_jr 8-row column tiles), the MMA columns are time; ldmatrix.trans reads the
    // raw rows, stmatrix.trans writes activated bf16 rows straight into the A operand slab.
    const int aw = warp - 2 - epiw;
    const int g = lane >> 2, t4 = lane & 3;
    const int rrows = ka.rstride;
    // constant tap fragments: up-FIR B[k][n] = 2 f[n + 11 - 2k] (bf16), down-FIR B_d[k][n] = f[16 d + k - 2n + 5] (fp16)
    uint32_t gup[2], fdn[3][2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k0 = 2 * t4 + 8 * half;
      gup[half] = pk_bf16(2.f * fir_tap(g + 11 - 2 * k0), 2.f * fir_tap(g + 11 - 2 * (k0 + 1)));
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int off = 16 * (d - 1) + 5 - 2 * g;
        fdn[d][half] = pk_f16(fir_tap(k0 + off), fir_tap(k0 + 1 + off));
      }
    }
    int sr = 0, pr = 0, sa = 0, pa = 0;
    int t = blockIdx.x;
    TileRef cur = t < total_tiles ? decode(t) : TileRef{0, 0, 0};
    int Lseg = a.seg_in[cur.b].len;
    while (t < total_tiles) {
      const TileRef tl = cur;
      const int L = Lseg;
      const int tn = t + gridDim.x;
      if (tn < total_tiles) { cur = decode(tn); Lseg = a.seg_in[cur.b].len; }
      const int tA0 = tl.q0 + ka.minoff;      // segment time of A-slab row 0 (raw slab row r = time tA0 - 8 + r)
      const bool tr = aw == 0 && lane == 0;
      for (int kb = 0; kb < ka.NKB; ++kb) {
        const int kcl = (kb == ka.NKB - 1) ? ka.kc_last_load : ka.KC;
        if (tr) TRACE(3, 0, t);
        mbar_wait(R_FULL(sr), pr);
        if (tr) TRACE(3, 1, t);
        mbar_wait(A_EMPTY(sa), pa ^ 1);
        if (tr) TRACE(3, 2, t);
        const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(r_smem + (size_t)sr * ka.r_stage_bytes);
        __nv_bfloat16* abase = reinterpret_cast<__nv_bfloat16*>(a_smem + (size_t)sa * ka.a_stage_bytes);
        const uint32_t r_s = smem_u32(rbase), a_s = smem_u32(abase);
        const int nitems = kcl * ka.act_ngc;
        for (int unit = aw; 2 * unit < nitems; unit += NACT) {
          // the two work items of this warp (the second one repeats the first when the count is odd)
          int ic[2], iJ0[2];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int it = 2 * unit + s2 < nitems ? 2 * unit + s2 : 2 * unit;
            ic[s2] = it / ka.act_ngc;
            iJ0[s2] = (it - ic[s2] * ka.act_ngc) * ka.act_jr;
          }
          float a2[2], hh[2];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int ch = (kb * ka.KC + ic[s2]) * 8 + g;
            a2[s2] = 2.f * ka.act_alpha[ch];
            hh[s2] = 0.5f * ka.act_inv_beta[ch];
          }
          // ldmatrix row address of this lane: matrices 0..3 = (item 0, rows +0..7), (item 1, +0..7), (item 0, +8..15), (item 1, +8..15)
          const int sl = (lane >> 3) & 1;
          const uint32_t ld_base = r_s + (uint32_t)(ic[sl] * ka.rstride + 8 * iJ0[sl] + (lane >> 4) * 8 + (lane & 7)) * 16;
          // stmatrix row address: matrices 0 / 1 = item 0 / 1
          const uint32_t st_base = a_s + (uint32_t)(ic[sl] * ka.astride + 8 * iJ0[sl] + (lane & 7)) * 16;
          // one up-FIR column tile j (relative to the item's first column tile): 8 activated samples of both items
          auto up_tile = [&](int j, uint32_t& p0, uint32_t& p1) {
            uint32_t xa[4];
            ldsm_x4_trans(ld_base + (uint32_t)(4 * j + 5) * 16, xa);   // raw rows of local time 4j-3 .. 4j+12 (+8 margin)
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            wmma_bf16(c, xa, gup[0], gup[1]);
            const float s0 = fmaf(-hh[0], __cosf(a2[0] * c[0]), c[0]), s1 = fmaf(-hh[0], __cosf(a2[0] * c[1]), c[1]);
            const float s2 = fmaf(-hh[1], __cosf(a2[1] * c[2]), c[2]), s3 = fmaf(-hh[1], __cosf(a2[1] * c[3]), c[3]);
            p0 = pk_f16(s0, s1);
            p1 = pk_f16(s2, s3);
          };
          uint32_t ap[4], ac[4], an[4];
          ap[0] = ap[1] = 0u;   // (zero taps)
          up_tile(-1, ap[2], ap[3]);
          up_tile(0, ac[0], ac[1]);
          up_tile(1, ac[2], ac[3]);
#pragma unroll 1
          for (int J = 0; J < ka.act_jr; ++J) {
            up_tile(2 * J + 2, an[0], an[1]);
            up_tile(2 * J + 3, an[2], an[3]);
            float c[4] = {hh[0], hh[0], hh[1], hh[1]};
            wmma_f16(c, ap, fdn[0][0], fdn[0][1]);
            wmma_f16(c, ac, fdn[1][0], fdn[1][1]);
            wmma_f16(c, an, fdn[2][0], fdn[2][1]);
            stsm_x2_trans(st_base + (uint32_t)(8 * J) * 16, pk_bf16(c[0], c[1]), pk_bf16(c[2], c[3]));
#pragma unroll
            for (int i = 0; i < 4; ++i) { ap[i] = ac[i]; ac[i] = an[i]; }
          }
          // rows outside the segment are the conv's "same" zero padding of the ACTIVATED signal; the three rows next to
          // each segment end see the replicate padding of the activated 2x signal: exact form.  lane = (row, channel pair)
          __syncwarp();
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            if (s2 == 1 && 2 * unit + 1 >= nitems) break;
            const int t_first = tA0 + 8 * iJ0[s2], t_end = t_first + 8 * ka.act_jr;
            if (t_first >= 3 && t_end <= L - 3) continue;
            const int ch = (kb * ka.KC + ic[s2]) * 8 + 2 * t4;
            const float fa0 = 2.f * ka.act_alpha[ch], fa1 = 2.f * ka.act_alpha[ch + 1];
            const float fh0 = 0.5f * ka.act_inv_beta[ch], fh1 = 0.5f * ka.act_inv_beta[ch + 1];
            const __nv_bfloat16* rcol = rbase + (size_t)(ic[s2] * ka.rstride) * 8 + 2 * t4;
            __nv_bfloat16* ycol = abase + (size_t)(ic[s2] * ka.astride + 8 * iJ0[s2]) * 8 + 2 * t4;
#pragma unroll 1
            for (int J = 0; J < ka.act_jr; ++J) {
              const int row = 8 * J + g, ts = t_first + row;
              if (ts < 0 || ts >= L) *reinterpret_cast<uint32_t*>(ycol + row * 8) = 0u;
              else if (ts < 3 || ts >= L - 3)
                actcore::stpair(ycol + row * 8, actcore::exact_clamped(rcol, tA0 - 8, rrows, ts, L, fa0, fa1, fh0, fh1));
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic st.shared -> tcgen05 (async proxy) reads
        __syncwarp();
        if (lane == 0) { mbar_arrive(A_FULL(sa)); mbar_arrive(R_EMPTY(sr)); }
        if (tr) TRACE(3, 3, t);
        if (++sr == ka.NR) { sr = 0; pr ^= 1; }
        if (++sa == ka.NA) { sa = 0; pa ^= 1; }
      }
      for (int kb = 0; kb < ka.n_extra * ka.NKB; ++kb)   // the residual / old-output k-blocks use the next stages of the A ring
        if (++sa == ka.NA) { sa = 0; pa ^= 1; }
      t = tn;
    }
  }
  {
    // the tracing lane of each role flushes its ring (only one lane per role has tr_n > 0)
    const int role = warp < 3 ? warp : 3;
    __syncwarp();
    if (ka.trace && blockIdx.x == 0 && tr_n > 0) {
      ka.trace[role] = (unsigned long long)tr_n;
      for (int i = 0; i < tr_n; ++i) ka.trace[4 + role * 1024 + i] = tr_smem[role * 1024 + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)ka.tmem_cols);
  }
  if (a.pdl_mode == 2) pdl_wait();   // (see ConvArgs::pdl_mode)
}

// fp32 [tap][Cin][N]  ->  bf16 images [ntile][kb][tap][chunk KC][n BN][8]
template <typename T>
__global__ void repack_umma_kernel(const float* __restrict__ wt, T* __restrict__ img, int ntaps, int Cin,
                                   int N, int KC, int NKB, int BN, int NT, const float* __restrict__ out_scale,
                                   const float* __restrict__ in_scale) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)NT * NKB * ntaps * KC * BN * 8;
  if (idx >= total) return;
  int e = idx & 7;
  size_t r = idx >> 3;
  int nn = r % BN; r /= BN;
  int c = r % KC; r /= KC;
  int tap = r % ntaps; r /= ntaps;
  int kb = r % NKB;
  int nt = r / NKB;
  int ci = (kb * KC + c) * 8 + e, n = nt * BN + nn;
  float v = (ci < Cin && n < N) ? wt[((size_t)tap * Cin + ci) * N + n] : 0.f;
  // per-channel scales folded into the image in fp32, before the one rounding to the storage type (plain Conv1d layers
  // only: n is the output channel there)
  if (out_scale && n < N) v *= out_scale[n];
  if (in_scale && ci < Cin) v *= in_scale[ci];
  img[idx] = from_f32<T>(v);
}

// identity images [kb][chunk KC][n BN][8] appended after the conv images of a square (Cin == N), single-n-tile layer
// K-packed image of a 24-channel layer: [stage][slot 0..3][n BN][8]; slot = (K-step within the stage, half); K-step s holds
// the items 2s and 2s+1 of the (tap, chunk) sequence, a pair that straddles two taps in swapped order (see the issuer's table)
template <typename T>
__global__ void repack_umma_packed_kernel(const float* __restrict__ wt, T* __restrict__ img, int ntaps, int Cin, int N,
                                          int BN, int nstg, const float* __restrict__ out_scale, const float* __restrict__ in_scale) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)nstg * 4 * BN * 8;
  if (idx >= total) return;
  const int e = idx & 7;
  size_t r = idx >> 3;
  const int nn = r % BN; r /= BN;
  const int slot = r % 4;
  const int stage = r / 4;
  const int sidx = 2 * stage + (slot >> 1), half = slot & 1, nitems = 3 * ntaps;
  const int i0 = 2 * sidx, i1 = 2 * sidx + 1;
  const bool swapped = i1 < nitems && (i0 / 3) != (i1 / 3);
  const int item = half == 0 ? (swapped ? i1 : i0) : (swapped ? i0 : i1);
  float v = 0.f;
  if (item < nitems && nn < N) {
    const int tap = item / 3, ci = (item - 3 * tap) * 8 + e;
    v = wt[((size_t)tap * Cin + ci) * N + nn];
    if (out_scale) v *= out_scale[nn];
    if (in_scale) v *= in_scale[ci];
  }
  img[idx] = from_f32<T>(v);
}
template <typename T>
__global__ void identity_umma_kernel(T* __restrict__ img, int N, int KC, int NKB, int BN, float value,
                                     const float* __restrict__ diag /* optional per-channel factor */) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)NKB * KC * BN * 8;
  if (idx >= total) return;
  int e = idx & 7;
  size_t r = idx >> 3;
  int nn = r % BN; r /= BN;
  int c = r % KC;
  int kb = r / KC;
  int ci = (kb * KC + c) * 8 + e;
  img[idx] = from_f32<T>((ci == nn && nn < N) ? (diag ? value * diag[nn] : value) : 0.f);
}

bool has_identity(const UmmaTiling& t, int Cin, int N) { return t.ok && t.NT == 1 && Cin == N; }

bool use_res_mma(const ConvArgs& a, const UmmaTiling& t) {
  static const int maxc = env_int("BVG_RES_MMA_MAXC", 96);
  return a.res && a.u == 1 && has_identity(t, a.Cin, a.Cout) && a.Cin <= maxc;
}
// the second identity set holds acc_img_scale * I; usable when out_scale * acc_img_scale == 1
bool use_acc_mma(const ConvArgs& a, const UmmaTiling& t) {
  static const int maxc = env_int("BVG_RES_MMA_MAXC", 96);
  const float p = a.out_scale * a.acc_img_scale;
  return a.accumulate && a.u == 1 && has_identity(t, a.Cin, a.Cout) && a.Cin <= maxc && p > 0.999999f && p < 1.000001f;
}

void tap_range(const ConvArgs& a, int& mn, int& mx) {
  mn = mx = a.tap_off[0];
  for (int j = 1; j < a.ntaps; ++j) { mn = a.tap_off[j] < mn ? a.tap_off[j] : mn; mx = a.tap_off[j] > mx ? a.tap_off[j] : mx; }
}

// Launch-time configuration for a given number of 128-row sub-tiles per tile.
bool configure(const ConvArgs& a, int msub, UmmaKernelArgs& ka, size_t& smem_bytes) {
  const int N = a.u * a.Cout;
  const UmmaTiling t = make_tiling(a.ntaps, a.Cin, N, a.bn_small != 0, a.k_packed != 0);
  if (!t.ok) return false;
  int mn, mx;
  tap_range(a, mn, mx);
  if (mx - mn > MAXSPAN || -mn > BVG_GUARD || mx > BVG_GUARD) return false;
  for (int j = 2; j < a.ntaps; ++j)   // the issue loop advances the tap shift by a constant step
    if (a.tap_off[j] - a.tap_off[j - 1] != a.tap_off[1] - a.tap_off[0]) return false;
  ka.c = a;
  ka.KC = t.KC; ka.NKB = t.NKB; ka.BN = t.BN; ka.BNC = t.BNC; ka.NT = t.NT;
  ka.packed = t.packed; ka.nks = t.nks; ka.nstg = t.nstg;
  ka.minoff = mn; ka.span = mx - mn;
  ka.MSUB = msub;
  ka.ACC = 512 / (ka.MSUB * ka.BNC);
  if (ka.ACC > 4) ka.ACC = 4;
  if (ka.ACC < 1) return false;
  ka.tmem_cols = 32;
  while (ka.tmem_cols < ka.ACC * ka.MSUB * ka.BNC) ka.tmem_cols <<= 1;
  ka.astride = 128 * ka.MSUB + MAXSPAN + 6;
  ka.a_stage_bytes = t.KC * ka.astride * 16;
  ka.b_stage_bytes = t.KC * t.BN * 16;
  ka.kc_last_load = a.Cin / 8 - (t.NKB - 1) * t.KC;
  static const int debug_env = env_int("BVG_CONV_DEBUG", 0);
  ka.debug = debug_env;
  ka.trace = nullptr;
  ka.res_mma = use_res_mma(a, t) ? 1 : 0;
  ka.acc_mma = use_acc_mma(a, t) ? 1 : 0;
  ka.n_extra = ka.res_mma + ka.acc_mma;
  // pipeline depths within the smem budget
  const int total_b = t.NKB * (t.nstg + ka.n_extra);
  ka.NA = t.NKB > 1 ? 2 : 3;
  ka.b_resident = 0;
  static const int allow_resident = env_int("BVG_CONV_RESIDENT", 1);
  if (allow_resident && t.NT == 1 && total_b <= MAX_STAGES &&
      (size_t)2 * ka.a_stage_bytes + (size_t)total_b * ka.b_stage_bytes <= (size_t)SMEM_BUDGET) {
    ka.b_resident = 1;
    ka.NB = total_b;
    int na = (int)((SMEM_BUDGET - (size_t)total_b * ka.b_stage_bytes) / ka.a_stage_bytes);
    ka.NA = na > 4 ? 4 : na;   // >= 2 by the test above
  } else {
    if ((size_t)ka.NA * ka.a_stage_bytes + 2 * (size_t)ka.b_stage_bytes > (size_t)SMEM_BUDGET) ka.NA = 2;
    if ((size_t)ka.NA * ka.a_stage_bytes + 2 * (size_t)ka.b_stage_bytes > (size_t)SMEM_BUDGET) return false;
    int nb = (SMEM_BUDGET - ka.NA * ka.a_stage_bytes) / ka.b_stage_bytes;
    if (nb > 8) nb = 8;
    if (nb > total_b) nb = total_b;
    if (nb < 1) return false;
    ka.NB = nb;
  }
  if (ka.packed && !ka.b_resident) return false;   // the K-packed issue path assumes resident weights
  smem_bytes = (size_t)ka.NA * ka.a_stage_bytes + (size_t)ka.NB * ka.b_stage_bytes +
               8 * (size_t)(2 * ka.NA + 2 * ka.NB + 2 * ka.ACC) + 192 + 4 * SBIAS_MAX;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one persistent CTA per SM (it owns the TMEM)
  return smem_bytes <= (size_t)SMEM_MAX;
}

// Fused (Activation1d in the producer stage) configuration.  Needs the whole N in one CTA tile.
bool configure_fused(const ConvArgs& a, int msub, UmmaKernelArgs& ka, size_t& smem_bytes) {
  if (!a.act_alpha || !a.act_inv_beta || a.u != 1) return false;
  const UmmaTiling t = make_tiling(a.ntaps, a.Cin, a.Cout);
  if (!t.ok || t.NT != 1) return false;
  int mn, mx;
  tap_range(a, mn, mx);
  if (mx - mn > MAXSPAN || -mn > BVG_GUARD || mx > BVG_GUARD) return false;
  ka.c = a;
  ka.KC = t.KC; ka.NKB = t.NKB; ka.BN = t.BN; ka.BNC = t.BNC; ka.NT = t.NT;
  ka.packed = 0; ka.nks = 0; ka.nstg = a.ntaps;
  ka.minoff = mn; ka.span = mx - mn;
  ka.MSUB = msub;
  ka.ACC = 512 / (ka.MSUB * ka.BNC);
  if (ka.ACC > 4) ka.ACC = 4;
  if (ka.ACC < 1) return false;
  ka.tmem_cols = 32;
  while (ka.tmem_cols < ka.ACC * ka.MSUB * ka.BNC) ka.tmem_cols <<= 1;
  const int arows = 128 * msub + ka.span;
  // activation work items: act_ngc row ranges per chunk, two items per warp pass, about one pass per tile
  const int kc_items = t.KC;
  ka.act_ngc = 2 * NACT / kc_items;
  if (ka.act_ngc < 1) ka.act_ngc = 1;
  const int njt = (arows + 7) / 8;
  ka.act_jr = (njt + ka.act_ngc - 1) / ka.act_ngc;
  ka.nblk = ka.act_ngc * ka.act_jr;
  ka.astride = 8 * ka.nblk;
  ka.rstride = ka.astride + 32;          // local time -8 .. astride + 23
  if (ka.rstride + 8 > BVG_TAIL_SLACK + BVG_GUARD) return false;   // raw loads past a segment end stay in bounds
  ka.a_stage_bytes = t.KC * ka.astride * 16;
  ka.r_stage_bytes = t.KC * ka.rstride * 16;
  ka.b_stage_bytes = t.KC * t.BN * 16;
  ka.kc_last_load = a.Cin / 8 - (t.NKB - 1) * t.KC;
  static const int debug_env = env_int("BVG_CONV_DEBUG", 0);
  ka.debug = debug_env;
  ka.trace = nullptr;
  ka.act_alpha = a.act_alpha; ka.act_inv_beta = a.act_inv_beta;
  ka.NA = 2; ka.NR = 2;
  ka.res_mma = use_res_mma(a, t) ? 1 : 0;
  ka.acc_mma = use_acc_mma(a, t) ? 1 : 0;
  ka.n_extra = ka.res_mma + ka.acc_mma;
  const size_t fixed = (size_t)ka.NA * ka.a_stage_bytes + (size_t)ka.NR * ka.r_stage_bytes;
  const int total_b = t.NKB * (a.ntaps + ka.n_extra);
  ka.b_resident = 0;
  if (fixed + (size_t)ka.b_stage_bytes > (size_t)SMEM_BUDGET) return false;
  if (total_b <= MAX_STAGES && fixed + (size_t)total_b * ka.b_stage_bytes <= (size_t)SMEM_BUDGET) {
    ka.b_resident = 1;
    ka.NB = total_b;
  } else {
    int nb = (int)(((size_t)SMEM_BUDGET - fixed) / ka.b_stage_bytes);
    if (nb > 8) nb = 8;
    if (nb > total_b) nb = total_b;
    if (nb < 2 && total_b >= 2) return false;
    ka.NB = nb;
  }
  smem_bytes = fixed + (size_t)ka.NB * ka.b_stage_bytes + 8 * (size_t)(3 * ka.NA + 2 * ka.NB + 2 * ka.ACC + 2 * ka.NR) + 192 + 4 * SBIAS_MAX;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  return smem_bytes <= (size_t)SMEM_MAX;
}

}  // namespace

// fp32 tensor-core mode: W' [tap][3 Cin][N] = [2^-11 W_hi; W_lo; W_hi] with W_hi = fp16(S W), W_lo = fp16(S W - W_hi) (all
// exactly representable, so the fp16 repack that follows does not round again).  The kernel reads the input as the three
// channel blocks [2^11 x_lo | x_hi | x_hi] (split_f32), so the GEMM delivers S (x_lo W_hi + x_hi W_lo + x_hi W_hi): every
// product is exact in the fp32 accumulator and only x_lo W_lo (2^-22 relative) is dropped.  S is a power of two that puts
// the layer's largest weight in [2^13, 2^14): W_lo then stays a normal fp16 number for weights down to 2^-13 of the largest.
// Block order: the tensor core truncates its fp32 accumulator once per MMA (measured: 2.6e-8 x |acc| per MMA of the chain,
// linear in the chain length), so the two correction blocks run first, while the accumulator is 2^-11 of its final size,
// and only the W_hi x_hi third of the chain pays the truncation.
__global__ void split3_weights_kernel(const float* __restrict__ w, float* __restrict__ w3, int ntaps, int Cin, int N, float scale) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)ntaps * Cin * N;
  if (idx >= total) return;
  const int n = idx % N;
  size_t r = idx / N;
  const int ci = r % Cin, tap = r / Cin;
  const float v = w[idx] * scale;
  const float hi = f16_round_sat(v);
  const float lo = f16_round_sat(v - hi);
  float* base = w3 + (size_t)tap * 3 * Cin * N;
  base[(size_t)ci * N + n] = f16_round_sat(hi * (1.f / BVG_SPLIT_LO_SCALE));
  base[(size_t)(Cin + ci) * N + n] = lo;
  base[(size_t)(2 * Cin + ci) * N + n] = hi;
}
cudaError_t launch_split3_weights(const float* w_tap_major, float* w3, int ntaps, int Cin, int N, float scale, cudaStream_t s) {
  size_t total = (size_t)ntaps * Cin * N;
  split3_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w_tap_major, w3, ntaps, Cin, N, scale);
  return cudaGetLastError();
}
float split3_weight_scale(float absmax) {
  if (!(absmax > 0.f) || !std::isfinite(absmax)) return 1.f;
  int e = 0;
  std::frexp(absmax, &e);               // absmax = m 2^e, m in [0.5, 1)
  return std::ldexp(1.f, 14 - e);       // S absmax in [2^13, 2^14)
}

bool umma_k_packed_default(int Cin, int N) {
  static const int on = [] {
    const char* f = getenv("BVG_FUSE_ACT");   // the experimental fused kernel reads the unpacked image
    if (f && atoi(f)) return 0;
    return env_int("BVG_CONV_PACKK", 1);
  }();
  return on && Cin == 24 && N <= 256;
}

size_t umma_weight_image_bytes(int ntaps, int Cin, int N, bool small, bool packed) {
  UmmaTiling t = make_tiling(ntaps, Cin, N, small, packed);
  if (!t.ok) return 0;
  return (size_t)(t.NT * t.NKB * t.nstg + (has_identity(t, Cin, N) ? 2 * t.NKB : 0)) * t.KC * t.BN * 16;
}

template <typename T>
static void repack_umma_t(const float* wp_tap_major, T* img, const UmmaTiling& t, int ntaps, int Cin, int N, float acc_img_scale,
                          const float* out_scale, const float* in_scale, const float* res_diag, cudaStream_t s) {
  size_t total = (size_t)t.NT * t.NKB * t.nstg * t.KC * t.BN * 8;
  if (t.packed)
    repack_umma_packed_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(wp_tap_major, img, ntaps, Cin, N, t.BN, t.nstg, out_scale,
                                                                                in_scale);
  else
  repack_umma_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(wp_tap_major, img, ntaps, Cin, N, t.KC, t.NKB, t.BN, t.NT,
                                                                       out_scale, in_scale);
  if (has_identity(t, Cin, N)) {
    const size_t itotal = (size_t)t.NKB * t.KC * t.BN * 8;
    // set 0: I (residual), set 1: acc_img_scale * I (old output of an accumulating layer; exact for small integers)
    identity_umma_kernel<T><<<(unsigned)((itotal + 255) / 256), 256, 0, s>>>(img + total, N, t.KC, t.NKB, t.BN, 1.f, res_diag);
    identity_umma_kernel<T><<<(unsigned)((itotal + 255) / 256), 256, 0, s>>>(img + total + itotal, N, t.KC, t.NKB, t.BN, acc_img_scale, nullptr);
  }
}

cudaError_t launch_repack_umma(const float* wp_tap_major, void* img, int dtype, int ntaps, int Cin, int N, float acc_img_scale,
                               bool small, cudaStream_t s, const float* out_scale, const float* in_scale, const float* res_diag,
                               bool packed) {
  UmmaTiling t = make_tiling(ntaps, Cin, N, small, packed);
  if (!t.ok || (dtype != 1 && dtype != 2)) return cudaErrorInvalidValue;
  if (dtype == 1) repack_umma_t(wp_tap_major, (__nv_bfloat16*)img, t, ntaps, Cin, N, acc_img_scale, out_scale, in_scale, res_diag, s);
  else repack_umma_t(wp_tap_major, (__half*)img, t, ntaps, Cin, N, acc_img_scale, out_scale, in_scale, res_diag, s);
  return cudaGetLastError();
}

// Sub-tiles per tile: as many as still leave >= 2 TMEM accumulator stages (epilogue/MMA overlap) and
// fit the shared-memory budget; small problems keep small tiles so all SMs get work.
int conv_umma_default_msub(const ConvArgs& a) {
  static const int forced = env_int("BVG_CONV_MSUB", 0);
  if (a.bn_small) return 1;   // the small-batch variant exists to maximise the number of CTAs
  const UmmaTiling t = make_tiling(a.ntaps, a.Cin, a.u * a.Cout);
  if (!t.ok) return 1;
  for (int msub = 4; msub >= 2; msub >>= 1) {
    if (forced && msub > forced) continue;
    if (a.max_q <= 128 * (msub / 2)) continue;
    if (msub * t.BNC * 2 > 512) continue;
    UmmaKernelArgs ka{};
    size_t smem;
    if (configure(a, msub, ka, smem)) return msub;
  }
  return 1;
}

// Sub-tiles per tile for the fused kernel, 0 if this layer / geometry cannot be fused.
int conv_umma_fused_msub(const ConvArgs& a, bool force) {
  // Opt-in (BVG_FUSE_ACT=1): measured slower than the separate passes on B200 in round 1, because the
  // activation is FP32-pipe bound and 10 activation warps per SM cannot outrun the stand-alone kernel.
  static const int enabled = env_int("BVG_FUSE_ACT", 0);
  static const int forced = env_int("BVG_CONV_MSUB", 0);
  if (!(enabled || force) || !a.act_alpha || a.dtype == 2 || a.bn_small || a.k_packed || a.f32io) return 0;   // the fused activation warps are bf16 only
  const UmmaTiling t = make_tiling(a.ntaps, a.Cin, a.Cout);
  if (!t.ok || t.NT != 1 || a.u != 1) return 0;
  for (int msub = 4; msub >= 1; msub >>= 1) {
    if (forced && msub > forced) continue;
    if (msub > 1 && a.max_q <= 128 * (msub / 2)) continue;
    if (msub > 1 && msub * t.BNC * 2 > 512) continue;
    UmmaKernelArgs ka{};
    size_t smem;
    if (configure_fused(a, msub, ka, smem)) return msub;
  }
  return 0;
}

bool conv_umma_supported(const ConvArgs& a) {
  if (a.dtype != 1 && a.dtype != 2) return false;
  if (a.act_alpha) {
    if (a.dtype != 1) return false;
    UmmaKernelArgs ka{};
    size_t smem;
    return a.tile_prefix != nullptr && configure_fused(a, a.msub, ka, smem);
  }
  UmmaKernelArgs ka{};
  size_t smem;
  if (a.u != 1 && (a.res || a.accumulate)) return false;   // the transposed epilogue has no residual / accumulate path
  return a.tile_prefix != nullptr && (a.msub == 1 || a.msub == 2 || a.msub == 4) && configure(a, a.msub, ka, smem);
}

cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.max_q <= 0 || a.total_mt <= 0) return cudaSuccess;
  UmmaKernelArgs ka{};
  size_t smem;
  const bool fuse = a.act_alpha != nullptr;
  if (fuse ? !configure_fused(a, a.msub, ka, smem) : !configure(a, a.msub, ka, smem)) return cudaErrorInvalidValue;
  ka.tile_prefix = a.tile_prefix;
  ka.total_mt = a.total_mt;
  // per-device one-shot state: MaxDynamicSharedMemorySize is a per-device function attribute and the grid is
  // sized from the device's own SM count (a process may drive several GPUs)
  static int sms_of_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!sms_of_dev[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma_kernel<false, EPIW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_umma_kernel<false, EPIW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e != cudaSuccess) return e;
    sms_of_dev[dev] = n;
  }
  const int num_sms = sms_of_dev[dev];
  static const int trace_cin = env_int("BVG_CONV_TRACE", 0);   // e.g. 24: trace the first k=3 conv with Cin == 24
  static bool traced = false;
  static const int trace_taps = env_int("BVG_CONV_TRACE_TAPS", 3);
  static const int trace_res = env_int("BVG_CONV_TRACE_RES", 0);   // 1: trace a layer with a residual input
  const bool do_trace = trace_cin > 0 && !traced && a.Cin == trace_cin && a.u == 1 && a.ntaps == trace_taps &&
                        (a.res != nullptr) == (trace_res != 0);
  if (do_trace) {
    cudaMalloc((void**)&ka.trace, (4 + 4 * 1024) * 8);
    cudaMemset(ka.trace, 0, (4 + 4 * 1024) * 8);
    traced = true;
    smem = SMEM_MAX;
  }
  const int total_tiles = ka.total_mt * ka.NT;
  // BVG_CONV_EPIW=4: 4 instead of 8 epilogue warps (192 threads, 32 K registers) -- leaves room for activation
  // blocks of a concurrent stream on the same SM (co-scheduling experiment, tools/two_stream.py)
  static const int epiw4 = env_int("BVG_CONV_EPIW", 8) == 4;
  dim3 grid(total_tiles < num_sms ? total_tiles : num_sms), block(fuse ? NTHREADS_FUSED : (epiw4 ? 64 + 32 * 4 : NTHREADS));
  cudaError_t le;
  if (a.f32io) {
    if (fuse || a.dtype != 2 || a.k_packed || a.bn_small || a.split3_chunks * 24 != a.Cin) return cudaErrorInvalidValue;
    le = launch_pdl(conv_umma_kernel<false, EPIW, true, true>, grid, dim3(NTHREADS), smem, s, ka);
  }
  else if (a.dtype == 2 && !fuse) le = launch_pdl(conv_umma_kernel<false, EPIW, true>, grid, dim3(NTHREADS), smem, s, ka);
  else if (fuse) le = launch_pdl(conv_umma_kernel<true>, grid, block, smem, s, ka);
  else if (epiw4) le = launch_pdl(conv_umma_kernel<false, 4>, grid, block, smem, s, ka);
  else le = launch_pdl(conv_umma_kernel<false>, grid, block, smem, s, ka);
  if (le != cudaSuccess) return le;
  if (do_trace) {
    cudaStreamSynchronize(s);
    static unsigned long long h[4 + 4 * 1024];
    cudaMemcpy(h, ka.trace, sizeof h, cudaMemcpyDeviceToHost);
    FILE* f = fopen("gpurun_out/conv_trace.txt", "w");
    if (f) {
      fprintf(f, "# NA %d NB %d ACC %d MSUB %d BN %d KC %d NKB %d resident %d tiles %d grid %d\n", ka.NA, ka.NB, ka.ACC,
              ka.MSUB, ka.BN, ka.KC, ka.NKB, ka.b_resident, total_tiles, (int)grid.x);
      for (int role = 0; role < 4; ++role)
        for (unsigned long long i = 0; i < h[role] && i < 1024; ++i) {
          unsigned long long v = h[4 + role * 1024 + i];
          fprintf(f, "%d %llu %llu %llu\n", role, v >> 56, (v >> 40) & 0xffff, v & 0xffffffffffULL);
        }
      fclose(f);
    }
    cudaFree(ka.trace);
  }
  return cudaGetLastError();
}
