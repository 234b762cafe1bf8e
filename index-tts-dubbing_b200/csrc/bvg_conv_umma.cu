// tcgen05 (5th-gen tensor core) implicit-GEMM 1-D convolution for sm_100a, bf16 in / fp32 accumulate.
//
// GEMM view (same as bvg_conv_simt.cu):  D[q, n] = sum_tap sum_ci X[q + tap_off[tap], ci] * W[tap][ci][n]
//   M = 128 time rows per CTA (UMMA_M = 128, cta_group::1), N = BN <= 256 columns per CTA,
//   K = taps x Cin, walked as k-blocks of KC 8-channel chunks (KC*8 channels) x taps.
//
// Operand staging (no tensor maps needed -- the HBM layouts ARE the shared-memory images):
//   A  activations, packed c8 layout [Cin/8][R][8] bf16.  For one k-block the producer issues KC
//      bulk copies (cp.async.bulk, mbarrier complete_tx) of AR = 128 + span contiguous rows, one per
//      8-channel chunk, giving the UMMA no-swizzle K-major layout [chunk][row][16 B]:
//      core matrix = 8 rows x 16 B contiguous, SBO (next 8 rows) = 128 B, LBO (next K chunk) =
//      ASTRIDE*16 B.  Because rows are 16 B apart, a dilated tap is just a descriptor start-address
//      shift of tap_off*16 B: the tile (with halo) is loaded ONCE per k-block and reused by all taps.
//      Zero "same" padding and the tile halo come from the zero guard rows of the packed layout.
//   B  weights, pre-packed by launch_repack_umma into per-(n-tile, k-block, tap) images
//      [chunk][n][16 B] (same canonical layout, LBO = BN*16 B): one bulk copy per pipeline stage.
//   D  fp32 accumulator in TMEM (BN columns x 128 lanes); epilogue warps read it with tcgen05.ld
//      (lane = time row, 8 consecutive columns = one 16-byte c8 vector) and apply
//      bias / residual / scale / accumulate before a coalesced 16-byte store.
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include <cstdlib>

#include "bvg_common.cuh"
#include "bvg_misc.cuh"

namespace {

constexpr int BM = 128;
constexpr int MAXSPAN = 50;
constexpr int ASTRIDE = BM + MAXSPAN + 6;   // 184 rows between K chunks of the A stage
constexpr int NTHREADS = 192;
constexpr int SMEM_LIMIT = 200 * 1024;

struct UmmaTiling {
  int KC;        // 8-channel chunks per k-block
  int NKB;       // k-blocks
  int BN;        // columns per CTA (multiple of 16, <= 256)
  int NT;        // n tiles
  int tmem_cols; // power of two >= 32
  int NA, NB;    // pipeline depths
  size_t a_stage_bytes, b_stage_bytes, smem_bytes;
  bool ok;
};

__host__ __device__ inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

inline UmmaTiling make_tiling(int ntaps, int Cin, int N) {
  UmmaTiling t{};
  t.ok = false;
  if (Cin % 8 || N % 8 || ntaps < 1 || ntaps > BVG_MAX_TAPS) return t;
  const int cin_pad = round_up_i(Cin, 16);
  if (cin_pad <= 128) { t.KC = cin_pad / 8; t.NKB = 1; }
  else if (cin_pad % 64 == 0) { t.KC = 8; t.NKB = cin_pad / 64; }
  else return t;
  const int n_pad = round_up_i(N, 16);
  t.NT = (n_pad + 255) / 256;
  t.BN = round_up_i((n_pad + t.NT - 1) / t.NT, 16);
  t.tmem_cols = 32;
  while (t.tmem_cols < t.BN) t.tmem_cols <<= 1;
  t.a_stage_bytes = (size_t)t.KC * ASTRIDE * 16;
  t.b_stage_bytes = (size_t)t.KC * t.BN * 16;
  t.NA = t.NKB > 1 ? 2 : 1;
  const int total_b = t.NKB * ntaps;
  t.NB = total_b < 4 ? total_b : 4;
  while (t.NB > 2 && t.NA * t.a_stage_bytes + t.NB * t.b_stage_bytes + 1024 > (size_t)SMEM_LIMIT) --t.NB;
  t.smem_bytes = t.NA * t.a_stage_bytes + t.NB * t.b_stage_bytes + 1024;
  t.ok = t.smem_bytes <= (size_t)SMEM_LIMIT;
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

struct UmmaKernelArgs {
  ConvArgs c;
  int KC, NKB, BN, tmem_cols, NA, NB;
  int a_stage_bytes, b_stage_bytes;
  int kc_last_load;   // real (non-padding) chunks of the last k-block
  int minoff, span;
  int swap_lbo_sbo;   // debug knob (BVG_UMMA_SWAP=1): exchange the LBO / SBO descriptor fields
};

__global__ void __launch_bounds__(NTHREADS, 1) conv_umma_kernel(const UmmaKernelArgs ka) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvArgs& a = ka.c;
  const int b = blockIdx.z;
  const SegDesc si = a.seg_in[b], so = a.seg_out[b];
  const int q0 = blockIdx.x * BM;
  const int Lq = si.len + a.q_extra;
  if (q0 >= Lq) return;
  const int ntile = blockIdx.y;
  const int n0 = ntile * ka.BN;
  const int N = a.u * a.Cout;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem carve-up: [A stages][B stages][barriers]
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + (size_t)ka.NA * ka.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)ka.NB * ka.b_stage_bytes);
  // bars: [0,NA) a_full, [NA,2NA) a_empty, [2NA, 2NA+NB) b_full, [2NA+NB, 2NA+2NB) b_empty, then tmem_full
  const uint32_t bar0 = smem_u32(bars);
  auto A_FULL = [&](int s) { return bar0 + 8u * s; };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (ka.NA + s); };
  auto B_FULL = [&](int s) { return bar0 + 8u * (2 * ka.NA + s); };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 * ka.NA + ka.NB + s); };
  const uint32_t TMEM_FULL = bar0 + 8u * (2 * ka.NA + 2 * ka.NB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ka.NA + 2 * ka.NB + 1);

  // padding chunks (Cin not a multiple of 16) must read as zero: clear the A stages once
  if (ka.kc_last_load < ka.KC) {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* p = reinterpret_cast<uint4*>(a_smem);
    const int n16 = ka.NA * ka.a_stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += NTHREADS) p[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < ka.NA; ++s) { mbar_init(A_FULL(s), 1); mbar_init(A_EMPTY(s), 1); }
    for (int s = 0; s < ka.NB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
    mbar_init(TMEM_FULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)ka.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int AR = BM + ka.span;   // rows loaded per chunk

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(a.x);
      const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.w);
      const long long row0 = (long long)si.off + q0 + ka.minoff;
      int sa = 0, pa = 0, sb = 0, pb = 0;
      for (int kb = 0; kb < ka.NKB; ++kb) {
        const int kcl = (kb == ka.NKB - 1) ? ka.kc_last_load : ka.KC;
        mbar_wait(A_EMPTY(sa), pa ^ 1);
        mbar_expect_tx(A_FULL(sa), (uint32_t)(kcl * AR * 16));
        const uint32_t adst = smem_u32(a_smem + (size_t)sa * ka.a_stage_bytes);
        for (int c = 0; c < kcl; ++c) {
          const __nv_bfloat16* src = xg + ((size_t)(kb * ka.KC + c) * a.Rx + row0) * 8;
          bulk_g2s(adst + (uint32_t)c * ASTRIDE * 16, src, (uint32_t)(AR * 16), A_FULL(sa));
        }
        for (int tap = 0; tap < a.ntaps; ++tap) {
          mbar_wait(B_EMPTY(sb), pb ^ 1);
          mbar_expect_tx(B_FULL(sb), (uint32_t)ka.b_stage_bytes);
          const uint8_t* src = wimg + ((size_t)(ntile * ka.NKB + kb) * a.ntaps + tap) * ka.b_stage_bytes;
          bulk_g2s(smem_u32(b_smem + (size_t)sb * ka.b_stage_bytes), src, (uint32_t)ka.b_stage_bytes, B_FULL(sb));
          if (++sb == ka.NB) { sb = 0; pb ^= 1; }
        }
        if (++sa == ka.NA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ka.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t lbo_a = ASTRIDE * 16, lbo_b = (uint32_t)ka.BN * 16;
      int sa = 0, pa = 0, sb = 0, pb = 0;
      uint32_t accum = 0;
      for (int kb = 0; kb < ka.NKB; ++kb) {
        mbar_wait(A_FULL(sa), pa);
        tc_fence_after();
        const uint32_t abase = smem_u32(a_smem + (size_t)sa * ka.a_stage_bytes);
        for (int tap = 0; tap < a.ntaps; ++tap) {
          mbar_wait(B_FULL(sb), pb);
          tc_fence_after();
          const uint32_t bbase = smem_u32(b_smem + (size_t)sb * ka.b_stage_bytes);
          const uint32_t ashift = (uint32_t)(a.tap_off[tap] - ka.minoff) * 16;
          for (int k16 = 0; k16 < ka.KC / 2; ++k16) {
            const uint32_t aaddr = abase + ashift + (uint32_t)(2 * k16) * lbo_a;
            const uint32_t baddr = bbase + (uint32_t)(2 * k16) * lbo_b;
            const uint64_t adesc = ka.swap_lbo_sbo ? make_desc(aaddr, 128, lbo_a) : make_desc(aaddr, lbo_a, 128);
            const uint64_t bdesc = ka.swap_lbo_sbo ? make_desc(baddr, 128, lbo_b) : make_desc(baddr, lbo_b, 128);
            umma_bf16(tmem_base, adesc, bdesc, idesc, accum);
            accum = 1;
          }
          umma_commit(B_EMPTY(sb));
          if (++sb == ka.NB) { sb = 0; pb ^= 1; }
        }
        umma_commit(A_EMPTY(sa));
        if (++sa == ka.NA) { sa = 0; pa ^= 1; }
      }
      umma_commit(TMEM_FULL);
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q = q0 + row;
    mbar_wait(TMEM_FULL, 0);
    tc_fence_after();
    __nv_bfloat16* yg = reinterpret_cast<__nv_bfloat16*>(a.y);
    const __nv_bfloat16* rg = reinterpret_cast<const __nv_bfloat16*>(a.res);
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int cg = 0; cg < ka.BN; cg += 16) {
      uint32_t r[16];
      __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
      tmem_ld16(trow + (uint32_t)cg, r);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + cg + 8 * h;
        if (n >= N) continue;
        const int phase = n / a.Cout, co = n - phase * a.Cout;
        const int orow = q * a.u + phase - a.p;
        if (!(q < Lq && orow >= 0 && orow < so.len)) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * h + j]);
        if (a.bias) {
          const float4* bp = reinterpret_cast<const float4*>(a.bias + (size_t)b * a.bias_bstride + co);
          float4 b0 = bp[0], b1 = bp[1];
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        const size_t o = ((size_t)(co >> 3) * a.Ry + so.off + orow) * 8;
        if (rg) {
          Vec8<__nv_bfloat16> rv;
          rv.load(rg + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rv.v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= a.out_scale;
        if (a.accumulate) {
          Vec8<__nv_bfloat16> ov;
          ov.load(yg + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += ov.v[j];
        }
        Vec8<__nv_bfloat16> outv;
#pragma unroll
        for (int j = 0; j < 8; ++j) outv.v[j] = v[j];
        outv.store(yg + o);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)ka.tmem_cols);
  }
}

// fp32 [tap][Cin][N]  ->  bf16 images [ntile][kb][tap][chunk KC][n BN][8]
__global__ void repack_umma_kernel(const float* __restrict__ wt, __nv_bfloat16* __restrict__ img, int ntaps, int Cin,
                                   int N, int KC, int NKB, int BN, int NT) {
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
  img[idx] = __float2bfloat16_rn(v);
}

}  // namespace

size_t umma_weight_image_bytes(int ntaps, int Cin, int N) {
  UmmaTiling t = make_tiling(ntaps, Cin, N);
  if (!t.ok) return 0;
  return (size_t)t.NT * t.NKB * ntaps * t.b_stage_bytes;
}

cudaError_t launch_repack_umma(const float* wp_tap_major, void* img, int ntaps, int Cin, int N, cudaStream_t s) {
  UmmaTiling t = make_tiling(ntaps, Cin, N);
  if (!t.ok) return cudaErrorInvalidValue;
  size_t total = (size_t)t.NT * t.NKB * ntaps * t.KC * t.BN * 8;
  repack_umma_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(wp_tap_major, (__nv_bfloat16*)img, ntaps, Cin, N,
                                                                    t.KC, t.NKB, t.BN, t.NT);
  return cudaGetLastError();
}

static void tap_range(const ConvArgs& a, int& mn, int& mx) {
  mn = mx = a.tap_off[0];
  for (int j = 1; j < a.ntaps; ++j) { mn = a.tap_off[j] < mn ? a.tap_off[j] : mn; mx = a.tap_off[j] > mx ? a.tap_off[j] : mx; }
}

bool conv_umma_supported(const ConvArgs& a) {
  UmmaTiling t = make_tiling(a.ntaps, a.Cin, a.u * a.Cout);
  if (!t.ok) return false;
  int mn, mx;
  tap_range(a, mn, mx);
  return (mx - mn) <= MAXSPAN && -mn <= BVG_GUARD && mx <= BVG_GUARD;
}

cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.max_q <= 0) return cudaSuccess;
  const int N = a.u * a.Cout;
  UmmaTiling t = make_tiling(a.ntaps, a.Cin, N);
  if (!t.ok) return cudaErrorInvalidValue;
  UmmaKernelArgs ka{};
  ka.c = a;
  ka.KC = t.KC; ka.NKB = t.NKB; ka.BN = t.BN; ka.tmem_cols = t.tmem_cols; ka.NA = t.NA; ka.NB = t.NB;
  ka.a_stage_bytes = (int)t.a_stage_bytes; ka.b_stage_bytes = (int)t.b_stage_bytes;
  const int real_chunks = a.Cin / 8;
  ka.kc_last_load = real_chunks - (t.NKB - 1) * t.KC;
  int mn, mx;
  tap_range(a, mn, mx);
  ka.minoff = mn; ka.span = mx - mn;
  {
    const char* e = getenv("BVG_UMMA_SWAP");
    ka.swap_lbo_sbo = (e && e[0] == '1') ? 1 : 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.max_q + BM - 1) / BM, t.NT, a.B), block(NTHREADS);
  conv_umma_kernel<<<grid, block, t.smem_bytes, s>>>(ka);
  return cudaGetLastError();
}
