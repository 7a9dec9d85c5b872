// Split-fp16 tcgen05 kernels for SYMMETRIC metric tables, latent_dim == 16 (sm_100a).
//
// Same mathematics as rlvae_tc.cu (G^{-1}[n] = sum_k exp(-||z_n-c_k||^2/T^2) M_k + lambda I,
// ref src/models/components/metric_tensor.py:115-135), but the weighted-sum GEMM runs on
// kind::f16 at twice the kind::tf32 rate with the same ~2^-22 relative accuracy:
//
//   P' = 2^14 exp(..)        in (0, 2^14]          P_hi = fp16(P'),  P_lo = fp16(P' - P_hi)
//   M' = 2^e M, max|M'| <= 2^14 (e per table)       M_hi = fp16(M'),  M_lo = fp16(M' - M_hi)
//   O  = P_hi.M_hi + P_lo.M_hi + P_hi.M_lo          (fp32 accumulate in TMEM; the dropped P_lo.M_lo
//                                                    term is 2^-24 relative)
// Both operands are exact sums of two fp16 numbers down to an ABSOLUTE floor of 2^-25 in the scaled
// units (fp16 subnormal spacing), i.e. 2^-39 relative to the largest weight / table entry, which is
// why the power-of-two pre-scaling matters and why it is exact.  The result is un-scaled by
// 2^-(14+e) in the epilogue.  The distance GEMM (GEMM1) is split the same way (z' = 2^ez z per point,
// c' = 2^ec c per table, S = 2^-(ez+ec) (z'_hi.c'_hi + z'_hi.c'_lo + z'_lo.c'_hi), K = 16 = one MMA per
// term), or -- EXACT mode, for temperatures where the expanded distance form is too inaccurate --
// replaced by exact differences on the FMA pipe.
//
// Because P is half as wide in TMEM (two fp16 per 32-bit column) one CTA now owns ALL 144 packed
// columns of its 128 points (no column halves -> GEMM1 and the exp stage are not duplicated) and
// there is room for TWO chunk accumulators, so the fp32 folding of a finished chunk overlaps the
// accumulation of the next one.
//
//   TMEM columns: [0,192) three S/P buffers (64 each: S fp32, overwritten in place by
//                 P_hi[0:32) | P_lo[0:32) | P_hi[32:64) | P_lo[32:64) as packed fp16),
//                 [192,336) and [352,496) the two chunk accumulators (N = 144),
//                 [336,344) z'_hi and [496,504) z'_lo (fp16 split of z: the A operand of GEMM1).
//   Warps       : 0 TMA (centroid tiles + bias), 1 MMA issuer, 2 TMA (table tiles), 3 idle;
//                 4-7 / 8-11 exp groups (even / odd super-blocks), 12-15 fold group.
//   FUSED       : the fold warpgroup keeps the finished 136 entries of its point in registers and
//                 runs the per-thread Cholesky of rlvae_perpoint.cu on them, so log det G and
//                 diag(G) (and packed / expanded G, expanded G^{-1}) leave the kernel directly, stored
//                 from the registers of the owning thread -- the "fused metric + log-det" kernel of
//                 the north star.  Points that are not positive definite go to the same
//                 fallback list.
#include <cuda_fp16.h>

#include <type_traits>

#include "rlvae_tc_common.cuh"

// -DRLVAE_TC_PROFILE: CTA 0 prints where its MMA warp and exp group A spend their cycles
#ifdef RLVAE_TC_PROFILE
#define PROF_T0() long long _pt = clock64()
#define PROF_ADD(acc) do { long long _n = clock64(); (acc) += _n - _pt; _pt = _n; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(acc) do {} while (0)
#endif

#ifdef RLVAE_TC_PROFILE
__device__ long long g_h16_trace[8][16];   // [event][block - 40]
#define HTRACE(ev, j) do { if (blockIdx.x == 0 && (j) >= 40 && (j) < 56 && lane == 0) g_h16_trace[ev][(j) - 40] = clock64(); } while (0)
#else
#define HTRACE(ev, j) do {} while (0)
#endif

namespace rlvae {
namespace tc {
namespace h16 {

constexpr int THREADS = 512;
constexpr int C_STAGES = 3;
constexpr int SP_BUFS = 3;
constexpr int M_STAGES = 3;                             // one stage = the hi AND the lo tile of a super-block
constexpr int AHEAD = 3;
constexpr bool COLSPLIT = false;                         // exp groups split every super-block by columns: measured slower here
                                                         // (7.31 vs 6.75 ms; three S|P buffers already hide the exp latency)
constexpr int NCOLS = 144;                               // packed columns = MMA N of GEMM2
constexpr uint32_t M_HALF_BYTES = NCOLS * 128;           // [144 rows x 64 centroids] fp16 (pair: 72 rows used)
constexpr uint32_t M_TILE_BYTES = 2 * M_HALF_BYTES;      // hi tile, then lo tile
constexpr uint32_t OFF_C = OFF_A2 + A_BYTES;
constexpr uint32_t OFF_M = OFF_C + C_STAGES * C_TILE_BYTES;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 4 + 3 + 2;   // + 3 pipe-trace barriers (profiling builds)
                                                                                   // + 2 "next position ready" barriers (HMC mode)
// HMC mode: per-chain state rows in the (otherwise unused) A-tile region of shared memory:
// [0,16) z, [16,32) rho_half, [32] zb, [33] s_scale, [34] log det G^{-1} and [35] its validity at the state the
// running MCMC iteration started from, [36,52) diag G there (re-used instead of re-evaluated after a rejection)
constexpr int ST_LD = 52;
constexpr uint32_t OFF_STATE = OFF_A1;
static_assert(TILE_M * ST_LD * 4 <= 2 * A_BYTES, "HMC state rows must fit the A-tile region");
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
constexpr uint32_t OFF_HYB = (OFF_TMEM_PTR + 16 + 15) & ~15u;   // HYBRID only: one scratch row per exp warp (8 warps)
constexpr uint32_t OFF_HYBC = OFF_HYB + 8 * HYB_ROW_BYTES;      // HYBRID only: natural fp32 rows of the block's 64 centroids, per C stage
constexpr uint32_t SMEM_BYTES_HYB = OFF_HYBC + C_STAGES * BK * 64 + 1024;
static_assert(OFF_HYB % 16 == 0 && SMEM_BYTES_HYB <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_SP = 0;        // + buf*64
constexpr uint32_t TM_ACC = 192;     // + buf*160
constexpr uint32_t TM_ZHI = 336;     // z (tf32 hi) 16 columns: A operand of GEMM1, in the hole between the accumulators
constexpr uint32_t TM_ZLO = 496;     // z - hi, 16 columns
constexpr float P_SHIFT = 14.f;      // P' = 2^14 P

}  // namespace h16

#define SYM_L(r, c) a[sym_index((c), (r))]   /* lower-triangular entry (r >= c) of the packed array */

// Per-thread Cholesky / inverse on the 136 packed entries (same algorithm as sym16_cholesky_kernel
// in rlvae_perpoint.cu).  a[] holds A on entry and packed G = A^{-1} on exit (when want_g).
__device__ __forceinline__ bool sym16_factor(float (&a)[144], float& lad, float (&dg)[16], bool want_g) {
  float rd[16];
  bool ok = true;
  // log det = sum_j log d_j as ONE logf: the pivots' mantissas are multiplied (16 values in [1,2): no overflow) and
  // their exponents added as integers -- 16 logf calls were ~a fifth of the instructions of this factorisation
  float mant = 1.f;
  int expo = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float d = SYM_L(j, j);
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-SYM_L(j, k), SYM_L(j, k), d);
    ok = ok && (d > 0.f);
    {
      const uint32_t db = __float_as_uint(d);
      expo += (int)((db >> 23) & 0xffu) - 127;
      mant *= __uint_as_float((db & 0x007fffffu) | 0x3f800000u);
    }
    // 1/sqrt(d) by MUFU.RSQ + one Newton step (< 1 ulp), L_jj = d / sqrt(d): this is the serial part of
    // the factorisation (16 dependent columns), sqrtf + an IEEE division were a third of its latency
    float inv = rsqrtf(d);
    inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);
    const float ljj = d * inv;
    rd[j] = inv;
    SYM_L(j, j) = ljj;
#pragma unroll
    for (int i = j + 1; i < 16; ++i) {
      float s = SYM_L(i, j);
#pragma unroll
      for (int k = 0; k < j; ++k) s = fmaf(-SYM_L(i, k), SYM_L(j, k), s);
      SYM_L(i, j) = s * inv;
    }
  }
  // (a non-positive or subnormal pivot makes the bit trick meaningless: those matrices are redone by the fallback
  // pass / flagged, their lad is never used)
  lad = fmaf((float)expo, 0.6931471805599453f, logf(mant));
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    SYM_L(j, j) = rd[j];
#pragma unroll
    for (int i = j + 1; i < 16; ++i) {
      float s = 0.f;
#pragma unroll
      for (int k = j; k < i; ++k) s = fmaf(SYM_L(i, k), SYM_L(k, j), s);
      SYM_L(i, j) = -s * rd[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float s = 0.f;
#pragma unroll
    for (int k = i; k < 16; ++k) s = fmaf(SYM_L(k, i), SYM_L(k, i), s);
    dg[i] = s;
  }
  if (want_g) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = i; k < 16; ++k) s = fmaf(SYM_L(k, i), SYM_L(k, j), s);
        SYM_L(i, j) = s;
      }
    }
  }
  return ok;
}
#undef SYM_L

struct FusedOut {
  float* a_full;       // [N,16,16] G^{-1} expanded (what the reference API returns) or NULL
  float* a_packed;     // [N,144] G^{-1} (lambda on the diagonal) or NULL
  float* g_packed;     // [N,144] G or NULL
  float* g_full;       // [N,16,16] G expanded or NULL
  float* logabsdet;    // [N] lad_scale * log det G^{-1} or NULL
  float* sign;         // [N] or NULL
  float* diag_g;       // [N,16] or NULL
  int* fail_ws;        // counter + list (needed when any of g_packed / logabsdet / sign / diag_g is set)
  float lad_scale;
  float* s_diag;       // [N,16] diagonal of sum_k w_k M_k WITHOUT lambda, or NULL (pythae variant: (G^{-1} - lambda I) z
                       // must not be formed from fl(S_ii + lambda) - lambda when S_ii << lambda)
};

// HMC mode of the forward kernel: ONE launch runs n_iters MCMC iterations of
// RiemannianHMCSampler.sample (ref src/models/samplers/hmc_sampler.py:120-163) for the 128 chains of
// the CTA -- n_lf + 1 metric evaluations per iteration, the chain state (z, rho_half) in shared memory /
// registers of the owning thread, the update / accept arithmetic in the fold group's epilogue, which
// already holds diag(G) and log det.  Only the variant-A drift (1 - lambda G_ii)/T^2 is fusable (the
// exact gradient needs the second contraction kernel).
struct HmcArgs {
  float* z;              // [N,16] in: chain state; out: accepted state (re-read as z_prev at every accept step)
  const float* gamma;    // [n_iters, N, 16]  draws of hmc_sampler.py:122
  const float* acc;      // [n_iters, N]      draws of hmc_sampler.py:158
  const float* scales;   // [n_iters * n_lf] device, or NULL -> scales_inl (n_iters == 1)
  float* h0;             // [n_iters, N] or NULL
  float* h1;             // [n_iters, N] or NULL
  float* alpha;          // [n_iters, N] or NULL
  float* moves;          // [n_iters, N] or NULL
  float* z_trace;        // [n_iters, N, 16] accepted state after every iteration, or NULL
  int* fail_count;       // incremented when a Cholesky pivot is not positive (certified tables: rounding only)
  int n_iters, n_lf;
  float eps, b0, T2;
  float scales_inl[64];
};

// 0.5*log(clamp(det G^{-1}, 1e-10)) from log det (the fused Cholesky only succeeds for det > 0);
// same arithmetic as log_pi_from_slogdet in rlvae_hmc.cu -- ref hmc_sampler.py:26-30
__device__ __forceinline__ float hmc_log_pi(float lad, bool pos) {
  const float floor_lp = 0.5f * logf(1e-10f);
  if (!pos) return floor_lp;
  if (lad > 88.72283f) return INFINITY;
  return fmaxf(0.5f * lad, floor_lp);
}

// EXACT: the weights come from exact differences sum_j (z_j - c_kj)^2 formed by the exp threads on the
// FMA pipe (packed fp32x2, natural centroid rows broadcast from shared memory) instead of the
// expanded form on the tensor core.  No GEMM1, no accuracy gate on T: this is the mode for small
// temperatures (the reference's T = 0.7 configuration), ~28 instead of ~7 instructions per
// (point, centroid) in the exp stage, which then bounds the kernel instead of the tensor pipe.
// HYBRID (only with !EXACT): the expanded form runs as usual, and every weight whose exponent is above
// hyb_thr -- the only ones large enough for the expanded form's absolute exponent error to matter in
// G^{-1} (see h16_mode) -- is recomputed from exact differences (natural centroid rows through L1).  At
// small T almost every weight is far below the threshold, so this keeps most of the tensor-core speed.
template <bool PAIR, bool EXACT, bool HYBRID, bool HMC = false>
__global__ void __launch_bounds__(h16::THREADS, 1)
inverse_metric_h16_kernel(const __grid_constant__ CUtensorMap tm_cstack,
                          const __grid_constant__ CUtensorMap tm_mh_hi,
                          const __grid_constant__ CUtensorMap tm_mh_lo,
                          const float* __restrict__ z, const float* __restrict__ cbias /* EXACT: 0 / -1e30 mask */,
                          const float* __restrict__ cnat /* EXACT: natural centroid rows [Kpad,16] */, int64_t n,
                          int num_blocks, float alpha /* log2(e)/T^2 */, float lambda,
                          float out_scale /* 2^-(14+e) */, float c_unscale /* 2^-ec */,
                          const float* __restrict__ cshift /* [16] centre of the expanded form */,
                          float hyb_thr /* HYBRID: refine weights with log2(2^14 w) above this */, FusedOut fo,
                          const __grid_constant__ HmcArgs hm) {
#ifdef RLVAE_TC_PROFILE
  const long long pk0 = clock64();
#endif
  // HMC: the main loop runs over ALL evaluations of the trajectory as one long block sequence
  // (num_blocks is even, so chunk / stage / buffer phases simply continue across evaluations)
  // (n_lf + 1 evaluations for the first MCMC iteration, n_lf for every later one: the metric at the state an
  // iteration starts from is the one its predecessor ended with -- accepted: the last evaluation, rejected: the
  // predecessor's own starting metric, kept in the state row)
  const int n_evals = HMC ? hm.n_iters * hm.n_lf + 1 : 1;
  const int total_blocks = n_evals * num_blocks;
  // local names shadow the tc:: constants of the 3xTF32 kernels
  constexpr int THREADS = h16::THREADS, C_STAGES = h16::C_STAGES, SP_BUFS = h16::SP_BUFS,
                M_STAGES = h16::M_STAGES, AHEAD = h16::AHEAD, NCOLS = h16::NCOLS;
  constexpr bool COLSPLIT = h16::COLSPLIT;
  constexpr uint32_t M_TILE_BYTES = h16::M_TILE_BYTES, M_HALF_BYTES = h16::M_HALF_BYTES, TM_SP = h16::TM_SP,
                     TM_ACC = h16::TM_ACC, TM_ZHI = h16::TM_ZHI, TM_ZLO = h16::TM_ZLO;
  constexpr float P_SHIFT = h16::P_SHIFT;
  constexpr int CB = 2;     // super-blocks per tensor-core accumulation chunk (compile-time: see the MMA issuer)
  (void)THREADS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

  const uint32_t bar0 = base + h16::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_BIAS_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };
  auto BAR_M_FULL = [&](int s) { return bar0 + 8u * (3 * C_STAGES + s); };
  auto BAR_M_EMPTY = [&](int s) { return bar0 + 8u * (3 * C_STAGES + M_STAGES + s); };
  auto BAR_S_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + b); };
  auto BAR_P_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + SP_BUFS + b); };
  auto BAR_CH_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + b); };
  auto BAR_CH_FREE = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 2 + b); };
  auto BAR_MON = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 4 + b); };
  (void)BAR_MON;
  // HMC: the fold group publishes the next position -- BAR_ZR: z' in TMEM, for the MMA issuer (leader's
  // barrier, both CTAs of a pair arrive); BAR_ZL: zb / s_scale / z in shared memory, for this CTA's exp groups
  const uint32_t BAR_ZR = bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 7);
  const uint32_t BAR_ZL = bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 8);
  float* const state = reinterpret_cast<float*>(gbase + h16::OFF_STATE);
  constexpr int ST_LD = h16::ST_LD;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + h16::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  constexpr int ROWS_BOX = NCOLS / NPAIR;                 // B-tile rows held by this CTA (144 / 72)
  constexpr uint32_t TILE_BYTES = ROWS_BOX * 128;
  constexpr uint32_t IDESC_G2 = make_idesc_f16(PAIR ? 256 : 128, NCOLS);
  const int row_cta = PAIR ? (int)rank * ROWS_BOX : 0;
  const int num_chunks = (num_blocks + CB - 1) / CB;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), COLSPLIT ? 8 : 4); mbar_init(BAR_BIAS_FULL(s), 1);
    }
    for (int s = 0; s < SP_BUFS; ++s) { mbar_init(BAR_S_FULL(s), 1); mbar_init(BAR_P_FULL(s), (COLSPLIT ? 8 : 4) * NPAIR); }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR_CH_FULL(b), 1); mbar_init(BAR_CH_FREE(b), 4 * NPAIR); }
    for (int b = 0; b < 3; ++b) mbar_init(BAR_MON(b), 1);
    mbar_init(BAR_ZR, 4 * NPAIR);
    mbar_init(BAR_ZL, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cstack) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mh_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mh_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + h16::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + h16::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }

  const int quarter = warp & 3;
  const int prow = quarter * 32 + lane;
  tc_fence_before();
  __syncthreads();            // TMEM base published (the exp threads store z into TMEM below)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  float zb = 0.f, s_scale = 0.f;      // exponent = S' * s_scale + bias_k + zb
  float2 nz[8];                       // EXACT: the negated point
#pragma unroll
  for (int j = 0; j < 8; ++j) nz[j] = make_float2(0.f, 0.f);
  if (wg == 1 || wg == 2) {
    const int64_t r = row0 + prow;
    float zv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) zv[j] = 0.f;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>((HMC ? hm.z : z) + r * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = HMC ? src[q] : __ldg(src + q);      // HMC: z is rewritten by this kernel
        zv[4 * q] = v.x; zv[4 * q + 1] = v.y; zv[4 * q + 2] = v.z; zv[4 * q + 3] = v.w;
      }
    }
    if (HMC && wg == 1) {            // the chain's position for the fold thread that owns it
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(state + prow * ST_LD + 4 * q) =
            make_float4(zv[4 * q], zv[4 * q + 1], zv[4 * q + 2], zv[4 * q + 3]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) nz[j] = make_float2(-zv[2 * j], -zv[2 * j + 1]);
    if (!EXACT) {                    // the expanded form works on z~ = z - shift (tables hold c~ = c - shift)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
        zv[4 * q] -= sh.x; zv[4 * q + 1] -= sh.y; zv[4 * q + 2] -= sh.z; zv[4 * q + 3] -= sh.w;
      }
    }
    float nrm = 0.f, zmax = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { nrm = fmaf(zv[j], zv[j], nrm); zmax = fmaxf(zmax, fabsf(zv[j])); }
    zb = EXACT ? P_SHIFT : (-nrm * alpha + P_SHIFT);     // P' = 2^14 P, folded into the exponent
    // GEMM1 on kind::f16: z' = 2^ez z (per point, max|z'| in [2^13, 2^14)), split hi + lo; S = 2^-(ez+ec) S'
    int ez = 0;
    if (zmax > 0.f && zmax < 3.0e38f) {
      const int ex = (int)((__float_as_uint(zmax) >> 23) & 0xffu) - 126;
      ez = 14 - ex;
      ez = ez > 50 ? 50 : (ez < -50 ? -50 : ez);
    }
    s_scale = 2.f * alpha * c_unscale * __uint_as_float((uint32_t)(127 - ez) << 23);
    if (wg == 1 && !EXACT) {         // exp group A writes the A operand of GEMM1 into TMEM
      const float zsc = __uint_as_float((uint32_t)(ez + 127) << 23);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_pair(zv[2 * j] * zsc, zv[2 * j + 1] * zsc, hi[j], lo[j]);
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      TMEM_ST8(tmem_base + lane_addr + TM_ZHI, hi);
      TMEM_ST8(tmem_base + lane_addr + TM_ZLO, lo);
      tmem_wait_st();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define MMA_H(d, a, b, acc) do { if (PAIR) mma_ts_f16_pair(d, a, b, IDESC_G2, acc); else mma_ts_f16(d, a, b, IDESC_G2, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (wg == 0) {
    reg_dec<72>();
    if (warp == 0) {
      // =========================================================== TMA producer 1: centroid tiles + bias
      for (int jg = 0, j = 0; jg < total_blocks; ++jg, j = (j + 1 == num_blocks) ? 0 : j + 1) {   // j: block within the evaluation
        const int cs = jg % C_STAGES;
        mbar_wait(BAR_C_EMPTY(cs), ((jg / C_STAGES) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t dst = base + h16::OFF_C + cs * C_TILE_BYTES;
          if (EXACT) {      // natural rows of the 64 centroids (4 KB) + mask, both for this CTA's exp groups
            mbar_expect_tx(BAR_BIAS_FULL(cs), BIAS_BYTES + BK * 64);
            bulk_load_1d(dst, cnat + (int64_t)j * BK * 16, BK * 64, BAR_BIAS_FULL(cs));
          } else {
            if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE_BYTES);
            if (PAIR) {
              tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32 * (int)rank);
            } else {
              tma_load_2d(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK);
              tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32);
            }
            if (HYBRID) {   // + the natural rows of the block (4 KB): the refinement reads them from shared memory
              mbar_expect_tx(BAR_BIAS_FULL(cs), BIAS_BYTES + BK * 64);
              bulk_load_1d(base + h16::OFF_HYBC + cs * (BK * 64), cnat + (int64_t)j * BK * 16, BK * 64, BAR_BIAS_FULL(cs));
            } else {
              mbar_expect_tx(BAR_BIAS_FULL(cs), BIAS_BYTES);
            }
          }
          bulk_load_1d(base + h16::OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_BIAS_FULL(cs));
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // =========================================================== TMA producer 2: M tiles, as far ahead
      // as the ring allows (its own warp: a stalled centroid stage must not delay the table stream)
      for (int jg = 0, jm = 0; jg < total_blocks; ++jg, jm = (jm + 1 == num_blocks) ? 0 : jm + 1) {
        const int ms = jg % M_STAGES;
        mbar_wait(BAR_M_EMPTY(ms), ((jg / M_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * 2 * TILE_BYTES);
          const uint32_t dst = base + h16::OFF_M + ms * M_TILE_BYTES;
          if (PAIR) {
            tma_load_2d_pair(dst, &tm_mh_hi, BAR_M_FULL(ms), jm * BK, row_cta);
            tma_load_2d_pair(dst + M_HALF_BYTES, &tm_mh_lo, BAR_M_FULL(ms), jm * BK, row_cta);
          } else {
            tma_load_2d(dst, &tm_mh_hi, BAR_M_FULL(ms), jm * BK, row_cta);
            tma_load_2d(dst + M_HALF_BYTES, &tm_mh_lo, BAR_M_FULL(ms), jm * BK, row_cta);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1 && leader) {
      // =========================================================== MMA issuer (warp-converged; pair: leader only)
      // The issue loop is unrolled 6x (lcm of the chunk, S/P-buffer, C-stage and M-stage periods) so
      // that every stage index, TMEM address and almost every barrier parity is a compile-time
      // constant: the tensor pipe idles whenever this warp's own instruction stream is slower than
      // the MMAs it feeds (measured: 1.7k cycles per super-block with runtime div/mod vs 1.15k of MMA).
      const uint64_t c_desc0 = make_desc_sw128(base + h16::OFF_C);
      const uint64_t m_desc0 = make_desc_sw128(base + h16::OFF_M);
      auto gemm1 = [&](auto CSc, auto SBc) {
        constexpr int cs = decltype(CSc)::value, sb = decltype(SBc)::value;
        if (elect_one()) {
          if (!EXACT) {
            // S' = z'_hi.c'_hi + z'_hi.c'_lo + z'_lo.c'_hi on kind::f16 (K = 16 = all latent dims per MMA),
            // A operand from TMEM: 3 MMAs of 32 cycles (3xTF32 needed 6); a centroid row is
            // [c'_hi (16) | c'_lo (16) | 0] fp16 = 2 K-steps of one 128-byte swizzle row
            constexpr uint32_t id1 = make_idesc_f16(PAIR ? 256 : 128, BK);
            const uint32_t d = tmem_base + TM_SP + sb * 64;
            const uint64_t bc = c_desc0 + ((cs * C_TILE_BYTES) >> 4);
            if (PAIR) {
              mma_ts_f16_pair(d, tmem_base + TM_ZHI, bc, id1, 0);
              mma_ts_f16_pair(d, tmem_base + TM_ZHI, bc + 2, id1, 1);
              mma_ts_f16_pair(d, tmem_base + TM_ZLO, bc, id1, 1);
            } else {
              mma_ts_f16(d, tmem_base + TM_ZHI, bc, id1, 0);
              mma_ts_f16(d, tmem_base + TM_ZHI, bc + 2, id1, 1);
              mma_ts_f16(d, tmem_base + TM_ZLO, bc, id1, 1);
            }
          }
          // EXACT: no distance GEMM; S_FULL then only says "GEMM2 has finished reading this P buffer"
          COMMIT(BAR_S_FULL(sb));
        }
        __syncwarp();
      };
      uint32_t free_phase = 0;     // bit ab: parity of the next CH_FREE(ab) completion to wait for
      int jl = 0, ev = 0;          // HMC: block index within the current evaluation, evaluation index
      // jg: global block index (== block within the evaluation unless HMC); the unrolled position J only
      // fixes stage / buffer indices and parities, which continue across evaluations
      auto block = [&](auto Jc, const int jg, const uint32_t qodd /* (jg / 6) & 1 */) {
        constexpr int J = decltype(Jc)::value;
        constexpr int first = (J % CB) == 0, sb = J % SP_BUFS;
        const uint32_t ab = ((J / CB) & 1) ^ qodd;          // 6 blocks = 3 chunks: the accumulator parity flips every pass
        constexpr int ms = J % M_STAGES;
        const int j = HMC ? jl : jg;
        if (first && jg >= 2 * CB) {                      // fold group drained this accumulator
          mbar_wait(BAR_CH_FREE(ab), (free_phase >> ab) & 1u);
          free_phase ^= 1u << ab;
        }
        tc_fence_after();
        const uint32_t p = tmem_base + TM_SP + sb * 64;     // k-step kk: P_hi at (kk>>1)*32 + (kk&1)*8, P_lo 16 further
        const uint32_t acc = tmem_base + TM_ACC + ab * 160;
        const uint64_t bh = m_desc0 + ((ms * M_TILE_BYTES) >> 4);
        const uint64_t bl = bh + (M_HALF_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_H(acc, p + (kk >> 1) * 32 + (kk & 1) * 8, bh + 2 * kk, !(first && kk == 0));
        }
        __syncwarp();
        // inputs of the NEXT steps, waited for behind queued MMAs (normally long complete)
        if (jg + 1 < total_blocks) mbar_wait(BAR_M_FULL((J + 1) % M_STAGES), ((J + 1) / M_STAGES) & 1);
        if (!EXACT && j + AHEAD < num_blocks) mbar_wait(BAR_C_FULL((J + AHEAD) % C_STAGES), ((J + AHEAD) / C_STAGES) & 1);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_H(acc, p + (kk >> 1) * 32 + 16 + (kk & 1) * 8, bh + 2 * kk, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_H(acc, p + (kk >> 1) * 32 + (kk & 1) * 8, bl + 2 * kk, 1);
          COMMIT(BAR_M_EMPTY(ms));
          if ((J % CB) == CB - 1 || j == num_blocks - 1) COMMIT(BAR_CH_FULL(ab));
        }
        __syncwarp();
        if (j + AHEAD < num_blocks) {
          tc_fence_after();
          gemm1(std::integral_constant<int, (J + AHEAD) % C_STAGES>{}, std::integral_constant<int, (J + AHEAD) % SP_BUFS>{});
        }
        if (HMC && j == num_blocks - 1 && jg + 1 < total_blocks) {
          // last block of an evaluation: the distance GEMMs of the next one need the chains' new
          // positions, which the fold group publishes after its epilogue (z' in TMEM of both CTAs)
          mbar_wait(BAR_ZR, (uint32_t)ev & 1u);
          tc_fence_after();
          if (1 <= num_blocks) {
            if (!EXACT) mbar_wait(BAR_C_FULL((J + 1) % C_STAGES), ((J + 1) / C_STAGES) & 1);
            gemm1(std::integral_constant<int, (J + 1) % C_STAGES>{}, std::integral_constant<int, (J + 1) % SP_BUFS>{});
          }
          if (2 <= num_blocks) {
            if (!EXACT) mbar_wait(BAR_C_FULL((J + 2) % C_STAGES), ((J + 2) / C_STAGES) & 1);
            gemm1(std::integral_constant<int, (J + 2) % C_STAGES>{}, std::integral_constant<int, (J + 2) % SP_BUFS>{});
          }
          if (3 <= num_blocks) {
            if (!EXACT) mbar_wait(BAR_C_FULL((J + 3) % C_STAGES), ((J + 3) / C_STAGES) & 1);
            gemm1(std::integral_constant<int, (J + 3) % C_STAGES>{}, std::integral_constant<int, (J + 3) % SP_BUFS>{});
          }
        }
        if (jg + 1 < total_blocks && (HMC || j + 1 < num_blocks))
          mbar_wait(BAR_P_FULL((J + 1) % SP_BUFS), ((J + 1) / SP_BUFS) & 1);
        if (HMC) { if (++jl == num_blocks) { jl = 0; ++ev; } }
      };
      // prologue: GEMM1 of the first AHEAD (= 3) blocks
      if (0 < num_blocks) { if (!EXACT) mbar_wait(BAR_C_FULL(0), 0); tc_fence_after(); gemm1(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{}); }
      if (1 < num_blocks) { if (!EXACT) mbar_wait(BAR_C_FULL(1), 0); tc_fence_after(); gemm1(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{}); }
      if (2 < num_blocks) { if (!EXACT) mbar_wait(BAR_C_FULL(2), 0); tc_fence_after(); gemm1(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{}); }
      mbar_wait(BAR_M_FULL(0), 0);
      mbar_wait(BAR_P_FULL(0), 0);
      static_assert(6 % CB == 0 && 6 % SP_BUFS == 0 && 6 % C_STAGES == 0 && 6 % M_STAGES == 0 && AHEAD == 3,
                    "the 6x unrolled issue loop assumes these periods");
      uint32_t qodd = 0;
      for (int j0 = 0; j0 < total_blocks; j0 += 6, qodd ^= 1u) {
#define RLVAE_BLK(J) if (j0 + J < total_blocks) block(std::integral_constant<int, J>{}, j0 + J, qodd);
        RLVAE_BLK(0) RLVAE_BLK(1) RLVAE_BLK(2) RLVAE_BLK(3) RLVAE_BLK(4) RLVAE_BLK(5)
#undef RLVAE_BLK
      }
    }
  } else if (wg == 1 || wg == 2) {
    // =========================================================== exp groups (one thread per point)
    reg_dec<104>();
    const int grp = wg - 1;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // HYBRID: this lane's scratch row (the launch adds HYB_BYTES of shared memory behind the barriers)
    float* hyb_row = reinterpret_cast<float*>(gbase + h16::OFF_HYB + (warp - 4) * HYB_ROW_BYTES) + lane * 4;
    (void)hyb_row;
    long long pe_wait = 0, pe_work = 0;
    (void)pe_wait; (void)pe_work;
    // COLSPLIT: both groups work on every super-block, 32 centroids each (halves the latency between
    // GEMM1(j) and GEMM2(j)); otherwise the groups alternate whole super-blocks
    int ev_cur = 0;                 // HMC: evaluation whose per-point constants this thread holds
    for (int jg = COLSPLIT ? 0 : grp; jg < total_blocks; jg += COLSPLIT ? 1 : 2) {
      PROF_T0();
      int j = jg;                   // block within the evaluation (selects the table rows)
      if (HMC) {
        const int e = jg / num_blocks;
        j = jg - e * num_blocks;
        if (e != ev_cur) {          // first block of a new evaluation: the chain has moved
          ev_cur = e;
          mbar_wait(BAR_ZL, (uint32_t)(e - 1) & 1u);
          const float* st = state + prow * ST_LD;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 v = *reinterpret_cast<const float2*>(st + 2 * q);
            nz[q] = make_float2(-v.x, -v.y);
          }
          zb = st[32];
          s_scale = st[33];
        }
      }
      const int cs = jg % C_STAGES, sb = jg % SP_BUFS;
      const uint32_t sp = tmem_base + lane_addr + TM_SP + sb * 64;
      mbar_wait(BAR_BIAS_FULL(cs), (jg / C_STAGES) & 1);
      mbar_wait(BAR_S_FULL(sb), (jg / SP_BUFS) & 1);
      tc_fence_after();
      if (quarter == 0) HTRACE(3, j);
      PROF_ADD(pe_wait);
#pragma unroll
      for (int rr = 0; rr < (COLSPLIT ? 1 : 2); ++rr) {
        const int rnd = COLSPLIT ? grp : rr;
        uint32_t ph[16], pl[16];
        const float4* bias4 = reinterpret_cast<const float4*>(gbase + h16::OFF_BIAS + cs * BIAS_BYTES) + rnd * 8;
        if (EXACT) {
          const float4* crow = reinterpret_cast<const float4*>(gbase + h16::OFF_C + cs * C_TILE_BYTES) + rnd * 32 * 4;
#pragma unroll 2
          for (int q = 0; q < 8; ++q) {
            const float4 bv = bias4[q];
            const float w0 = ex2_approx(fmaf(dist2_row16(crow + (4 * q) * 4, nz), -alpha, bv.x + zb));
            const float w1 = ex2_approx(fmaf(dist2_row16(crow + (4 * q + 1) * 4, nz), -alpha, bv.y + zb));
            const float w2 = ex2_approx(fmaf(dist2_row16(crow + (4 * q + 2) * 4, nz), -alpha, bv.z + zb));
            const float w3 = ex2_approx(fmaf(dist2_row16(crow + (4 * q + 3) * 4, nz), -alpha, bv.w + zb));
            split_pair(w0, w1, ph[2 * q], pl[2 * q]);
            split_pair(w2, w3, ph[2 * q + 1], pl[2 * q + 1]);
          }
        } else {
          uint32_t s[32];
          TMEM_LD32(sp + rnd * 32, s);
          tmem_wait_ld();
          if (HYBRID) {
            uint32_t live = 0;             // weights of this thread that need exact differences
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bv = bias4[q];
              const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * q + e;
                const float ex = fmaf(__uint_as_float(s[i]), s_scale, b4[e] + zb);
                s[i] = __float_as_uint(ex);
                live |= (ex > hyb_thr ? 1u : 0u) << i;
              }
            }
            refine_exponents<true>(s, live,
                             reinterpret_cast<const float4*>(gbase + h16::OFF_HYBC + cs * (BK * 64)) + rnd * 32 * 4, nz,
                             -alpha, P_SHIFT, hyb_row);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float w0 = ex2_approx(__uint_as_float(s[4 * q])), w1 = ex2_approx(__uint_as_float(s[4 * q + 1]));
              const float w2 = ex2_approx(__uint_as_float(s[4 * q + 2])), w3 = ex2_approx(__uint_as_float(s[4 * q + 3]));
              split_pair(w0, w1, ph[2 * q], pl[2 * q]);
              split_pair(w2, w3, ph[2 * q + 1], pl[2 * q + 1]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bv = bias4[q];
              const float w0 = ex2_approx(fmaf(__uint_as_float(s[4 * q]), s_scale, bv.x + zb));
              const float w1 = ex2_approx(fmaf(__uint_as_float(s[4 * q + 1]), s_scale, bv.y + zb));
              const float w2 = ex2_approx(fmaf(__uint_as_float(s[4 * q + 2]), s_scale, bv.z + zb));
              const float w3 = ex2_approx(fmaf(__uint_as_float(s[4 * q + 3]), s_scale, bv.w + zb));
              split_pair(w0, w1, ph[2 * q], pl[2 * q]);
              split_pair(w2, w3, ph[2 * q + 1], pl[2 * q + 1]);
            }
          }
        }
        TMEM_ST16(sp + rnd * 32, ph);
        TMEM_ST16(sp + rnd * 32 + 16, pl);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_P_FULL(sb)); else mbar_arrive(BAR_P_FULL(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
      if (quarter == 0) HTRACE(4, j);
      PROF_ADD(pe_work);
    }
#ifdef RLVAE_TC_PROFILE
    if ((blockIdx.x == 0 || blockIdx.x == 4096) && threadIdx.x == 128)
      printf("[h16 prof %d] exp group A per own super-block: wait S %lld  work %lld\n", (int)blockIdx.x,
             pe_wait / (num_blocks / 2), pe_work / (num_blocks / 2));
#endif
  } else {
    // =========================================================== fold group: fp32 running total + output
    reg_inc<232>();
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float total[NCOLS];
#pragma unroll
    for (int i = 0; i < NCOLS; ++i) total[i] = 0.f;
    long long pf_wait = 0, pf_work = 0;
    (void)pf_wait; (void)pf_work;
#ifdef RLVAE_TC_PROFILE
    const long long pf0 = clock64();
#endif
    const int total_chunks = n_evals * num_chunks;
    const int64_t rows_here = (n - row0 < TILE_M) ? ((n - row0 > 0) ? (n - row0) : 0) : TILE_M;   // 0 for a pair's padding tile
    const bool live = prow < rows_here;
    float h0_keep = 0.f;             // HMC: H0 of the running MCMC iteration
    for (int ev = 0, cg0 = 0; ev < n_evals; ++ev, cg0 += num_chunks) {
    for (int cl = 0; cl < num_chunks; ++cl) {
      const int c = cg0 + cl;        // global chunk index: accumulator / phase bookkeeping continues across evaluations
      PROF_T0();
      const int ab = c & 1;
      mbar_wait(BAR_CH_FULL(ab), (c >> 1) & 1);
      tc_fence_after();
      if (quarter == 0) HTRACE(5, c * CB);
      PROF_ADD(pf_wait);
      const uint32_t src = tmem_base + lane_addr + TM_ACC + ab * 160;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t a[32];
        TMEM_LD32(src + cb * 32, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) total[cb * 32 + i] += __uint_as_float(a[i]);
      }
      {
        uint32_t a[16];
        TMEM_LD16(src + 128, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) total[128 + i] += __uint_as_float(a[i]);
      }
      if (c + 2 < total_chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE(ab)); else mbar_arrive(BAR_CH_FREE(ab)); }
      }
      if (quarter == 0) HTRACE(6, c * CB);
      PROF_ADD(pf_work);
    }
#ifdef RLVAE_TC_PROFILE
    const long long pf1 = clock64();     // (printed after the epilogue: a printf here would be timed as epilogue)
#endif
    // ---------------------------------------------------------- epilogue (all TMA / MMA work of this evaluation is complete)
    if (!HMC && fo.s_diag != nullptr && live) {
      float4* dst = reinterpret_cast<float4*>(fo.s_diag + (row0 + prow) * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_float4(total[sym_index(4 * q, 4 * q)] * out_scale, total[sym_index(4 * q + 1, 4 * q + 1)] * out_scale,
                             total[sym_index(4 * q + 2, 4 * q + 2)] * out_scale,
                             total[sym_index(4 * q + 3, 4 * q + 3)] * out_scale);
    }
#pragma unroll
    for (int pc = 0; pc < NCOLS; ++pc) {
      bool diag = false;
#pragma unroll
      for (int i = 0; i < 16; ++i) diag |= (pc == sym_index(i, i));
      total[pc] = (pc < 136) ? fmaf(total[pc], out_scale, diag ? lambda : 0.f) : 0.f;
    }
    if (HMC) {
      // ---- one leapfrog stage of ref hmc_sampler.py:120-163 for the chain this thread owns.
      // Evaluation k of an iteration (k = 0 .. n_lf) is the metric at the chain's position after k
      // position updates: one metric evaluation per leapfrog step (the reference re-evaluates the same
      // gradient at the end of step k and the start of step k + 1).
      int it = 0, k = 0;               // evaluation 0 = start of iteration 0; then n_lf evaluations per iteration
      if (ev > 0) { it = (ev - 1) / hm.n_lf; k = (ev - 1) - it * hm.n_lf + 1; }
      const int64_t r = row0 + prow;
      if (!live) {
#pragma unroll
        for (int i = 0; i < 136; ++i) total[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) total[sym_index(i, i)] = 1.f;
      }
      float lad, dg[16];
      bool ok = sym16_factor(total, lad, dg, false);
      if (!ok) {                       // certified tables: rounding only.  Counted; the host redoes the call unfused.
        if (live) atomicAdd(hm.fail_count, 1);
#pragma unroll
        for (int i = 0; i < 16; ++i) dg[i] = 0.f;
      }
      float* st = state + prow * ST_LD;
      float zc[16], rh[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(st + 4 * q);
        zc[4 * q] = v.x; zc[4 * q + 1] = v.y; zc[4 * q + 2] = v.z; zc[4 * q + 3] = v.w;
      }
      const float half_eps = hm.eps / 2.f;
      bool moved_on = true;            // false after the accept step of the LAST iteration (nothing left to evaluate)
      bool begin_iteration = (k == 0); // run the first stage of iteration `it_begin` on (dg, lad, ok) at zc
      int it_begin = it;
      if (k > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(st + 16 + 4 * q);
          rh[4 * q] = v.x; rh[4 * q + 1] = v.y; rh[4 * q + 2] = v.z; rh[4 * q + 3] = v.w;
        }
        const float scale = (hm.scales != nullptr) ? __ldg(hm.scales + it * hm.n_lf + (k - 1)) : hm.scales_inl[k - 1];
        if (k < hm.n_lf) {
          // second half step + tempering (:141-148), then the next step's first half (:132-138)
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float g = -((1.f - lambda * dg[jj]) / hm.T2);
            const float rho = scale * (rh[jj] - half_eps * g);
            rh[jj] = rho - half_eps * g;
            zc[jj] = zc[jj] + hm.eps * rh[jj];
          }
        } else {
          // end of the trajectory: H (:153), alpha (:156-157), accept / reject (:158-162)
          float ss = 0.f;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float g = -((1.f - lambda * dg[jj]) / hm.T2);
            const float rho = scale * (rh[jj] - half_eps * g);
            ss = fmaf(rho, rho, ss);
          }
          const float nrm = sqrtf(ss);
          const float H = -hmc_log_pi(lad, ok) + 0.5f * (nrm * nrm);
          float a = expf(-H) / (expf(-h0_keep) + 1e-10f);
          a = fminf(fmaxf(a, 0.f), 1.f);
          const float u = live ? __ldg(hm.acc + (int64_t)it * n + r) : 2.f;
          const bool mv = u < a;
          if (live) {
            float4* zrow = reinterpret_cast<float4*>(hm.z + r * 16);
            if (!mv) {                 // back to the state this iteration started from
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 v = zrow[q];
                zc[4 * q] = v.x; zc[4 * q + 1] = v.y; zc[4 * q + 2] = v.z; zc[4 * q + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) zrow[q] = make_float4(zc[4 * q], zc[4 * q + 1], zc[4 * q + 2], zc[4 * q + 3]);
            }
            if (hm.z_trace != nullptr) {
              float4* tr = reinterpret_cast<float4*>(hm.z_trace + ((int64_t)it * n + r) * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) tr[q] = make_float4(zc[4 * q], zc[4 * q + 1], zc[4 * q + 2], zc[4 * q + 3]);
            }
            if (hm.h1 != nullptr) hm.h1[(int64_t)it * n + r] = H;
            if (hm.alpha != nullptr) hm.alpha[(int64_t)it * n + r] = a;
            if (hm.moves != nullptr) hm.moves[(int64_t)it * n + r] = mv ? 1.f : 0.f;
          }
          moved_on = it + 1 < hm.n_iters;
          if (moved_on) {
            // the next iteration starts from the selected state, whose metric is already known: this
            // evaluation's if the move was accepted, the one this iteration started from if not
            if (!mv) {
              lad = st[34];
              ok = st[35] != 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(st + 36 + 4 * q);
                dg[4 * q] = v.x; dg[4 * q + 1] = v.y; dg[4 * q + 2] = v.z; dg[4 * q + 3] = v.w;
              }
            }
            begin_iteration = true;
            it_begin = it + 1;
          }
        }
      }
      if (begin_iteration) {
        // rho = gamma / beta_zero_sqrt (:123), H0 (:127), first half step (:132-138); remember the metric here
        st[34] = lad;
        st[35] = ok ? 1.f : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(st + 36 + 4 * q) = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) gm = __ldg(reinterpret_cast<const float4*>(hm.gamma + ((int64_t)it_begin * n + r) * 16) + q);
          const float g4[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int jj = 4 * q + e;
            const float rho = g4[e] / hm.b0;
            ss = fmaf(rho, rho, ss);
            const float g = -((1.f - lambda * dg[jj]) / hm.T2);
            rh[jj] = rho - half_eps * g;
            zc[jj] = zc[jj] + hm.eps * rh[jj];
          }
        }
        const float nrm = sqrtf(ss);
        h0_keep = -hmc_log_pi(lad, ok) + 0.5f * (nrm * nrm);
        if (live && hm.h0 != nullptr) hm.h0[(int64_t)it_begin * n + r] = h0_keep;
      }
      if (moved_on) {
        // publish the position of the next evaluation: state row (z, rho_half, exponent constants) for this
        // CTA's exp groups, z' = 2^ez (z - shift) split in fp16 into TMEM for GEMM1
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float4*>(st + 4 * q) = make_float4(zc[4 * q], zc[4 * q + 1], zc[4 * q + 2], zc[4 * q + 3]);
          *reinterpret_cast<float4*>(st + 16 + 4 * q) = make_float4(rh[4 * q], rh[4 * q + 1], rh[4 * q + 2], rh[4 * q + 3]);
        }
        float zs[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) zs[jj] = live ? zc[jj] : 0.f;
        if (!EXACT) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
            zs[4 * q] -= sh.x; zs[4 * q + 1] -= sh.y; zs[4 * q + 2] -= sh.z; zs[4 * q + 3] -= sh.w;
          }
        }
        float nrm2 = 0.f, zmax = 0.f;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) { nrm2 = fmaf(zs[jj], zs[jj], nrm2); zmax = fmaxf(zmax, fabsf(zs[jj])); }
        int ez = 0;
        if (zmax > 0.f && zmax < 3.0e38f) {
          const int ex = (int)((__float_as_uint(zmax) >> 23) & 0xffu) - 126;
          ez = 14 - ex;
          ez = ez > 50 ? 50 : (ez < -50 ? -50 : ez);
        }
        st[32] = EXACT ? P_SHIFT : (-nrm2 * alpha + P_SHIFT);
        st[33] = 2.f * alpha * c_unscale * __uint_as_float((uint32_t)(127 - ez) << 23);
        if (!EXACT) {
          const float zsc = __uint_as_float((uint32_t)(ez + 127) << 23);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) split_pair(zs[2 * jj] * zsc, zs[2 * jj + 1] * zsc, hi[jj], lo[jj]);
          TMEM_ST8(tmem_base + lane_addr + TM_ZHI, hi);
          TMEM_ST8(tmem_base + lane_addr + TM_ZLO, lo);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR_ZL);
          if (PAIR) mbar_arrive_leader(BAR_ZR); else mbar_arrive(BAR_ZR);
        }
#pragma unroll
        for (int i = 0; i < NCOLS; ++i) total[i] = 0.f;
      }
      continue;                        // next evaluation
    }
    // Outputs are stored straight from the registers of the thread that owns the point: 36 (packed) or 64
    // (expanded) fire-and-forget 128-bit stores per row, each row a contiguous 576 B / 1 KB run, so every
    // 32-byte sector is written completely (by two consecutive stores) and the stores drain while the
    // thread goes on with the Cholesky.  (Staging through shared memory + a cooperative or bulk copy
    // serialised ~10k cycles per CTA behind barriers: measured +0.9 / +0.65 ms per 2^20 points.)
    auto store_rows = [&](float* dst_base) {                    // total[] -> packed [rows, 144]
      if (live) {
        float4* dst = reinterpret_cast<float4*>(dst_base + (row0 + prow) * NCOLS);
#pragma unroll
        for (int q = 0; q < NCOLS / 4; ++q)
          dst[q] = make_float4(total[4 * q], total[4 * q + 1], total[4 * q + 2], total[4 * q + 3]);
      }
    };
    if (fo.a_packed != nullptr) store_rows(fo.a_packed);
    if (fo.a_full != nullptr && live) {
      float4* dst = reinterpret_cast<float4*>(fo.a_full + (row0 + prow) * 256);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            v[e] = total[i <= j ? sym_index(i, j) : sym_index(j, i)];
          }
          dst[i * 4 + q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    const bool fused = fo.g_packed != nullptr || fo.g_full != nullptr || fo.logabsdet != nullptr ||
                       fo.sign != nullptr || fo.diag_g != nullptr;
    if (fused) {
      if (!live) {
#pragma unroll
        for (int i = 0; i < 136; ++i) total[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) total[sym_index(i, i)] = 1.f;
      }
      float lad, dg[16];
      const bool ok = sym16_factor(total, lad, dg, fo.g_packed != nullptr || fo.g_full != nullptr);
      if (live) {
        const int64_t r = row0 + prow;
        if (fo.logabsdet != nullptr) fo.logabsdet[r] = fo.lad_scale * lad;
        if (fo.sign != nullptr) fo.sign[r] = 1.f;
        if (fo.diag_g != nullptr) {
          float4* dst = reinterpret_cast<float4*>(fo.diag_g + r * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
        }
        if (!ok) fo.fail_ws[1 + atomicAdd(fo.fail_ws, 1)] = (int)r;
      }
      if (fo.g_packed != nullptr) {
#pragma unroll
        for (int i = 136; i < NCOLS; ++i) total[i] = 0.f;
        store_rows(fo.g_packed);
      }
      if (fo.g_full != nullptr && live) {
        float4* dst = reinterpret_cast<float4*>(fo.g_full + (row0 + prow) * 256);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q + e;
              v[e] = total[i <= j ? sym_index(i, j) : sym_index(j, i)];
            }
            dst[i * 4 + q] = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
    }
#ifdef RLVAE_TC_PROFILE
    if (!HMC && (blockIdx.x == 0 || blockIdx.x == 4096) && threadIdx.x == 384)
    {
      const long long pf2 = clock64();
      printf("[h16 prof %d] fold group per chunk: wait CH_FULL %lld  fold %lld | prologue %lld  mainloop %lld  epilogue %lld cycles\n",
             (int)blockIdx.x, pf_wait / num_chunks, pf_work / num_chunks, pf0 - pk0, pf1 - pf0, pf2 - pf1);
    }
#endif
    }   // evaluations (one, unless HMC)
  }
#undef MMA_H
#undef MMA_TS
#undef COMMIT

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// HYBRID mode of the gradient kernel.  (Measured against the alternatives of the forward kernel -- natural rows staged in
// shared memory, t_k through a shared-memory row, refinement deferred past the hand-off of u: same-box A/B 11.6 ms for
// this version, 12.0 ms with all three, 12.9-13.4 ms with subsets.  Near pairs are rare per (warp, block) when every point
// sits next to a few centroids, and the unconditional per-block cost of the staging outweighs cheaper rounds.)
// Like refine_exponents, the pairs flagged in `live` get their exponent from
// exact differences -- and, because those are exactly the pairs whose weight is not negligible, i.e. the centroids
// NEAR the point, their whole contribution u (c_k - z) is accumulated here from the exact difference vector and
// taken out of the tensor contraction (exponent -> -1e30 -> u = 0 for GEMM3 and for sum_k u).  The contraction's
// form sum_k u c~_k - z~ sum_k u loses |c~| / |c_k - z| digits next to a centroid (1.2e-4 relative at T = 0.1,
// measured against fp64); with the near pairs handled here only far pairs go through it, where nothing cancels.
__device__ __forceinline__ void refine_direct(uint32_t (&ex)[32], const uint32_t (&tv)[32], uint32_t live,
                                              const float4* __restrict__ crows, const float2 (&nz)[8], float neg_alpha,
                                              float shift, float2 (&direct)[8]) {
  if (!__any_sync(0xffffffffu, live != 0u)) return;
  constexpr uint32_t NEG_BIG = 0xf149f2cau;        // -1e30f
  const int total = __reduce_add_sync(0xffffffffu, __popc(live));
  if (total > 160) {                               // dense tile: uniform sweep, broadcast loads
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float2 dv[8];
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 c = __ldg(crows + i * 4 + q);
        dv[2 * q] = fadd2(make_float2(c.x, c.y), nz[2 * q]);
        dv[2 * q + 1] = fadd2(make_float2(c.z, c.w), nz[2 * q + 1]);
        ffma2_acc(acc, dv[2 * q], dv[2 * q]);
        ffma2_acc(acc, dv[2 * q + 1], dv[2 * q + 1]);
      }
      if ((live >> i) & 1u) {
        const float uv = ex2_approx(fmaf(acc.x + acc.y, neg_alpha, shift)) * __uint_as_float(tv[i]);
        const float2 u2 = make_float2(uv, uv);
#pragma unroll
        for (int q = 0; q < 8; ++q) ffma2_acc(direct[q], u2, dv[q]);
        ex[i] = NEG_BIG;
      }
    }
    return;
  }
  while (__any_sync(0xffffffffu, live != 0u)) {
    if (live != 0u) {
      const int b = __ffs(live) - 1;
      live &= live - 1u;
      float2 dv[8];
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 c = __ldg(crows + b * 4 + q);
        dv[2 * q] = fadd2(make_float2(c.x, c.y), nz[2 * q]);
        dv[2 * q + 1] = fadd2(make_float2(c.z, c.w), nz[2 * q + 1]);
        ffma2_acc(acc, dv[2 * q], dv[2 * q]);
        ffma2_acc(acc, dv[2 * q + 1], dv[2 * q + 1]);
      }
      uint32_t tb = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) tb = (i == b) ? tv[i] : tb;
      const float uv = ex2_approx(fmaf(acc.x + acc.y, neg_alpha, shift)) * __uint_as_float(tb);
      const float2 u2 = make_float2(uv, uv);
#pragma unroll
      for (int q = 0; q < 8; ++q) ffma2_acc(direct[q], u2, dv[q]);
#pragma unroll
      for (int i = 0; i < 32; ++i) ex[i] = (i == b) ? NEG_BIG : ex[i];
    }
  }
}

// ==========================================================================================
// Gradient / backward kernel, split-fp16 version (symmetric tables, d == 16):
//   out[n,:] = scale * sum_k w_nk <U_n, M_k> (c_k - z_n)
// Per 64-centroid super-block j (ONE CTA contracts all 136 packed columns, so nothing is duplicated):
//   GEMM1   S[128 x 64]  = Z.C^T                                   (3xTF32, as everywhere)
//   T-GEMM  T[128 x 64]  = U'_hi.M'_hi^T + U'_lo.M'_hi^T + U'_hi.M'_lo^T   (kind::f16, K = 144 = 9 steps)
//           U' = 2^eU Ut (per point, max|U'| in [2^13, 2^14)), Ut_p = U_ij + U_ji (i<j), U_ii: valid for
//           ANY U; M' = 2^eM M as in the forward kernel.  U' (hi|lo fp16) is resident in TMEM.
//   exp     u = exp2(..) * t, split u_hi | u_lo (fp32 / TF32), written over S | T
//   GEMM3   OUT[128 x 16] += u_hi.Ct_hi + u_lo.Ct_hi + u_hi.Ct_lo   (3xTF32, N = 16; Ct = c^T; sum_k u is
//           accumulated by the exp threads themselves: one FADD per element instead of 16 more columns)
// OUT accumulates in two alternating chunk accumulators (2 super-blocks each) that the exp groups
// fold into fp32 registers.  out = scale 2^-(eU+eM) (OUT - z sum_k u).
// TMEM: [0,72) U'_hi, [72,144) U'_lo, [144,400) two (S|u_hi 64, T|u_lo 64) buffers, [400,464) OUT x 2,
//       [464,496) z_hi | z_lo (TF32 split, A operand of GEMM1).
// ==========================================================================================
namespace g16 {
constexpr int THREADS = 384;        // TMA warp (C), MMA warp, 2 exp groups of 4 warps, TMA warps (M, Ct)
constexpr int C_STAGES = 4;
constexpr int M_STAGES = 2;                          // one stage = hi AND lo tile of a super-block
constexpr int KSTEPS = 9;                           // 144 packed columns / 16
constexpr bool COLSPLIT = true;                     // exp groups split every super-block by columns (see the exp loop)
constexpr uint32_t CT_TILE_BYTES = 2 * 16 * 128;    // [16 rows (c^T) x 64 centroids] fp32 = 2 atoms of 32 centroids
constexpr uint32_t M_HALF_BYTES = 3 * BK * 128;     // 3 column atoms (64 fp16) x 64 centroid rows
constexpr uint32_t M_TILE_BYTES = 2 * M_HALF_BYTES;
constexpr uint32_t OFF_C = 0;                       // (z lives in TMEM: no A tiles in shared memory)
constexpr uint32_t OFF_CT = OFF_C + C_STAGES * C_TILE_BYTES;
constexpr uint32_t OFF_M = OFF_CT + C_STAGES * 2 * CT_TILE_BYTES;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 5 * C_STAGES + 2 * M_STAGES + 11;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
constexpr uint32_t OFF_HYB = (OFF_TMEM_PTR + 16 + 15) & ~15u;   // HYBRID only: one scratch row per exp warp (8 warps)
constexpr uint32_t SMEM_BYTES_HYB = OFF_HYB + 8 * HYB_ROW_BYTES + 1024;   // (rows: the unit-weight mode's refine_exponents)
static_assert(OFF_HYB % 16 == 0 && SMEM_BYTES_HYB <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_UHI = 0, TM_ULO = 72, TM_ST = 144, TM_OUT = 400, TM_ZHI = 464, TM_ZLO = 480;
constexpr int RED_LD = 20;
}  // namespace g16

__device__ __forceinline__ void split_pair_scaled(float even, float odd, float sc, uint32_t& hi2, uint32_t& lo2) {
  split_pair(even * sc, odd * sc, hi2, lo2);
}

// EXACT: weights from exact differences on the FMA pipe (see inverse_metric_h16_kernel); no GEMM1.
// HYBRID: as in the forward kernel -- weights above hyb_thr are recomputed from exact differences.
template <bool PAIR, bool EXACT, bool HYBRID>
__global__ void __launch_bounds__(g16::THREADS, 1)
metric_grad_h16_kernel(const __grid_constant__ CUtensorMap tm_cstack,
                       const __grid_constant__ CUtensorMap tm_mn_hi,
                       const __grid_constant__ CUtensorMap tm_mn_lo,
                       const __grid_constant__ CUtensorMap tm_ct_hi,
                       const __grid_constant__ CUtensorMap tm_ct_lo,
                       const float* __restrict__ z, const float* __restrict__ u,
                       const float* __restrict__ cbias /* EXACT: 0 / -1e30 mask */,
                       const float* __restrict__ cnat /* EXACT: natural centroid rows */, int64_t n, int num_blocks,
                       float alpha, float scale /* includes 2^-eM */, float c_unscale /* 2^-ec */,
                       const float* __restrict__ cshift /* [16] centre of the expanded form */,
                       float hyb_thr /* HYBRID: refine weights with log2 w above this */,
                       float* __restrict__ out, int u_packed) {
  constexpr int C_STAGES = g16::C_STAGES, M_STAGES = g16::M_STAGES, RED_LD = g16::RED_LD, KSTEPS = g16::KSTEPS;
  constexpr bool COLSPLIT = g16::COLSPLIT;
  constexpr uint32_t CT_TILE_BYTES = g16::CT_TILE_BYTES, M_TILE_BYTES = g16::M_TILE_BYTES,
                     M_HALF_BYTES = g16::M_HALF_BYTES, OFF_C = g16::OFF_C,
                     OFF_CT = g16::OFF_CT, OFF_M = g16::OFF_M, OFF_BIAS = g16::OFF_BIAS,
                     TM_UHI = g16::TM_UHI, TM_ULO = g16::TM_ULO, TM_ST = g16::TM_ST, TM_OUT = g16::TM_OUT,
                     TM_ZHI = g16::TM_ZHI, TM_ZLO = g16::TM_ZLO;
  constexpr int CHUNK = 2;             // super-blocks per OUT chunk accumulator
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + g16::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_B_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };     // bias (local)
  auto BAR_CT_FULL = [&](int s) { return bar0 + 8u * (3 * C_STAGES + s); };
  auto BAR_CT_EMPTY = [&](int s) { return bar0 + 8u * (4 * C_STAGES + s); };
  auto BAR_M_FULL = [&](int s) { return bar0 + 8u * (5 * C_STAGES + s); };
  auto BAR_M_EMPTY = [&](int s) { return bar0 + 8u * (5 * C_STAGES + M_STAGES + s); };
  auto BAR_ST_FULL = [&](int b) { return bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + b); };
  auto BAR_U_FULL = [&](int b) { return bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + 2 + b); };
  auto BAR_CH_FULL = [&](int b) { return bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + 4 + b); };
  auto BAR_CH_FREE = [&](int b) { return bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + 6 + b); };
  const uint32_t BAR_DONE = bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + 8);
  auto BAR_G3_DONE = [&](int b) { return bar0 + 8u * (5 * C_STAGES + 2 * M_STAGES + 9 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + g16::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  constexpr uint32_t ROWS_CTA = PAIR ? BK / 2 : BK;
  constexpr uint32_t ATOM_BYTES = ROWS_CTA * 128;                   // M tile atom held by this CTA
  constexpr uint32_t TILE_BYTES = 3 * ATOM_BYTES;
  constexpr uint32_t ATOM_DESC = ATOM_BYTES >> 4;
  constexpr uint32_t CT_ROWS = PAIR ? 8 : 16;
  constexpr uint32_t CT_ATOM_BYTES = CT_ROWS * 128;
  constexpr uint32_t CT_BYTES = 2 * CT_ATOM_BYTES;                  // per hi / lo tile per CTA
  constexpr uint32_t CT_ATOM_DESC = CT_ATOM_BYTES >> 4;
  constexpr uint32_t IDESC_T = make_idesc_f16(PAIR ? 256 : 128, BK);
  constexpr uint32_t IDESC_3 = make_idesc(PAIR ? 256 : 128, 16);   // N = 16: the 16 components of sum_k u c_k
  const int num_chunks = (num_blocks + CHUNK - 1) / CHUNK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), COLSPLIT ? 8 : 4); mbar_init(BAR_B_FULL(s), 1);
      mbar_init(BAR_CT_FULL(s), 1); mbar_init(BAR_CT_EMPTY(s), 1);
    }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR_ST_FULL(b), 1); mbar_init(BAR_U_FULL(b), (COLSPLIT ? 8 : 4) * NPAIR);
      mbar_init(BAR_CH_FULL(b), 1); mbar_init(BAR_CH_FREE(b), 4 * NPAIR);
    }
    mbar_init(BAR_DONE, 1);
    for (int b = 0; b < 2; ++b) mbar_init(BAR_G3_DONE(b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cstack) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ct_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ct_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + g16::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + g16::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();            // TMEM base published before the exp threads store U into it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int quarter = warp & 3;
  const int prow = quarter * 32 + lane;
  const int grp = (warp >= 6) ? 1 : 0;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
  float zb = 0.f, s_scale = 0.f;   // exponent = S' * s_scale + bias_k + zb
  float u_unscale = 1.f;      // 2^-eU
  float zrow[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) zrow[j] = 0.f;
  if (warp >= 2 && warp < 10) {
    const int64_t r = row0 + prow;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = __ldg(src + q);
        zrow[4 * q] = v.x; zrow[4 * q + 1] = v.y; zrow[4 * q + 2] = v.z; zrow[4 * q + 3] = v.w;
      }
    }
    if (!EXACT) {   // GEMM1 on kind::f16 about the table's centre: z~ = z - shift, z' = 2^ez z~ (max|z'| in
                    // [2^13, 2^14)), split hi + lo; S = 2^-(ez+ec) S'
      float zs[16];
      float nrm = 0.f, zmax = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
        zs[4 * q] = zrow[4 * q] - sh.x; zs[4 * q + 1] = zrow[4 * q + 1] - sh.y;
        zs[4 * q + 2] = zrow[4 * q + 2] - sh.z; zs[4 * q + 3] = zrow[4 * q + 3] - sh.w;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { nrm = fmaf(zs[j], zs[j], nrm); zmax = fmaxf(zmax, fabsf(zs[j])); }
      zb = -nrm * alpha;
      int ez = 0;
      if (zmax > 0.f && zmax < 3.0e38f) {
        const int ex = (int)((__float_as_uint(zmax) >> 23) & 0xffu) - 126;
        ez = 14 - ex;
        ez = ez > 50 ? 50 : (ez < -50 ? -50 : ez);
      }
      s_scale = 2.f * alpha * c_unscale * __uint_as_float((uint32_t)(127 - ez) << 23);
      if (grp == 0) {
        const float zsc = __uint_as_float((uint32_t)(ez + 127) << 23);
        uint32_t zh[8], zl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_pair(zs[2 * j] * zsc, zs[2 * j + 1] * zsc, zh[j], zl[j]);
        TMEM_ST8(tmem_base + lane_addr + TM_ZHI, zh);
        TMEM_ST8(tmem_base + lane_addr + TM_ZLO, zl);
      }
    }
    // ---- U' = 2^eU Ut, split into fp16 hi | lo, resident in TMEM as the A operand of the T GEMM.
    // group 0 converts packed columns [0,64), group 1 [64,144); both scan the whole row for the scale.
    // unit mode (u_packed == 2, pythae variant): U' = e_136 against the table's constant pad column (= 1), so
    // t_k = 1 and the kernel returns scale * sum_k w_k b_k for the table behind tm_ct_* (no "- z sum_k u" term)
    const bool unit = u_packed == 2;
    const int row_len = u_packed ? SYM_COLS : NCOL;
    const float* urow = u + r * row_len;
    float m = 0.f;
    if (r < n && !unit) {
      const float4* u4 = reinterpret_cast<const float4*>(urow);
      const int nq = u_packed ? 34 : 64;
      for (int q = 0; q < nq; ++q) {
        const float4 v = __ldg(u4 + q);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
      m *= 2.f;                                     // |Ut_p| <= 2 max|U|
    }
    int eu = 0;
    if (m > 0.f && m < 3.0e38f) {
      const int ex = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126;    // m = f 2^ex, f in [0.5, 1)
      eu = 14 - ex;
      eu = eu > 50 ? 50 : (eu < -50 ? -50 : eu);
    }
    const float usc = __uint_as_float((uint32_t)(eu + 127) << 23);
    u_unscale = __uint_as_float((uint32_t)(127 - eu) << 23);
    {
      uint32_t h[32], l[32], h2[8], l2[8];
      int p = grp * 64;
      int ri = 0, rb = 0;
      while (ri < 15 && p >= rb + (16 - ri)) { rb += 16 - ri; ++ri; }
      int cj = ri + (p - rb);
      auto next = [&]() -> float {      // Ut at packed index p, then advance
        float v = 0.f;
        if (unit) {
          v = (p == 136) ? 1.f : 0.f;
        } else if (r < n && p < 136) {
          if (u_packed) {
            v = __ldg(urow + p);
            if (cj != ri) v += v;
          } else {
            v = __ldg(urow + ri * 16 + cj);
            if (cj != ri) v += __ldg(urow + cj * 16 + ri);
          }
        }
        ++p; ++cj;
        if (cj == 16) { ++ri; cj = ri; }
        return v;
      };
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e0 = next();
        const float e1 = next();
        split_pair_scaled(e0, e1, usc, h[i], l[i]);
      }
      TMEM_ST32(tmem_base + lane_addr + TM_UHI + grp * 32, h);
      TMEM_ST32(tmem_base + lane_addr + TM_ULO + grp * 32, l);
      if (grp == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e0 = next();
          const float e1 = next();
          split_pair_scaled(e0, e1, usc, h2[i], l2[i]);
        }
        TMEM_ST8(tmem_base + lane_addr + TM_UHI + 64, h2);
        TMEM_ST8(tmem_base + lane_addr + TM_ULO + 64, l2);
      }
    }
    tmem_wait_st();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_T16(d, a, b, acc) do { if (PAIR) mma_ts_f16_pair(d, a, b, IDESC_T, acc); else mma_ts_f16(d, a, b, IDESC_T, acc); } while (0)
#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (warp == 0) {
    // =========================================================== TMA producer 1: centroid tiles + bias,
    // then the Ct tiles (hi, lo) of the same block: [32 (pair: 16) rows x 64 centroids] as two
    // 32-centroid atoms each
    for (int j = 0; j < num_blocks; ++j) {
      const int cs = j % C_STAGES;
      mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t dst = base + OFF_C + cs * C_TILE_BYTES;
        if (EXACT) {        // natural rows of the 64 centroids (4 KB) + mask for this CTA's exp groups
          mbar_expect_tx(BAR_B_FULL(cs), BIAS_BYTES + BK * 64);
          bulk_load_1d(dst, cnat + (int64_t)j * BK * 16, BK * 64, BAR_B_FULL(cs));
        } else {
          if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE_BYTES);
          if (PAIR) {
            tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32 * (int)rank);
          } else {
            tma_load_2d(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK);
            tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32);
          }
          mbar_expect_tx(BAR_B_FULL(cs), BIAS_BYTES);
        }
        bulk_load_1d(base + OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_B_FULL(cs));
      }
      __syncwarp();
      mbar_wait(BAR_CT_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_CT_FULL(cs), NPAIR * 2 * CT_BYTES);
        const uint32_t dst = base + OFF_CT + cs * 2 * CT_TILE_BYTES;
        const int row = PAIR ? 8 * (int)rank : 0;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (PAIR) {
            tma_load_2d_pair(dst + a * CT_ATOM_BYTES, &tm_ct_hi, BAR_CT_FULL(cs), j * BK + 32 * a, row);
            tma_load_2d_pair(dst + CT_TILE_BYTES + a * CT_ATOM_BYTES, &tm_ct_lo, BAR_CT_FULL(cs), j * BK + 32 * a, row);
          } else {
            tma_load_2d(dst + a * CT_ATOM_BYTES, &tm_ct_hi, BAR_CT_FULL(cs), j * BK + 32 * a, row);
            tma_load_2d(dst + CT_TILE_BYTES + a * CT_ATOM_BYTES, &tm_ct_lo, BAR_CT_FULL(cs), j * BK + 32 * a, row);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // =========================================================== TMA producer 2: M tiles (hi, lo alternate)
    for (int j = 0; j < num_blocks; ++j) {
      const int ms = j % M_STAGES;
      mbar_wait(BAR_M_EMPTY(ms), ((j / M_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * 2 * TILE_BYTES);
        const uint32_t dst = base + OFF_M + ms * M_TILE_BYTES;
        const int row = j * BK + (PAIR ? 32 * (int)rank : 0);
        if (PAIR) {
          tma_load_3d_pair(dst, &tm_mn_hi, BAR_M_FULL(ms), 0, row, 0);
          tma_load_3d_pair(dst + M_HALF_BYTES, &tm_mn_lo, BAR_M_FULL(ms), 0, row, 0);
        } else {
          tma_load_3d(dst, &tm_mn_hi, BAR_M_FULL(ms), 0, row, 0);
          tma_load_3d(dst + M_HALF_BYTES, &tm_mn_lo, BAR_M_FULL(ms), 0, row, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer 1 (pair: leader only):
    // [GEMM1, T-GEMM] of every super-block.  The final contraction (GEMM3) is issued by a SECOND warp
    // (warp 11): with one issuer the 57 MMAs per super-block made the kernel issue-bound (measured:
    // ~26 cycles of issue per MMA against 1.44k cycles of tensor time).  The two streams touch disjoint
    // TMEM regions; their only ordering constraint -- S|T(j+2) overwrites the buffer GEMM3(j) reads --
    // is the G3_DONE barrier.  Unrolled 4x: stage indices and parities are compile-time constants.
    if (leader) {
      static_assert(C_STAGES == 4 && M_STAGES == 2 && CHUNK == 2, "the 4x unrolled issue loops assume these periods");
      const uint64_t c_desc0 = make_desc_sw128(base + OFF_C);
      const uint64_t m_desc0 = make_desc_sw128(base + OFF_M);
      constexpr uint32_t ID1 = make_idesc_f16(PAIR ? 256 : 128, BK);
      auto issue_st = [&](auto Jc, const int j, const uint32_t qodd /* (j / 4) & 1 */) {
        constexpr int J = decltype(Jc)::value;
        constexpr int cs = J % C_STAGES, sb = J & 1, ms = J % M_STAGES;
        if (!EXACT) mbar_wait(BAR_C_FULL(cs), qodd);
        mbar_wait(BAR_M_FULL(ms), (J / M_STAGES) & 1);
        if (j >= 2) mbar_wait(BAR_G3_DONE(sb), ((J >> 1) + 1) & 1);     // GEMM3(j-2) has consumed this buffer
        tc_fence_after();
        const uint32_t s_t = tmem_base + TM_ST + sb * 128;
        const uint32_t t_t = s_t + 64;
        const uint64_t bc = c_desc0 + ((cs * C_TILE_BYTES) >> 4);
        const uint64_t bh = m_desc0 + ((ms * M_TILE_BYTES) >> 4);
        const uint64_t bl = bh + (M_HALF_BYTES >> 4);
        if (elect_one()) {
          if (!EXACT) {     // S' = z'_hi.c'_hi + z'_hi.c'_lo + z'_lo.c'_hi (kind::f16, K = 16: one MMA each)
            if (PAIR) {
              mma_ts_f16_pair(s_t, tmem_base + TM_ZHI, bc, ID1, 0);
              mma_ts_f16_pair(s_t, tmem_base + TM_ZHI, bc + 2, ID1, 1);
              mma_ts_f16_pair(s_t, tmem_base + TM_ZLO, bc, ID1, 1);
            } else {
              mma_ts_f16(s_t, tmem_base + TM_ZHI, bc, ID1, 0);
              mma_ts_f16(s_t, tmem_base + TM_ZHI, bc + 2, ID1, 1);
              mma_ts_f16(s_t, tmem_base + TM_ZLO, bc, ID1, 1);
            }
          }
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_T16(t_t, tmem_base + TM_UHI + 8 * kk, bh + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), kk > 0);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_T16(t_t, tmem_base + TM_ULO + 8 * kk, bh + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), 1);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_T16(t_t, tmem_base + TM_UHI + 8 * kk, bl + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), 1);
          COMMIT(BAR_M_EMPTY(ms));
          COMMIT(BAR_ST_FULL(sb));
        }
        __syncwarp();
      };
      uint32_t qodd = 0;
      for (int j0 = 0; j0 < num_blocks; j0 += 4, qodd ^= 1u) {
#define RLVAE_BLK(J) if (j0 + J < num_blocks) issue_st(std::integral_constant<int, J>{}, j0 + J, qodd);
        RLVAE_BLK(0) RLVAE_BLK(1) RLVAE_BLK(2) RLVAE_BLK(3)
#undef RLVAE_BLK
      }
    }
  } else if (warp == 11) {
    // =========================================================== MMA issuer 2 (pair: leader only): GEMM3
    if (leader) {
      const uint64_t ct_desc0 = make_desc_sw128(base + OFF_CT);
      uint32_t free_phase = 0;     // bit b: parity of the next CH_FREE(b) completion to wait for
      auto issue_g3 = [&](auto Jc, const int j, const uint32_t qodd /* (j / 4) & 1 */) {
        constexpr int J = decltype(Jc)::value;
        constexpr int cs = J % C_STAGES, sb = J & 1, cb = (J >> 1) & 1;
        constexpr int first = (J % CHUNK) == 0;
        if (first && j >= 2 * CHUNK) {
          mbar_wait(BAR_CH_FREE(cb), (free_phase >> cb) & 1u);
          free_phase ^= 1u << cb;
        }
        mbar_wait(BAR_CT_FULL(cs), qodd);
        mbar_wait(BAR_U_FULL(sb), (J >> 1) & 1);
        tc_fence_after();
        const uint32_t u_hi = tmem_base + TM_ST + sb * 128;
        const uint32_t u_lo = u_hi + 64;
        const uint32_t acc = tmem_base + TM_OUT + cb * 32;
        const uint64_t ch = ct_desc0 + ((cs * 2 * CT_TILE_BYTES) >> 4);
        const uint64_t cl = ch + (CT_TILE_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, u_hi + 8 * kk, ch + (kk >> 2) * CT_ATOM_DESC + 2 * (kk & 3), IDESC_3, !(first && kk == 0));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, u_lo + 8 * kk, ch + (kk >> 2) * CT_ATOM_DESC + 2 * (kk & 3), IDESC_3, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, u_hi + 8 * kk, cl + (kk >> 2) * CT_ATOM_DESC + 2 * (kk & 3), IDESC_3, 1);
          COMMIT(BAR_CT_EMPTY(cs));
          COMMIT(BAR_G3_DONE(sb));
          if ((J % CHUNK) == CHUNK - 1 || j == num_blocks - 1) COMMIT(BAR_CH_FULL(cb));
        }
        __syncwarp();
      };
      uint32_t qodd = 0;
      for (int j0 = 0; j0 < num_blocks; j0 += 4, qodd ^= 1u) {
#define RLVAE_BLK(J) if (j0 + J < num_blocks) issue_g3(std::integral_constant<int, J>{}, j0 + J, qodd);
        RLVAE_BLK(0) RLVAE_BLK(1) RLVAE_BLK(2) RLVAE_BLK(3)
#undef RLVAE_BLK
      }
      if (elect_one()) COMMIT(BAR_DONE);
      __syncwarp();
    }
  } else {
    // =========================================================== exp groups (one thread per point)
    float2 nz[8];                        // EXACT: the negated point
#pragma unroll
    for (int q = 0; q < 8; ++q) nz[q] = make_float2(-zrow[2 * q], -zrow[2 * q + 1]);
    float tot[17];                       // this group's share of OUT (chunks of parity grp) and of sum_k u
#pragma unroll
    for (int e = 0; e < 17; ++e) tot[e] = 0.f;
    float2 direct[8];                    // HYBRID: sum of u (c_k - z) over the near pairs, from exact differences
#pragma unroll
    for (int q = 0; q < 8; ++q) direct[q] = make_float2(0.f, 0.f);
    float* hyb_row = reinterpret_cast<float*>(gbase + g16::OFF_HYB + (warp - 2) * HYB_ROW_BYTES) + lane * 4;
    (void)hyb_row;
    auto fold_chunk = [&](int c, bool signal) {
      uint32_t a[16];
      TMEM_LD16(tmem_base + lane_addr + TM_OUT + (c & 1) * 32, a);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) tot[i] += __uint_as_float(a[i]);
      if (signal) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE(c & 1)); else mbar_arrive(BAR_CH_FREE(c & 1)); }
      }
    };
    int next_chunk = grp;                // chunks c with (c & 1) == grp belong to this group
    long long pe_wait = 0, pe_work = 0, pe_fold = 0;
    (void)pe_wait; (void)pe_work; (void)pe_fold;
    // COLSPLIT: both groups work on EVERY super-block, 32 centroids each, instead of alternating whole
    // super-blocks.  The tensor pipe runs in order (T(j+1), GEMM3(j), T(j+2), ...), so u(j) has to be
    // ready within the ~960 cycles of T(j+1): halving the exp latency per super-block removes that stall.
    for (int j = COLSPLIT ? 0 : grp; j < num_blocks; j += COLSPLIT ? 1 : 2) {
      PROF_T0();
      const int cs = j % C_STAGES, sb = j & 1;
      const uint32_t st = tmem_base + lane_addr + TM_ST + sb * 128;
      mbar_wait(BAR_B_FULL(cs), (j / C_STAGES) & 1);
      mbar_wait(BAR_ST_FULL(sb), (j >> 1) & 1);
      tc_fence_after();
      PROF_ADD(pe_wait);
      float su_blk = 0.f;                // sum of u over this super-block (two-level fp32 summation)
#pragma unroll
      for (int rr = 0; rr < (COLSPLIT ? 1 : 2); ++rr) {
        const int rnd = COLSPLIT ? grp : rr;
        uint32_t sv[32], tv[32];
        if (!EXACT) TMEM_LD32(st + rnd * 32, sv);
        TMEM_LD32(st + 64 + rnd * 32, tv);
        const float4* bias4 = reinterpret_cast<const float4*>(gbase + OFF_BIAS + cs * BIAS_BYTES) + rnd * 8;
        const float4* crow = reinterpret_cast<const float4*>(gbase + OFF_C + cs * C_TILE_BYTES) + rnd * 32 * 4;
        (void)crow;
        tmem_wait_ld();
        if (HYBRID) {
          uint32_t live = 0;               // weights of this thread that need exact differences
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bv = bias4[q];
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q + e;
              const float ex = fmaf(__uint_as_float(sv[i]), s_scale, b4[e] + zb);
              live |= (ex > hyb_thr ? 1u : 0u) << i;
              sv[i] = __float_as_uint(ex);
            }
          }
          if (u_packed == 2)      // unit-weight mode: the table behind the contraction is not the centroids
            refine_exponents<false>(sv, live, reinterpret_cast<const float4*>(cnat) + ((int64_t)j * BK + rnd * 32) * 4,
                                    nz, -alpha, 0.f, hyb_row);
          else
            refine_direct(sv, tv, live, reinterpret_cast<const float4*>(cnat) + ((int64_t)j * BK + rnd * 32) * 4, nz,
                          -alpha, 0.f, direct);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float uv = ex2_approx(__uint_as_float(sv[i])) * __uint_as_float(tv[i]);
            su_blk += uv;
            const uint32_t uh = __float_as_uint(uv) & 0xFFFFE000u;
            sv[i] = uh;
            tv[i] = __float_as_uint(uv - __uint_as_float(uh));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bv = bias4[q];
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q + e;
              const float w = EXACT ? ex2_approx(fmaf(dist2_row16(crow + i * 4, nz), -alpha, b4[e]))
                                    : ex2_approx(fmaf(__uint_as_float(sv[i]), s_scale, b4[e] + zb));
              const float uv = w * __uint_as_float(tv[i]);
              su_blk += uv;
              const uint32_t uh = __float_as_uint(uv) & 0xFFFFE000u;
              sv[i] = uh;
              tv[i] = __float_as_uint(uv - __uint_as_float(uh));
            }
          }
        }
        TMEM_ST32(st + rnd * 32, sv);
        TMEM_ST32(st + 64 + rnd * 32, tv);
      }
      tot[16] += su_blk;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_U_FULL(sb)); else mbar_arrive(BAR_U_FULL(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
      PROF_ADD(pe_work);
      while (next_chunk < num_chunks && min((next_chunk + 1) * CHUNK - 1, num_blocks - 1) <= j - 1) {
        mbar_wait(BAR_CH_FULL(next_chunk & 1), (next_chunk >> 1) & 1);
        tc_fence_after();
        fold_chunk(next_chunk, true);
        next_chunk += 2;
      }
      PROF_ADD(pe_fold);
    }
#ifdef RLVAE_TC_PROFILE
    if ((blockIdx.x == 0 || blockIdx.x == 4096) && threadIdx.x == 64)
      printf("[g16 prof] exp group A per own super-block: wait S|T %lld  work %lld  fold %lld\n",
             pe_wait / (num_blocks / 2), pe_work / (num_blocks / 2), pe_fold / (num_blocks / 2));
#endif
    mbar_wait(BAR_DONE, 0);
    tc_fence_after();
    while (next_chunk < num_chunks) { fold_chunk(next_chunk, false); next_chunk += 2; }
    if (HYBRID) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { tot[2 * q] += direct[q].x; tot[2 * q + 1] += direct[q].y; }
    }
    // ---------------------------------------------------------- combine the two groups
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* red = reinterpret_cast<float*>(gbase + OFF_M);
    if (grp == 1) {
#pragma unroll
      for (int e = 0; e < 17; ++e) red[prow * RED_LD + e] = tot[e];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0) {
      const int64_t r = row0 + prow;
      const float su = tot[16] + red[prow * RED_LD + 16];
      if (r < n) {
        float o[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float ge = tot[e] + red[prow * RED_LD + e];
          const float zt = (u_packed == 2) ? 0.f : zrow[e] - __ldg(cshift + e);   // Ct holds c - shift
          o[e] = (ge - zt * su) * u_unscale * scale;
        }
        float4* dst = reinterpret_cast<float4*>(out + r * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
  }
#undef MMA_T16
#undef MMA_TS
#undef COMMIT

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ==========================================================================================
// Two nearest centroids on the tensor core (d == 16).
//   ref src/models/samplers/riemannian_sampler.py:58-67,125-131: dist = norm(mu - c) [N,K], topk(2, smallest)
// The [N,K] distance matrix is GEMM1 of the kernels above (S = Z.C^T, 3xTF32, z resident in TMEM);
// the scan groups turn S into keys ||c_k||^2 - 2 S_k (same order as the distances) and keep the FOUR
// smallest per point.  The expanded form is only a pre-filter: its ~1e-6 |z||c| absolute error could
// swap near-ties, so the final two are chosen among the 8 candidates (4 per scan group) by the
// EXACT differences sum_j (mu_j - c_kj)^2, accumulated exactly like nearest2_kernel (same value, same
// lowest-index tie rule) -- indices are bit-identical to the direct kernel / the reference's topk.
// One stage index for everything: C tile + ||c||^2 slice (TMA) -> S buffer (TMEM) -> scan -> FREE.
// ==========================================================================================
namespace n2 {
constexpr int THREADS = 320;         // TMA warp, MMA warp, two scan groups of 4 warps
constexpr int STAGES = 4;
constexpr uint32_t CN_BYTES = BK * 4;
constexpr uint32_t OFF_C = 0;
constexpr uint32_t OFF_CN = OFF_C + STAGES * C_TILE_BYTES;
constexpr uint32_t OFF_MERGE = OFF_CN + STAGES * CN_BYTES;         // [128 rows][8] (key, idx) of group B
constexpr uint32_t OFF_BAR = OFF_MERGE + TILE_M * 8 * 4;
constexpr int NUM_BARS = 4 * STAGES;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
constexpr uint32_t TM_S = 0;         // + stage*64
constexpr uint32_t TM_ZHI = 256, TM_ZLO = 272;
}  // namespace n2


// order-preserving float -> int map (so that keys can be ranked with integer min / max)
__device__ __forceinline__ int ordered_int(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
// sorted insertion into the four smallest (key, index) pairs; strict '<' keeps the earlier entry on ties
__device__ __forceinline__ void top4_insert(int key, int id, int& k0, int& k1, int& k2, int& k3, int& i0, int& i1,
                                            int& i2, int& i3) {
  if (key < k1) {
    if (key < k0) { k3 = k2; i3 = i2; k2 = k1; i2 = i1; k1 = k0; i1 = i0; k0 = key; i0 = id; }
    else { k3 = k2; i3 = i2; k2 = k1; i2 = i1; k1 = key; i1 = id; }
  } else {
    if (key < k2) { k3 = k2; i3 = i2; k2 = key; i2 = id; }
    else { k3 = key; i3 = id; }
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(n2::THREADS, 1)
nearest2_tc_kernel(const __grid_constant__ CUtensorMap tm_cstack, const float* __restrict__ mu,
                   const float* __restrict__ cn_inf /* ||c||^2, +huge on padding rows */,
                   const float* __restrict__ cnat /* [Kpad,16] */, int64_t n, int num_blocks, int n_centroids,
                   int64_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  constexpr int STAGES = n2::STAGES;
  constexpr uint32_t OFF_C = n2::OFF_C, OFF_CN = n2::OFF_CN, CN_BYTES = n2::CN_BYTES, TM_S = n2::TM_S,
                     TM_ZHI = n2::TM_ZHI, TM_ZLO = n2::TM_ZLO;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + n2::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_CN_FULL = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto BAR_S_FULL = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto BAR_FREE = [&](int s) { return bar0 + 8u * (3 * STAGES + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + n2::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_CN_FULL(s), 1); mbar_init(BAR_S_FULL(s), 1);
      mbar_init(BAR_FREE(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cstack) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + n2::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + n2::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int quarter = warp & 3;
  const int prow = quarter * 32 + lane;
  const int grp = (warp >= 6) ? 1 : 0;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
  float zrow[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) zrow[j] = 0.f;
  if (warp >= 2) {
    const int64_t r = row0 + prow;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(mu + r * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(src + q);
        zrow[4 * q] = v.x; zrow[4 * q + 1] = v.y; zrow[4 * q + 2] = v.z; zrow[4 * q + 3] = v.w;
      }
    }
    if (grp == 0) {
      uint32_t zh[16], zl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float hh = tf32_rna(zrow[j]);
        zh[j] = __float_as_uint(hh);
        zl[j] = __float_as_uint(zrow[j] - hh);
      }
      TMEM_ST16(tmem_base + lane_addr + TM_ZHI, zh);
      TMEM_ST16(tmem_base + lane_addr + TM_ZLO, zl);
      tmem_wait_st();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (warp == 0) {
    // =========================================================== TMA producer: centroid tile + ||c||^2 slice
    for (int j = 0; j < num_blocks; ++j) {
      const int st = j % STAGES;
      mbar_wait(BAR_FREE(st), ((j / STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_C_FULL(st), C_TILE_BYTES);
        const uint32_t dst = base + OFF_C + st * C_TILE_BYTES;
        if (PAIR) {
          tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(st), 0, j * BK + 32 * (int)rank);
        } else {
          tma_load_2d(dst, &tm_cstack, BAR_C_FULL(st), 0, j * BK);
          tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(st), 0, j * BK + 32);
        }
        mbar_expect_tx(BAR_CN_FULL(st), CN_BYTES);
        bulk_load_1d(base + OFF_CN + st * CN_BYTES, cn_inf + (int64_t)j * BK, CN_BYTES, BAR_CN_FULL(st));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer: the distance GEMM only
    if (leader) {
      const uint64_t c_desc0 = make_desc_sw128(base + OFF_C);
      constexpr uint32_t ID1 = make_idesc(PAIR ? 256 : 128, BK);
      for (int j = 0; j < num_blocks; ++j) {
        const int st = j % STAGES;
        mbar_wait(BAR_C_FULL(st), (j / STAGES) & 1);     // implies the scan groups of both CTAs freed S(st)
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + TM_S + st * 64;
          const uint64_t bc = c_desc0 + ((st * C_TILE_BYTES) >> 4);
          MMA_TS(d, tmem_base + TM_ZHI, bc, ID1, 0);
          MMA_TS(d, tmem_base + TM_ZHI + 8, bc + 2, ID1, 1);
          MMA_TS(d, tmem_base + TM_ZHI, bc + 4, ID1, 1);
          MMA_TS(d, tmem_base + TM_ZHI + 8, bc + 6, ID1, 1);
          MMA_TS(d, tmem_base + TM_ZLO, bc, ID1, 1);
          MMA_TS(d, tmem_base + TM_ZLO + 8, bc + 2, ID1, 1);
          COMMIT(BAR_S_FULL(st));
        }
        __syncwarp();
      }
    }
  } else {
    // =========================================================== scan groups: four smallest keys per point.
    // Per super-block and thread: 64 keys -> ordered ints with the in-block index in the low 6 bits ->
    // the three smallest of the even and of the odd elements by a branch-free min/max insertion network
    // (two independent chains for ILP; ~9 instructions per element).  Only those 6 block candidates go
    // through the branchy top-4 insertion (a per-element insertion is if-converted by ptxas into ~35
    // instructions for EVERY element, or -- as a real branch -- pays ~150 cycles of divergence 12% of
    // the time: both measured at 10-15 ms per 2^20 x 10k, no better than the scalar kernel).
    constexpr int IMAX = 0x7fffffff;
    int k0 = IMAX, k1 = IMAX, k2 = IMAX, k3 = IMAX;
    int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
    for (int j = grp; j < num_blocks; j += 2) {
      const int st = j % STAGES;
      mbar_wait(BAR_CN_FULL(st), (j / STAGES) & 1);
      mbar_wait(BAR_S_FULL(st), (j / STAGES) & 1);
      tc_fence_after();
      int a0 = IMAX, a1 = IMAX, a2 = IMAX, b0 = IMAX, b1 = IMAX, b2 = IMAX;
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {
        uint32_t sv[32];
        TMEM_LD32(tmem_base + lane_addr + TM_S + st * 64 + rnd * 32, sv);
        const float4* cn4 = reinterpret_cast<const float4*>(gbase + OFF_CN + st * CN_BYTES) + rnd * 8;
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 cv = cn4[q];
          const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = rnd * 32 + 4 * q + e;
            const int v = (ordered_int(fmaf(-2.f, __uint_as_float(sv[4 * q + e]), c4[e])) & ~63) | i;
            if (e & 1) {
              const int x = max(b0, v); b0 = min(b0, v);
              const int y = max(b1, x); b1 = min(b1, x);
              b2 = min(b2, y);
            } else {
              const int x = max(a0, v); a0 = min(a0, v);
              const int y = max(a1, x); a1 = min(a1, x);
              a2 = min(a2, y);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_FREE(st));
      const int cand[6] = {a0, a1, a2, b0, b1, b2};
#pragma unroll
      for (int q = 0; q < 6; ++q)
        if (cand[q] < k3) top4_insert(cand[q], j * BK + (cand[q] & 63), k0, k1, k2, k3, i0, i1, i2, i3);
    }
    // ---- merge the two groups' candidates, then decide by exact differences
    int* merge = reinterpret_cast<int*>(gbase + n2::OFF_MERGE);
    if (grp == 1) {
      merge[prow * 8 + 0] = i0; merge[prow * 8 + 1] = i1; merge[prow * 8 + 2] = i2; merge[prow * 8 + 3] = i3;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0) {
      const int64_t r = row0 + prow;
      if (r < n) {
        int cand[8] = {i0, i1, i2, i3, merge[prow * 8 + 0], merge[prow * 8 + 1], merge[prow * 8 + 2],
                       merge[prow * 8 + 3]};
        float b0 = 3.4e38f, b1 = 3.4e38f;
        int j0 = 0x7fffffff, j1 = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int id = cand[q];
          if (id < 0 || id >= n_centroids) continue;      // empty slot or a padding row
          const float4* crow = reinterpret_cast<const float4*>(cnat + (int64_t)id * 16);
          float sq = 0.f;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const float4 cv = __ldg(crow + v4);
            float df = zrow[4 * v4] - cv.x; sq = fmaf(df, df, sq);
            df = zrow[4 * v4 + 1] - cv.y; sq = fmaf(df, df, sq);
            df = zrow[4 * v4 + 2] - cv.z; sq = fmaf(df, df, sq);
            df = zrow[4 * v4 + 3] - cv.w; sq = fmaf(df, df, sq);
          }
          // (sq, index) lexicographic: the index-ordered strict-less scan of nearest2_kernel
          if (sq < b0 || (sq == b0 && id < j0)) { b1 = b0; j1 = j0; b0 = sq; j0 = id; }
          else if (sq < b1 || (sq == b1 && id < j1)) { b1 = sq; j1 = id; }
        }
        idx_out[r * 2 + 0] = j0; idx_out[r * 2 + 1] = j1;
        dist_out[r * 2 + 0] = sqrtf(b0); dist_out[r * 2 + 1] = sqrtf(b1);
      }
    }
  }
#undef MMA_TS
#undef COMMIT

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled16)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

static int make_map_h(PFN_encodeTiled16 enc, CUtensorMap* map, void* ptr, uint64_t inner, uint64_t outer,
                      uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * sizeof(__half)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp16) failed with CUresult " + std::to_string((int)r));
    return 4;
  }
  return 0;
}


// natural fp16 table [Kpad, 192] viewed as [3 column atoms][Kpad][64]: one box = 3 atoms x 64 (pair: 32) rows
static int make_map_h_atoms(PFN_encodeTiled16 enc, CUtensorMap* map, void* ptr, uint64_t Kpad, uint32_t box_rows) {
  cuuint64_t dims[3] = {64, Kpad, 3};
  cuuint64_t strides[2] = {192 * sizeof(__half), 64 * sizeof(__half)};
  cuuint32_t box[3] = {64, box_rows, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp16, 3d) failed with CUresult " + std::to_string((int)r));
    return 4;
  }
  return 0;
}

int tc_build_h16_descriptors(rlvae_tables* t) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  RLVAE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  RLVAE_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  PFN_encodeTiled16 enc = reinterpret_cast<PFN_encodeTiled16>(fn);
  const uint64_t Kpad = (uint64_t)t->Kpad;
  // packed-transposed fp16 tables [144, Kpad]: one box = 64 centroids (128 B) x 144 (pair: 72) rows
  if (int rc = make_map_h(enc, &t->tm_mh_hi, t->Mh_hi, Kpad, tc::h16::NCOLS, tc::BK, tc::h16::NCOLS)) return rc;
  if (int rc = make_map_h(enc, &t->tm_mh_lo, t->Mh_lo, Kpad, tc::h16::NCOLS, tc::BK, tc::h16::NCOLS)) return rc;
  if (int rc = make_map_h(enc, &t->tm_mh2_hi, t->Mh_hi, Kpad, tc::h16::NCOLS, tc::BK, tc::h16::NCOLS / 2)) return rc;
  if (int rc = make_map_h(enc, &t->tm_mh2_lo, t->Mh_lo, Kpad, tc::h16::NCOLS, tc::BK, tc::h16::NCOLS / 2)) return rc;
  if (int rc = make_map_h(enc, &t->tm_c16h, t->c16h, 64, Kpad, 64, 32)) return rc;   // 32 centroid rows (128 B each) per box
  if (int rc = make_map_h_atoms(enc, &t->tm_mnh_hi, t->Mnh_hi, Kpad, tc::BK)) return rc;
  if (int rc = make_map_h_atoms(enc, &t->tm_mnh_lo, t->Mnh_lo, Kpad, tc::BK)) return rc;
  if (int rc = make_map_h_atoms(enc, &t->tm_mnh2_hi, t->Mnh_hi, Kpad, tc::BK / 2)) return rc;
  if (int rc = make_map_h_atoms(enc, &t->tm_mnh2_lo, t->Mnh_lo, Kpad, tc::BK / 2)) return rc;
  return 0;
}

static bool h16_use_pairs() {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_PAIR");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();   // initialised once, thread-safe (C++11 magic static)
  return v == 1;
}

// The gradient kernel has its own switch (RLVAE_TC_PAIR_GRAD, default: follow RLVAE_TC_PAIR): with its small
// N = 64 / N = 16 MMAs the CTA-pair form saves little B-tile traffic and pays for cross-CTA barrier hops.
static bool g16_use_pairs() {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_PAIR_GRAD");
    return (e == nullptr) ? -1 : ((e[0] == '0') ? 0 : 1);
  }();
  return v < 0 ? h16_use_pairs() : v == 1;
}

template <bool PAIR, bool EXACT, bool HYBRID = false, bool HMC = false>
static int launch_h16(const rlvae_tables* t, const float* z, int64_t n, const tc::FusedOut& fo, cudaStream_t s,
                      const tc::HmcArgs& hm = tc::HmcArgs{}) {
  auto kern = tc::inverse_metric_h16_kernel<PAIR, EXACT, HYBRID, HMC>;
  RLVAE_OPT_IN_SMEM(kern, (int)(HYBRID ? tc::h16::SMEM_BYTES_HYB : tc::h16::SMEM_BYTES));
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 1, 1);
  cfg.blockDim = dim3(tc::h16::THREADS, 1, 1);
  cfg.dynamicSmemBytes = HYBRID ? tc::h16::SMEM_BYTES_HYB : tc::h16::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float alpha = 1.4426950408889634f / t->T2;
  const float* cbias = EXACT ? t->cmask : t->cbias_h;
  const float* cnat = t->c;
  const int nb = t->Kpad / tc::BK;
  const float lambda = t->lambda;
  const float out_scale = t->h16_out_scale;
  const float cu = t->c16_unscale;
  if (PAIR) {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c16h, t->tm_mh2_hi, t->tm_mh2_lo, z, cbias, cnat, n, nb,
                                       alpha, lambda, out_scale, cu, t->cshift, 14.f - t->hybrid_bits, fo, hm));
  } else {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c16h, t->tm_mh_hi, t->tm_mh_lo, z, cbias, cnat, n, nb,
                                       alpha, lambda, out_scale, cu, t->cshift, 14.f - t->hybrid_bits, fo, hm));
  }
  return 0;
}

// How the weights are formed (d == 16, symmetric tables):
//   0  expanded form on the tensor core              -- when its accuracy gate passes (expanded_ok)
//   2  hybrid: expanded form + exact refinement of   -- small T with lambda > 0: only weights above
//      the weights that matter                          2^-hybrid_bits can move G^{-1} by more than 1e-6 lambda
//   1  exact differences on the FMA pipe             -- everything else, or RLVAE_TC_EXACT=1
// RLVAE_TC_EXACT=2 forces the hybrid mode where its threshold exists.
int h16_mode(const rlvae_tables* t) {
  const char* e = getenv("RLVAE_TC_EXACT");       // read per call: the tests switch modes within one process
  const int forced = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
  if (forced == 1) return 1;
  if (forced == 2 && t->hybrid_ok) return 2;
  if (t->expanded_ok) return 0;
  return t->hybrid_ok ? 2 : 1;
}

// Symmetric tables, d == 16: any of { packed G^{-1}, packed G, lad_scale * log|det G^{-1}|, sign,
// diag(G) } from ONE kernel (+ the pivoting fallback pass for points that are not positive definite,
// which needs packed G^{-1}: a_packed must be given whenever a factor output is requested).
int launch_inverse_metric_h16(const rlvae_tables* t, const float* z, int64_t n, float* a_packed,
                              float* g_packed, float* logabsdet, float lad_scale, float* sign, float* diag_g,
                              int* fail_ws, cudaStream_t s, float* a_full, float* g_full, int a_packed_wanted,
                              float* s_diag) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->tensor_capable && t->symmetric && t->Mh_hi != nullptr,
                "split-fp16 tensor path needs latent_dim == 16 and symmetric tables");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0, "tensor path needs 16-byte aligned z");
  const bool fused = g_packed || g_full || logabsdet || sign || diag_g;
  RLVAE_REQUIRE(!fused || (fail_ws != nullptr && a_packed != nullptr),
                "fused factor outputs need the packed G^{-1} buffer and the fallback workspace");
  RLVAE_REQUIRE(n < (int64_t)1 << 31, "batch too large for the 32-bit fallback list");
  if (fused) RLVAE_CUDA_OK(cudaMemsetAsync(fail_ws, 0, sizeof(int), s));
  RLVAE_REQUIRE(a_full == nullptr || (reinterpret_cast<uintptr_t>(a_full) & 15) == 0, "G^-1 output must be 16-byte aligned");
  RLVAE_REQUIRE(g_full == nullptr || (reinterpret_cast<uintptr_t>(g_full) & 15) == 0, "G output must be 16-byte aligned");
  // certified tables: nobody reads the packed G^{-1} unless the caller asked for it (the fallback list is
  // recomputed from the tables in the rare rounding-level failure), so it is not stored
  const bool skip_packed = !a_packed_wanted && t->psd_certified;
  RLVAE_REQUIRE(s_diag == nullptr || (reinterpret_cast<uintptr_t>(s_diag) & 15) == 0, "s_diag must be 16-byte aligned");
  tc::FusedOut fo{a_full, skip_packed ? nullptr : a_packed, g_packed, g_full, logabsdet, sign, diag_g, fail_ws, lad_scale,
                  s_diag};
  int rc;
  const int mode = h16_mode(t);
  prof_mark(0, s);
  if (mode == 1) rc = h16_use_pairs() ? launch_h16<true, true>(t, z, n, fo, s) : launch_h16<false, true>(t, z, n, fo, s);
  else if (mode == 2) rc = h16_use_pairs() ? launch_h16<true, false, true>(t, z, n, fo, s) : launch_h16<false, false, true>(t, z, n, fo, s);
  else rc = h16_use_pairs() ? launch_h16<true, false>(t, z, n, fo, s) : launch_h16<false, false>(t, z, n, fo, s);
  prof_mark(1, s);
  if (rc) return rc;
  RLVAE_REQUIRE(g_full == nullptr || g_packed != nullptr, "the expanded G output needs the packed G buffer too");
  if (fused)
    return launch_sym16_fallback(a_packed, n, g_packed, logabsdet, lad_scale, sign, diag_g, fail_ws, s, g_full,
                                 skip_packed ? t : nullptr, z);
  return 0;
}

// The whole HMC trajectory in one launch (tables certified positive semi-definite, variant-A drift):
// n_iters MCMC iterations x (n_lf + 1) metric evaluations per CTA pair, chain state on chip.
int h16_hmc_available(const rlvae_tables* t) {
  return t->d == 16 && t->tensor_capable && t->symmetric && t->Mh_hi != nullptr && t->psd_certified &&
         h16_use_pairs() && (t->Kpad / tc::BK) % 2 == 0;
}

int launch_hmc_trajectory_h16(const rlvae_tables* t, float* z, const float* gamma, const float* acc, int64_t n,
                              int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* scales_dev,
                              const float* h_scales, float* h0, float* h1, float* alpha, float* moves,
                              float* z_trace, int* fail_count, cudaStream_t s) {
  if (n == 0 || n_iters == 0) return 0;
  RLVAE_REQUIRE(h16_hmc_available(t), "fused HMC trajectory needs latent_dim == 16, symmetric certified tables");
  RLVAE_REQUIRE(n_lf >= 1 && n_iters >= 1, "fused HMC trajectory: need n_lf >= 1 and n_iters >= 1");
  RLVAE_REQUIRE(scales_dev != nullptr || (n_iters == 1 && n_lf <= 64 && h_scales != nullptr),
                "fused HMC trajectory: tempering scales on the device are required for n_iters > 1 or n_lf > 64");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0,
                "fused HMC trajectory needs 16-byte aligned z and gamma");
  RLVAE_REQUIRE(z_trace == nullptr || (reinterpret_cast<uintptr_t>(z_trace) & 15) == 0, "z_trace must be 16-byte aligned");
  RLVAE_REQUIRE(((int64_t)n_iters * n_lf + 1) * (t->Kpad / tc::BK) < ((int64_t)1 << 30), "trajectory too long for one launch");
  tc::HmcArgs hm{};
  hm.z = z; hm.gamma = gamma; hm.acc = acc; hm.scales = scales_dev;
  hm.h0 = h0; hm.h1 = h1; hm.alpha = alpha; hm.moves = moves; hm.z_trace = z_trace; hm.fail_count = fail_count;
  hm.n_iters = n_iters; hm.n_lf = n_lf; hm.eps = eps_lf; hm.b0 = beta_zero_sqrt; hm.T2 = t->T2;
  if (scales_dev == nullptr)
    for (int i = 0; i < n_lf; ++i) hm.scales_inl[i] = h_scales[i];
  tc::FusedOut fo{};
  const int mode = h16_mode(t);
  if (mode == 1) return launch_h16<true, true, false, true>(t, z, n, fo, s, hm);
  if (mode == 2) return launch_h16<true, false, true, true>(t, z, n, fo, s, hm);
  return launch_h16<true, false, false, true>(t, z, n, fo, s, hm);
}

template <bool PAIR, bool EXACT, bool HYBRID = false>
static int launch_g16(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale, float* out,
                      cudaStream_t s, int u_packed) {
  auto kern = tc::metric_grad_h16_kernel<PAIR, EXACT, HYBRID>;
  RLVAE_OPT_IN_SMEM(kern, (int)(HYBRID ? tc::g16::SMEM_BYTES_HYB : tc::g16::SMEM_BYTES));
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 1, 1);
  cfg.blockDim = dim3(tc::g16::THREADS, 1, 1);
  cfg.dynamicSmemBytes = HYBRID ? tc::g16::SMEM_BYTES_HYB : tc::g16::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float alpha = 1.4426950408889634f / t->T2;
  const float* cbias = EXACT ? t->cmask : t->cbias_h;
  const float* cnat = t->c;
  const int nb = t->Kpad / tc::BK;
  const bool unit = u_packed == 2;                     // t_k = 1 exactly: nothing to un-scale
  const float sc = unit ? scale : scale * t->h16_m_unscale;          // 2^-eM of the table scaling
  const float cu = t->c16_unscale;
  if (PAIR) {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c16h, t->tm_mnh2_hi, t->tm_mnh2_lo,
                                     unit ? t->tm_bt8_hi : t->tm_ct8_hi, unit ? t->tm_bt8_lo : t->tm_ct8_lo, z, u, cbias,
                                     cnat, n, nb, alpha, sc, cu, t->cshift, -t->hybrid_bits, out, u_packed));
  } else {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c16h, t->tm_mnh_hi, t->tm_mnh_lo,
                                     unit ? t->tm_bt16_hi : t->tm_ct16_hi, unit ? t->tm_bt16_lo : t->tm_ct16_lo, z, u,
                                     cbias, cnat, n, nb, alpha, sc, cu, t->cshift, -t->hybrid_bits, out, u_packed));
  }
  return 0;
}

// (scale) * sum_k w_k <U, M_k> (c_k - z) for symmetric tables; u is [N,256] (any U) or, with
// u_packed, a symmetric U in the packed [N,144] layout.
int launch_metric_grad_h16(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale,
                           float* out, cudaStream_t s, int u_packed) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->tensor_capable && t->symmetric && t->Mnh_hi != nullptr,
                "split-fp16 gradient path needs latent_dim == 16 and symmetric tables");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0, "tensor path needs 16-byte aligned z, u and out");
  RLVAE_REQUIRE(u_packed != 2 || t->bt_hi != nullptr, "unit-weight mode needs the pythae table");
  const int mode = h16_mode(t);
  if (mode == 1)
    return g16_use_pairs() ? launch_g16<true, true>(t, z, u, n, scale, out, s, u_packed)
                           : launch_g16<false, true>(t, z, u, n, scale, out, s, u_packed);
  if (mode == 2)
    return g16_use_pairs() ? launch_g16<true, false, true>(t, z, u, n, scale, out, s, u_packed)
                           : launch_g16<false, false, true>(t, z, u, n, scale, out, s, u_packed);
  return g16_use_pairs() ? launch_g16<true, false>(t, z, u, n, scale, out, s, u_packed)
                         : launch_g16<false, false>(t, z, u, n, scale, out, s, u_packed);
}

template <bool PAIR>
static int launch_n2(const rlvae_tables* t, const float* mu, int64_t n, int64_t* idx, float* dist, cudaStream_t s) {
  auto kern = tc::nearest2_tc_kernel<PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::n2::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 1, 1);
  cfg.blockDim = dim3(tc::n2::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::n2::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float* cn = t->cn_inf;
  const float* cnat = t->c;
  const int nb = t->Kpad / tc::BK;
  const int K = t->K;
  RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_cstack, mu, cn, cnat, n, nb, K, idx, dist));
  return 0;
}

// two nearest centroids, d == 16: tensor-core pre-filter + exact decision (bit-identical to the direct kernel)
int launch_nearest2_tc(const rlvae_tables* t, const float* mu, int64_t n, int64_t* idx, float* dist, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->cstack != nullptr && t->cn_inf != nullptr, "tensor nearest2 needs latent_dim == 16");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(mu) & 15) == 0, "tensor nearest2 needs 16-byte aligned mu");
  return h16_use_pairs() ? launch_n2<true>(t, mu, n, idx, dist, s) : launch_n2<false>(t, mu, n, idx, dist, s);
}

}  // namespace rlvae
