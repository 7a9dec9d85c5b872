// tcgen05 / TMEM / TMA implementation of the metric hot path for latent_dim == 16 (sm_100a).
//
//   G^{-1}[n] = sum_k exp(-||z_n - c_k||^2 / T^2) M_k + lambda I
//   ref: src/models/components/metric_tensor.py:115-135 (the [N,K,d,d] product + reduction)
//
// It is un-normalised flash attention with asymmetric head dims (QK dim 16, V dim 256):
//
//   GEMM1  S[128 x 32]   = Z[128 x 16] . C^T           3xTF32: (z_hi|z_hi).(c_hi|c_lo) + z_lo.c_hi
//   exp    P = 2^(S*2a + bias_k + zb_n),  a = log2(e)/T^2, bias_k = -a||c_k||^2, zb_n = -a||z_n||^2
//          P is split P_hi (top 19 bits = exactly what the TF32 datapath reads) and P_lo = P - P_hi
//   GEMM2  O[128 x 128] += P_hi.M_hi + P_lo.M_hi + P_hi.M_lo      (3xTF32, fp32 accumulate in TMEM)
//
// One CTA = 128 latent points (TMEM lanes) x 128 of the 256 output columns (blockIdx.y picks the
// half; both halves recompute the cheap S/exp stage).  Warp roles:
//   warp 0     TMA producer   : centroid tiles (4 KB) + bias (128 B) ring, M_hi/M_lo tile (16 KB) ring
//   warp 1     MMA issuer     : one elected thread issues every tcgen05.mma; owns TMEM alloc/dealloc
//   warps 2-5  exp group A    : even blocks; one thread per point: tcgen05.ld S -> exp2 -> split ->
//   warps 6-9  exp group B    : odd blocks;   tcgen05.st P (A operand of GEMM2 is read from TMEM),
//                               chunk folding (64 columns each) and the epilogue
//
// Accumulation accuracy.  The tensor core adds into its fp32 accumulator with truncation, so a
// K=10k reduction (3750 accumulating MMAs per output) drifts by ~1e-4 relative -- measured 5e-5 at
// K=3000 -- which breaks the 1e-5 contract.  The MMA therefore accumulates only CHUNK_BLOCKS*32
// centroids at a time into one of two "chunk" accumulators; the exp warps fold every finished
// chunk into the running total with round-to-nearest fp32 adds (Ootomo & Yokota's remedy).
// The running total lives in the exp threads' registers (64 columns each).
// The exp stage is the latency-critical path (S -> P must finish inside the MMA time of the blocks
// in flight), so two exp warpgroups alternate blocks and GEMM1 runs two blocks ahead of GEMM2.
// A tcgen05.mma never takes less than ~45 cycles on this part (measured, scripts/micro/mma_rate.cu),
// so GEMM1 is issued per 64-centroid super-block (N = 64), not per 32.
// TMEM columns: [0,128) / [128,256) chunk accumulators, [256,512) two S/P buffers of
//               (64 S|P_hi + 64 P_lo).
// [N,K] never exists outside TMEM; the tables stream L2 -> smem once per 128 points.
#include "rlvae_tc_common.cuh"

// -DRLVAE_TC_PROFILE: CTA (0,0) prints where its MMA thread and exp group A spend their cycles
#ifdef RLVAE_TC_PROFILE
__device__ long long g_trace[4][16][8];
#define TRACE(role, j, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (j) >= 40 && (j) < 56 && lane == 0) g_trace[role][(j) - 40][slot] = clock64(); } while (0)
#else
#define TRACE(role, j, slot) do {} while (0)
#endif
#ifdef RLVAE_TC_PROFILE
#define PROF_T0() long long _pt = clock64()
#define PROF_ADD(acc) do { long long _n = clock64(); (acc) += _n - _pt; _pt = _n; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(acc) do {} while (0)
#endif

namespace rlvae {
namespace tc {

// ------------------------------------------------------------------------------------------ forward kernel
// Warp-specialised, 512 threads = 4 warpgroups:
//   WG0  warp 0 TMA producer, warp 1 MMA issuer (warps 2,3 idle)          -> 40 registers
//   WG1  exp group A (even super-blocks), WG2 exp group B (odd)            -> 112 registers
//   WG3  fold group: owns the fp32 running total, folds every finished chunk, writes the output -> 232
// Pipeline per 64-centroid super-block j (all MMAs execute in issue order):
//   GEMM1(j+3) is issued right after GEMM2(j), so S(j+3) is ready ~2 iterations before P(j+3) is
//   needed: the exp stage (MUFU-bound, ~1.2-1.5k cycles of latency) is off the critical path.
//   Three S/P buffers + ONE chunk accumulator = 512 TMEM columns; the accumulator is handed to the
//   fold group at every chunk end (CH_FULL) and taken back (CH_FREE) while GEMM1(j+3) keeps the
//   tensor pipe busy.
namespace fwd {
constexpr int THREADS = 512;
constexpr int C_STAGES = 4;
constexpr int SP_BUFS = 3;
constexpr int M_STAGES = 5;
constexpr int AHEAD = 3;             // GEMM1 runs this many super-blocks ahead of GEMM2
constexpr uint32_t OFF_C = OFF_A2 + A_BYTES;
constexpr uint32_t OFF_M = OFF_C + C_STAGES * C_TILE_BYTES;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 2;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TILE_M * OUT_LD * 4 <= M_STAGES * M_TILE_BYTES, "epilogue staging must fit the M ring");
constexpr uint32_t TM_ACC = 0;       // chunk accumulator (<= 128 columns)
constexpr uint32_t TM_SP = 128;      // + buf*128 : S/P_hi (64) ; + 64 : P_lo (64)
}  // namespace fwd


template <bool SYM, bool PAIR>
__global__ void __launch_bounds__(fwd::THREADS, 1)
inverse_metric_tc_kernel(const __grid_constant__ CUtensorMap tm_cstack,
                         const __grid_constant__ CUtensorMap tm_mt_hi,
                         const __grid_constant__ CUtensorMap tm_mt_lo,
                         const float* __restrict__ z, const float* __restrict__ cbias, int64_t n,
                         int num_blocks, int chunk_blocks, float alpha /* log2(e)/T^2 */, float lambda,
                         float* __restrict__ out) {
  constexpr int C_STAGES = fwd::C_STAGES, SP_BUFS = fwd::SP_BUFS, M_STAGES = fwd::M_STAGES, AHEAD = fwd::AHEAD;
  constexpr uint32_t OFF_C = fwd::OFF_C, OFF_M = fwd::OFF_M, OFF_BIAS = fwd::OFF_BIAS, TM_ACC = fwd::TM_ACC,
                     TM_SP = fwd::TM_SP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

  const uint32_t bar0 = base + fwd::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_BIAS_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };
  auto BAR_M_FULL = [&](int s) { return bar0 + 8u * (3 * C_STAGES + s); };
  auto BAR_M_EMPTY = [&](int s) { return bar0 + 8u * (3 * C_STAGES + M_STAGES + s); };
  auto BAR_S_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + b); };
  auto BAR_P_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + SP_BUFS + b); };
  const uint32_t BAR_CH_FULL = bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS);
  const uint32_t BAR_CH_FREE = BAR_CH_FULL + 8u;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + fwd::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const int half = blockIdx.y;         // which 128 of the 256 output columns (SYM: 80 + 64 of 144)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader of the CTA pair
  const bool leader = rank == 0;
  const int ncols = SYM ? (half == 0 ? SYM_H0 : SYM_COLS - SYM_H0) : NHALF;   // MMA N of GEMM2
  const int col_base = SYM ? half * SYM_H0 : half * NHALF;
  // rows of the B tile this CTA holds per 32-centroid atom (a pair splits N between its CTAs)
  constexpr int ROWS_BOX = (SYM ? SYM_H0 : NHALF) / (PAIR ? 2 : 1);
  constexpr uint32_t ATOM_BYTES = ROWS_BOX * 128;
  constexpr uint32_t TILE_BYTES = 2 * ATOM_BYTES;        // bytes TMA delivers per stage per CTA
  constexpr uint32_t ATOM_DESC = ATOM_BYTES >> 4;
  const int row_cta = col_base + (PAIR ? (int)rank * (ncols / 2) : 0);
  const uint32_t idesc_g2 = make_idesc(PAIR ? 256 : 128, ncols);
  constexpr int NPAIR = PAIR ? 2 : 1;
  const int num_chunks = (num_blocks + chunk_blocks - 1) / chunk_blocks;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), 4); mbar_init(BAR_BIAS_FULL(s), 1);
    }
    for (int s = 0; s < SP_BUFS; ++s) { mbar_init(BAR_S_FULL(s), 1); mbar_init(BAR_P_FULL(s), 4 * NPAIR); }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    mbar_init(BAR_CH_FULL, 1);
    mbar_init(BAR_CH_FREE, 4 * NPAIR);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cstack) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mt_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mt_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + fwd::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + fwd::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }

  // row identity of the exp / fold threads: TMEM lane quarter = warp % 4, point = quarter*32 + lane
  const int quarter = warp & 3;
  const int prow = quarter * 32 + lane;
  float zb = 0.f;
  if (wg == 1) {          // exp group A writes the GEMM1 operand tiles
    zb = write_z_tiles(gbase, z, row0 + prow, n, prow, alpha);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (UMMA)
  } else if (wg == 2) {
    const int64_t r = row0 + prow;
    float nrm = 0.f;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = __ldg(src + q);
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
      }
    }
    zb = -nrm * alpha;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // barriers initialised in BOTH CTAs before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (wg == 0) {
    reg_dec<40>();
    if (warp == 0) {
      // =========================================================== TMA producer (warp-converged)
      // In a pair every CTA fetches only ITS half of each B tile; all transaction bytes complete on
      // the leader's FULL barrier.
      auto load_m = [&](int jm) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int it = 2 * jm + h, ms = it % M_STAGES;
          mbar_wait(BAR_M_EMPTY(ms), ((it / M_STAGES) & 1) ^ 1);
          if (elect_one()) {
            if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * TILE_BYTES);
            const CUtensorMap* map = (h == 0) ? &tm_mt_hi : &tm_mt_lo;
            const uint32_t dst = base + OFF_M + ms * M_TILE_BYTES;
            if (PAIR) {
              tma_load_2d_pair(dst, map, BAR_M_FULL(ms), jm * BK, row_cta);
              tma_load_2d_pair(dst + ATOM_BYTES, map, BAR_M_FULL(ms), jm * BK + 32, row_cta);
            } else {
              tma_load_2d(dst, map, BAR_M_FULL(ms), jm * BK, row_cta);
              tma_load_2d(dst + ATOM_BYTES, map, BAR_M_FULL(ms), jm * BK + 32, row_cta);
            }
          }
          __syncwarp();
        }
      };
      for (int j = 0; j < num_blocks; ++j) {
        const int cs = j % C_STAGES;
        mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE_BYTES);
          const uint32_t dst = base + OFF_C + cs * C_TILE_BYTES;
          if (PAIR) {           // 32 of the 64 centroid rows per CTA (box = 32 rows)
            tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32 * (int)rank);
          } else {
            tma_load_2d(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK);
            tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32);
          }
          mbar_expect_tx(BAR_BIAS_FULL(cs), BIAS_BYTES);
          bulk_load_1d(base + OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_BIAS_FULL(cs));
        }
        __syncwarp();
        // the M tiles trail the centroid tiles like GEMM2 trails GEMM1
        if (j >= AHEAD) load_m(j - AHEAD);
      }
      for (int jm = (num_blocks >= AHEAD ? num_blocks - AHEAD : 0); jm < num_blocks; ++jm) load_m(jm);
    } else if (warp == 1 && leader) {
      // =========================================================== MMA issuer (warp-converged; pair: leader only)
      const uint64_t a1_desc = make_desc_sw128(base + OFF_A1);
      const uint64_t a2_desc = make_desc_sw128(base + OFF_A2);
      auto gemm1 = [&](int j, bool waited) {
        const int cs = j % C_STAGES, sb = j % SP_BUFS;
        if (!waited) mbar_wait(BAR_C_FULL(cs), (j / C_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_gemm1<PAIR>(tmem_base + TM_SP + sb * 128, a1_desc, a2_desc,
                             make_desc_sw128(base + OFF_C + cs * C_TILE_BYTES));
          COMMIT(BAR_S_FULL(sb));   // the C stage itself is released by the exp warps (after S_FULL)
        }
        __syncwarp();
      };
      auto wait_m = [&](int it) { mbar_wait(BAR_M_FULL(it % M_STAGES), (it / M_STAGES) & 1); };
      for (int j = 0; j < AHEAD && j < num_blocks; ++j) gemm1(j, false);
      wait_m(0);
      wait_m(1);
      mbar_wait(BAR_P_FULL(0), 0);
      // Every wait below is for something that is normally long complete; each sits behind a
      // group of 8 queued MMAs so that its latency never drains the tensor pipe.
      for (int j = 0; j < num_blocks; ++j) {
        const int chunk = j / chunk_blocks;
        const int first = (j % chunk_blocks) == 0;     // a new chunk overwrites the accumulator ...
        if (first && chunk >= 1) mbar_wait(BAR_CH_FREE, (chunk - 1) & 1);   // ... once the fold group has drained it
        tc_fence_after();
        const int sb = j % SP_BUFS;
        const int ms_hi = (2 * j) % M_STAGES, ms_lo = (2 * j + 1) % M_STAGES;
        const uint32_t p_hi = tmem_base + TM_SP + sb * 128;
        const uint32_t p_lo = p_hi + 64;
        const uint32_t acc = tmem_base + TM_ACC;
        const uint64_t bh = make_desc_sw128(base + OFF_M + ms_hi * M_TILE_BYTES);
        const uint64_t bl = make_desc_sw128(base + OFF_M + ms_lo * M_TILE_BYTES);
        // K index kk = 8 centroids; atom = kk / 4 (ATOM_BYTES apart)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, p_hi + 8 * kk, bh + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), idesc_g2, !(first && kk == 0));
        }
        __syncwarp();
        if (j + 1 < num_blocks) wait_m(2 * j + 2);
        if (j + AHEAD < num_blocks) mbar_wait(BAR_C_FULL((j + AHEAD) % C_STAGES), ((j + AHEAD) / C_STAGES) & 1);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, p_lo + 8 * kk, bh + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), idesc_g2, 1);
          COMMIT(BAR_M_EMPTY(ms_hi));
        }
        __syncwarp();
        if (j + 1 < num_blocks) wait_m(2 * j + 3);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            MMA_TS(acc, p_hi + 8 * kk, bl + (kk >> 2) * ATOM_DESC + 2 * (kk & 3), idesc_g2, 1);
          COMMIT(BAR_M_EMPTY(ms_lo));
          if ((j % chunk_blocks) == chunk_blocks - 1 || j == num_blocks - 1)
            COMMIT(BAR_CH_FULL);                      // chunk complete: hand the accumulator to the fold group
        }
        __syncwarp();
        // GEMM1 AHEAD super-blocks ahead re-uses this S/P buffer: ordered behind GEMM2(j) in the pipe,
        // and keeps the pipe busy while the fold group drains the accumulator at a chunk boundary
        if (j + AHEAD < num_blocks) gemm1(j + AHEAD, true);
        if (j + 1 < num_blocks) mbar_wait(BAR_P_FULL((j + 1) % SP_BUFS), ((j + 1) / SP_BUFS) & 1);
      }
    }
  } else if (wg == 1 || wg == 2) {
    // =========================================================== exp groups (one thread per point)
    reg_dec<112>();
    const int grp = wg - 1;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float two_alpha = 2.f * alpha;
    for (int j = grp; j < num_blocks; j += 2) {
      const int cs = j % C_STAGES, sb = j % SP_BUFS;
      const uint32_t sp = tmem_base + lane_addr + TM_SP + sb * 128;
      mbar_wait(BAR_BIAS_FULL(cs), (j / C_STAGES) & 1);   // bias bytes visible to this thread
      mbar_wait(BAR_S_FULL(sb), (j / SP_BUFS) & 1);
      tc_fence_after();
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {
        uint32_t s[32], l[32];
        TMEM_LD32(sp + rnd * 32, s);
        const float4* bias4 = reinterpret_cast<const float4*>(gbase + OFF_BIAS + cs * BIAS_BYTES) + rnd * 8;
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = bias4[q];
          const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * q + e;
            const float w = ex2_approx(fmaf(__uint_as_float(s[i]), two_alpha, b4[e] + zb));
            const uint32_t wh = __float_as_uint(w) & 0xFFFFE000u;   // what the TF32 datapath will read
            s[i] = wh;
            l[i] = __float_as_uint(w - __uint_as_float(wh));
          }
        }
        TMEM_ST32(sp + rnd * 32, s);
        TMEM_ST32(sp + 64 + rnd * 32, l);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_P_FULL(sb)); else mbar_arrive(BAR_P_FULL(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
    }
  } else {
    // =========================================================== fold group: fp32 running total + output
    reg_inc<232>();
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    constexpr int NTOT = SYM ? SYM_H0 : NHALF;          // columns kept per row (80 packed / 128 dense)
    float total[NTOT];
#pragma unroll
    for (int i = 0; i < NTOT; ++i) total[i] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      mbar_wait(BAR_CH_FULL, c & 1);
      tc_fence_after();
      const uint32_t src = tmem_base + lane_addr + TM_ACC;
#pragma unroll
      for (int cb = 0; cb < NTOT / 32; ++cb) {
        uint32_t a[32];
        TMEM_LD32(src + cb * 32, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) total[cb * 32 + i] += __uint_as_float(a[i]);
      }
      if (SYM) {      // columns 64..79 (half 1 only fills 64: the rest is never-written TMEM, never stored)
        uint32_t a[16];
        TMEM_LD16(src + 64, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) total[64 + i] += __uint_as_float(a[i]);
      }
      if (c + 1 < num_chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE); else mbar_arrive(BAR_CH_FREE); }
      }
    }
    // ---------------------------------------------------------- epilogue (all TMA / MMA work is complete)
    float* stage = reinterpret_cast<float*>(gbase + OFF_M);
    const int t = threadIdx.x - 384;
    const int64_t rows_here = (n - row0 < TILE_M) ? (n - row0) : TILE_M;
    if (!SYM) {
#pragma unroll
      for (int q = 0; q < NHALF / 4; ++q) {
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = half * NHALF + q * 4 + e;
          op[e] = total[q * 4 + e] + ((col % 17 == 0) ? lambda : 0.f);
        }
        *reinterpret_cast<float4*>(stage + prow * OUT_LD + q * 4) = o;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four fold warps only
      float* dst = out + row0 * NCOL + half * NHALF;
#pragma unroll 4
      for (int i = t; i < TILE_M * (NHALF / 4); i += 128) {
        const int r = i >> 5, c4 = i & 31;
        if (r < rows_here)
          *reinterpret_cast<float4*>(dst + (int64_t)r * NCOL + c4 * 4) =
              *reinterpret_cast<const float4*>(stage + r * OUT_LD + c4 * 4);
      }
    } else {
      // packed output [N, 144]: this CTA owns packed columns [col_base, col_base + ncols)
#pragma unroll
      for (int q = 0; q < SYM_H0 / 4; ++q) {
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pc = col_base + q * 4 + e;      // packed index
          bool diag = false;                        // p(i,i) = 16 i - i (i-1)/2
#pragma unroll
          for (int i = 0; i < 16; ++i) diag |= (pc == sym_index(i, i));
          op[e] = total[q * 4 + e] + (diag ? lambda : 0.f);
        }
        *reinterpret_cast<float4*>(stage + prow * SYM_OUT_LD + q * 4) = o;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* dst = out + row0 * SYM_COLS + col_base;
      const int c4n = ncols / 4;                                      // 20 or 16 float4 per row
      for (int i = t; i < TILE_M * c4n; i += 128) {
        const int r = i / c4n, c4 = i - r * c4n;
        if (r < rows_here)
          *reinterpret_cast<float4*>(dst + (int64_t)r * SYM_COLS + c4 * 4) =
              *reinterpret_cast<const float4*>(stage + r * SYM_OUT_LD + c4 * 4);
      }
    }
  }
#undef MMA_TS
#undef COMMIT

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // no CTA of a pair exits while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ==========================================================================================
// Gradient / backward kernel:  out[n,:] += scale * sum_k w_nk <U_n, M_k> (c_k - z_n)
//   (with U = G^T, scale = -2/T^2 it is grad_z log det G; with
//    U = dL/dG^{-1}, scale = 2/T^2 it is the autograd backward of metric_tensor.py:98-137)
//
// Same skeleton as the forward kernel.  Per 64-centroid super-block the tensor core produces
//   S[128 x 64] = Z.C^T                      (GEMM1, as in the forward kernel)
//   T[128 x 64] = U_hi.Mhi^T + U_lo.Mhi^T + U_hi.Mlo^T      (3xTF32, N = 64, K = 8 per MMA)
// with U (this CTA's share of the (i,j) columns, hi/lo split) resident in TMEM as the A operand
// and the natural-layout table tile [64 centroids x cols] as the K-major B operand.  The exp groups
// then form w = exp2(..), u = w*t and accumulate g += u*c_k, su += u in fp32 registers (packed
// fma.rn.f32x2); the two column halves (blockIdx.y) add their partial results into `out` (zeroed by
// the launcher; exactly two addends per element).  T is re-started every super-block, so the
// accumulator-truncation drift is bounded by <= 51 MMAs.
//   SYM : symmetric tables contract only the 136 packed columns: <U,M> = sum_{i<=j} Ut_p M_p with
//         Ut_p = U_ij + U_ji (i<j), U_ii -- valid for ANY U.  Half 0 takes packed columns [0,72),
//         half 1 [72,136) (9 / 8 K-steps instead of 16).
//   PAIR: CTA pair (cta_group::2): each CTA supplies 32 of the 64 centroid rows of every B tile.
// TMEM columns: [0,128) U_hi, [128,256) U_lo, [256,512) two (S 64 | T 64) buffers.
// ==========================================================================================
namespace grad {
constexpr int C_STAGES = 3;
constexpr int M_STAGES = 4;
constexpr uint32_t C_TILE_BYTES = BK * 128;        // [hi|lo] rows for GEMM1
constexpr uint32_t CN_BYTES = BK * 64;             // natural fp32 centroid rows for the FMA stage
constexpr uint32_t BIAS_BYTES = BK * 4;
constexpr uint32_t M_ATOM_BYTES = BK * 128;        // 64 centroid rows x 32 fp32 (pair: 32 rows used)
constexpr uint32_t M_TILE_BYTES = 4 * M_ATOM_BYTES;
constexpr uint32_t OFF_A1 = 0;
constexpr uint32_t OFF_A2 = OFF_A1 + A_BYTES;
constexpr uint32_t OFF_C = OFF_A2 + A_BYTES;
constexpr uint32_t OFF_M = OFF_C + C_STAGES * C_TILE_BYTES;
constexpr uint32_t OFF_CN = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BIAS = OFF_CN + C_STAGES * CN_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 3 * C_STAGES + 2 * M_STAGES + 4;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
constexpr int RED_LD = 20;                          // group-B partials staged in the M ring
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_UHI = 0, TM_ULO = 128, TM_ST = 256;   // ST + buf*128 : S (64) | T (64)
constexpr int SYM_SPLIT = 72;                       // packed columns of half 0 (9 K-steps); half 1: 64
constexpr int SYM_NAT_COLS = 160;                   // packed natural row length (5 atoms of 32)
}  // namespace grad

__device__ __forceinline__ void ffma2(float2& acc, float2 a, float2 b) {   // acc += a * b, two lanes
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
}

template <bool SYM, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
metric_grad_tc_kernel(const __grid_constant__ CUtensorMap tm_cstack,
                      const __grid_constant__ CUtensorMap tm_mn_hi,
                      const __grid_constant__ CUtensorMap tm_mn_lo,
                      const float* __restrict__ z, const float* __restrict__ u,
                      const float* __restrict__ cnat, const float* __restrict__ cbias, int64_t n,
                      int num_blocks, float alpha, float scale, float* __restrict__ out, int u_packed) {
  // local names shadow the forward kernel's constants of the same name
  constexpr int C_STAGES = grad::C_STAGES, M_STAGES = grad::M_STAGES, RED_LD = grad::RED_LD;
  constexpr uint32_t C_TILE_BYTES = grad::C_TILE_BYTES, CN_BYTES = grad::CN_BYTES,
                     BIAS_BYTES = grad::BIAS_BYTES, M_TILE_BYTES = grad::M_TILE_BYTES,
                     OFF_CN = grad::OFF_CN, TM_UHI = grad::TM_UHI, TM_ULO = grad::TM_ULO,
                     TM_ST = grad::TM_ST;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + grad::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_CB_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };   // natural c rows + bias (local)
  auto BAR_M_FULL = [&](int s) { return bar0 + 8u * (3 * C_STAGES + s); };
  auto BAR_M_EMPTY = [&](int s) { return bar0 + 8u * (3 * C_STAGES + M_STAGES + s); };
  auto BAR_ST_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + b); };
  auto BAR_ST_FREE = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + grad::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const int half = blockIdx.y;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  // contraction (K) extent of this half, in 8-column K-steps, and where it starts inside the TMA box
  const int ksteps = SYM ? (half == 0 ? grad::SYM_SPLIT / 8 : (136 - grad::SYM_SPLIT) / 8) : 16;
  const int kstep0 = SYM ? (half == 0 ? 0 : (grad::SYM_SPLIT - 64) / 8) : 0;   // half 1 box starts at column 64
  constexpr int BOX_ATOMS = SYM ? 3 : 4;
  constexpr uint32_t ROWS_CTA = PAIR ? BK / 2 : BK;                  // centroid rows of a B tile held here
  constexpr uint32_t ATOM_BYTES = ROWS_CTA * 128;
  constexpr uint32_t TILE_BYTES = BOX_ATOMS * ATOM_BYTES;
  constexpr uint32_t ATOM_DESC = ATOM_BYTES >> 4;
  constexpr uint32_t IDESC_T = make_idesc(PAIR ? 256 : 128, BK);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), 4); mbar_init(BAR_CB_FULL(s), 1);
    }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR_ST_FULL(b), 1); mbar_init(BAR_ST_FREE(b), 4 * NPAIR); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cstack) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + grad::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + grad::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
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
  float zb = 0.f;
  float zrow[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) zrow[j] = 0.f;
  if (warp >= 2) {
    const int64_t r = row0 + prow;
    if (grp == 0) {
      zb = write_z_tiles(gbase, z, r, n, prow, alpha);   // A1/A2 offsets are the forward kernel's
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * 16);
      float nrm = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = __ldg(src + q);
        zrow[4 * q] = v.x; zrow[4 * q + 1] = v.y; zrow[4 * q + 2] = v.z; zrow[4 * q + 3] = v.w;
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
      }
      if (grp == 1) zb = -nrm * alpha;
    }
    // this thread's share of U (hi/lo split) -> TMEM, the A operand of the T GEMM.
    // dense: 64 of this half's 128 columns; packed: up to 40 of this half's 72 / 64 columns.
    const float* urow = u + r * (u_packed ? SYM_COLS : NCOL);   // packed: a symmetric U, [N,144]
    if (!SYM) {
      const float* usrc = urow + half * NHALF + grp * 64;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t h[32], l[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = (r < n) ? __ldg(reinterpret_cast<const float4*>(usrc + cb * 32) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hi = tf32_rna(vv[e]);
            h[4 * q + e] = __float_as_uint(hi);
            l[4 * q + e] = __float_as_uint(vv[e] - hi);
          }
        }
        TMEM_ST32(tmem_base + lane_addr + TM_UHI + grp * 64 + cb * 32, h);
        TMEM_ST32(tmem_base + lane_addr + TM_ULO + grp * 64 + cb * 32, l);
      }
    } else {
      // packed column p = pbase + grp*40 + i  (i < 40); columns past the half's extent are zero
      const int pbase = half == 0 ? 0 : grad::SYM_SPLIT;
      const int pend = half == 0 ? grad::SYM_SPLIT : 136;
      uint32_t h[32], l[32], h2[8], l2[8];
      // (i,j) of packed index p by walking the upper-triangle rows
      int p = pbase + grp * 40;
      int ri = 0, rb = 0;
      while (ri < 15 && p >= rb + (16 - ri)) { rb += 16 - ri; ++ri; }
      int cj = ri + (p - rb);
#pragma unroll
      for (int i = 0; i < 40; ++i) {
        float v = 0.f;
        if (r < n && p < pend) {
          if (u_packed) {
            v = __ldg(urow + p);
            if (cj != ri) v += v;
          } else {
            v = __ldg(urow + ri * 16 + cj);
            if (cj != ri) v += __ldg(urow + cj * 16 + ri);
          }
        }
        const float hi = tf32_rna(v);
        if (i < 32) { h[i] = __float_as_uint(hi); l[i] = __float_as_uint(v - hi); }
        else { h2[i - 32] = __float_as_uint(hi); l2[i - 32] = __float_as_uint(v - hi); }
        ++p; ++cj;
        if (cj == 16) { ++ri; cj = ri; }
      }
      TMEM_ST32(tmem_base + lane_addr + TM_UHI + grp * 40, h);
      TMEM_ST8(tmem_base + lane_addr + TM_UHI + grp * 40 + 32, h2);
      TMEM_ST32(tmem_base + lane_addr + TM_ULO + grp * 40, l);
      TMEM_ST8(tmem_base + lane_addr + TM_ULO + grp * 40 + 32, l2);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (warp == 0) {
    // =========================================================== TMA producer (warp-converged)
    for (int j = 0; j < num_blocks; ++j) {
      const int cs = j % C_STAGES;
      mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE_BYTES);
        const uint32_t dst = base + grad::OFF_C + cs * C_TILE_BYTES;
        if (PAIR) {
          tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32 * (int)rank);
        } else {
          tma_load_2d(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK);
          tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32);
        }
        mbar_expect_tx(BAR_CB_FULL(cs), CN_BYTES + BIAS_BYTES);
        bulk_load_1d(base + OFF_CN + cs * CN_BYTES, cnat + (int64_t)j * BK * 16, CN_BYTES, BAR_CB_FULL(cs));
        bulk_load_1d(base + grad::OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_CB_FULL(cs));
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int it = 2 * j + h, ms = it % M_STAGES;
        mbar_wait(BAR_M_EMPTY(ms), ((it / M_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * TILE_BYTES);
          const CUtensorMap* map = h == 0 ? &tm_mn_hi : &tm_mn_lo;
          const uint32_t dst = base + grad::OFF_M + ms * M_TILE_BYTES;
          const int row = j * BK + (PAIR ? 32 * (int)rank : 0);
          const int atom0 = SYM ? half * 2 : half * 4;
          if (PAIR) tma_load_3d_pair(dst, map, BAR_M_FULL(ms), 0, row, atom0);
          else tma_load_3d(dst, map, BAR_M_FULL(ms), 0, row, atom0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (warp-converged; pair: leader only)
    if (leader) {
      const uint64_t a1_desc = make_desc_sw128(base + grad::OFF_A1);
      const uint64_t a2_desc = make_desc_sw128(base + grad::OFF_A2);
      long long gw_free = 0, gw_c = 0, gw_m = 0, gw_issue = 0;
      (void)gw_free; (void)gw_c; (void)gw_m; (void)gw_issue;
#ifdef RLVAE_TC_PROFILE
      const long long gl0 = clock64();
#endif
      for (int j = 0; j < num_blocks; ++j) {
        PROF_T0();
        const int cs = j % C_STAGES, sb = j & 1;
        const int ms_hi = (2 * j) % M_STAGES, ms_lo = (2 * j + 1) % M_STAGES;
        mbar_wait(BAR_ST_FREE(sb), ((j >> 1) & 1) ^ 1);       // exp groups done with super-block j-2
        PROF_ADD(gw_free);
        mbar_wait(BAR_C_FULL(cs), (j / C_STAGES) & 1);
        PROF_ADD(gw_c);
        mbar_wait(BAR_M_FULL(ms_hi), ((2 * j) / M_STAGES) & 1);
        PROF_ADD(gw_m);
        tc_fence_after();
        const uint32_t s_t = tmem_base + TM_ST + sb * 128;
        const uint32_t t_t = s_t + 64;
        const uint64_t bh = make_desc_sw128(base + grad::OFF_M + ms_hi * M_TILE_BYTES);
        const uint64_t bl = make_desc_sw128(base + grad::OFF_M + ms_lo * M_TILE_BYTES);
        if (elect_one()) {
          issue_gemm1<PAIR>(s_t, a1_desc, a2_desc, make_desc_sw128(base + grad::OFF_C + cs * C_TILE_BYTES));
          // K-step kk = 8 (i,j) columns; box step s = kstep0 + kk; atom = s / 4 (ATOM_BYTES apart)
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_UHI + 8 * kk, bh + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, kk > 0);
          }
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_ULO + 8 * kk, bh + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, 1);
          }
          COMMIT(BAR_M_EMPTY(ms_hi));
        }
        __syncwarp();
        PROF_ADD(gw_issue);
        mbar_wait(BAR_M_FULL(ms_lo), ((2 * j + 1) / M_STAGES) & 1);
        PROF_ADD(gw_m);
        tc_fence_after();
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_UHI + 8 * kk, bl + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, 1);
          }
          COMMIT(BAR_M_EMPTY(ms_lo));
          COMMIT(BAR_ST_FULL(sb));
        }
        __syncwarp();
        PROF_ADD(gw_issue);
      }
#ifdef RLVAE_TC_PROFILE
      if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
        printf("[grad prof] MMA warp per super-block: total %lld | wait ST_FREE %lld  wait C %lld  wait M %lld  issue %lld\n",
               (clock64() - gl0) / num_blocks, gw_free / num_blocks, gw_c / num_blocks, gw_m / num_blocks,
               gw_issue / num_blocks);
#endif
    }
  } else {
    // =========================================================== exp groups (one thread per point)
    const float two_alpha = 2.f * alpha;
    float2 g2[8];
    float su = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) g2[e] = make_float2(0.f, 0.f);
    long long ge_wait = 0, ge_work = 0;
    (void)ge_wait; (void)ge_work;
    for (int j = grp; j < num_blocks; j += 2) {
      PROF_T0();
      const int cs = j % C_STAGES, sb = j & 1;
      const uint32_t st = tmem_base + lane_addr + TM_ST + sb * 128;
      mbar_wait(BAR_CB_FULL(cs), (j / C_STAGES) & 1);
      mbar_wait(BAR_ST_FULL(sb), (j >> 1) & 1);
      tc_fence_after();
      PROF_ADD(ge_wait);
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {
        uint32_t sv[32], tv[32];
        TMEM_LD32(st + rnd * 32, sv);
        TMEM_LD32(st + 64 + rnd * 32, tv);
        const float* bias = reinterpret_cast<const float*>(gbase + grad::OFF_BIAS + cs * BIAS_BYTES) + rnd * 32;
        const float4* crow = reinterpret_cast<const float4*>(gbase + OFF_CN + cs * CN_BYTES) + rnd * 32 * 4;
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float w = ex2_approx(fmaf(__uint_as_float(sv[i]), two_alpha, bias[i] + zb));
          const float uv = w * __uint_as_float(tv[i]);
          su += uv;
          const float2 uu = make_float2(uv, uv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 c4 = crow[i * 4 + q];
            ffma2(g2[2 * q], uu, make_float2(c4.x, c4.y));
            ffma2(g2[2 * q + 1], uu, make_float2(c4.z, c4.w));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_ST_FREE(sb)); else mbar_arrive(BAR_ST_FREE(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
      PROF_ADD(ge_work);
    }
#ifdef RLVAE_TC_PROFILE
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64)
      printf("[grad prof] exp group A per own super-block: wait %lld  work %lld\n", ge_wait / (num_blocks / 2),
             ge_work / (num_blocks / 2));
#endif
    // ---------------------------------------------------------- combine the two groups, then the halves
    // Every TMA / MMA of this CTA has been consumed once both groups leave their loops, so the M
    // ring can be reused as scratch.  (In a pair the peer may still be streaming into ITS smem only.)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* red = reinterpret_cast<float*>(gbase + grad::OFF_M);
    if (grp == 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { red[prow * RED_LD + 2 * e] = g2[e].x; red[prow * RED_LD + 2 * e + 1] = g2[e].y; }
      red[prow * RED_LD + 16] = su;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0) {
      const int64_t r = row0 + prow;
      su += red[prow * RED_LD + 16];
      if (r < n) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float ge = ((e & 1) ? g2[e >> 1].y : g2[e >> 1].x) + red[prow * RED_LD + e];
          atomicAdd(out + r * 16 + e, scale * (ge - zrow[e] * su));   // exactly two addends per element
        }
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

// ==========================================================================================
// Gradient kernel for SYMMETRIC tables with the final contraction on the tensor core as well
// (the FMA contraction of metric_grad_tc_kernel costs ~25 issue slots per (point, centroid) and is
// SM-issue bound at ~2.7k cycles per super-block; measured).  Per 64-centroid super-block j:
//   [GEMM1, T-GEMM](j)  ->  S | T in TMEM buffer j&1                     (as in metric_grad_tc_kernel)
//   exp group           ->  u = exp2(..) * t, split u_hi | u_lo, written over S | T
//   GEMM3(j)            ->  OUT[128 x 32] += u_hi.Ct_hi + u_lo.Ct_hi + u_hi.Ct_lo     (3xTF32, N = 32)
// where Ct = [c_k (16 rows) ; 1 ; 0 ...] so that OUT[:, :16] = sum_k u c_k and OUT[:, 16] = sum_k u.
// OUT is accumulated in two alternating 32-column chunk accumulators (2 super-blocks each) that the
// exp groups fold into fp32 registers -- the same remedy for the accumulator truncation as in the
// forward kernel.  TMEM: [0,80) U_hi, [80,160) U_lo, [160,416) two (S|u_hi 64, T|u_lo 64) buffers,
// [416,480) two OUT chunk accumulators.
// ==========================================================================================
namespace gsym {
constexpr int C_STAGES = 3;
constexpr int M_STAGES = 4;
constexpr uint32_t C_TILE_BYTES = BK * 128;
constexpr uint32_t CT_TILE_BYTES = 2 * 32 * 128;   // [32 rows x 64 centroids] = 2 atoms of 4 KB (pair: 2 KB used each)
constexpr uint32_t BIAS_BYTES = BK * 4;
constexpr uint32_t M_TILE_BYTES = 3 * BK * 128;    // 3 column atoms x 64 centroid rows
constexpr uint32_t OFF_C = OFF_A2 + A_BYTES;
constexpr uint32_t OFF_CT = OFF_C + C_STAGES * C_TILE_BYTES;            // hi then lo per stage
constexpr uint32_t OFF_M = OFF_CT + C_STAGES * 2 * CT_TILE_BYTES;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 5 * C_STAGES + 2 * M_STAGES + 2 + 2 + 2 + 2 + 1;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_UHI = 0, TM_ULO = 80, TM_ST = 160, TM_OUT = 416;
constexpr int RED_LD = 36;
}  // namespace gsym

template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
metric_grad_sym_kernel(const __grid_constant__ CUtensorMap tm_cstack,
                       const __grid_constant__ CUtensorMap tm_mn_hi,
                       const __grid_constant__ CUtensorMap tm_mn_lo,
                       const __grid_constant__ CUtensorMap tm_ct_hi,
                       const __grid_constant__ CUtensorMap tm_ct_lo,
                       const float* __restrict__ z, const float* __restrict__ u,
                       const float* __restrict__ cbias, int64_t n, int num_blocks, float alpha, float scale,
                       float* __restrict__ out, int u_packed) {
  constexpr int C_STAGES = gsym::C_STAGES, M_STAGES = gsym::M_STAGES, RED_LD = gsym::RED_LD;
  constexpr uint32_t C_TILE_BYTES = gsym::C_TILE_BYTES, CT_TILE_BYTES = gsym::CT_TILE_BYTES,
                     BIAS_BYTES = gsym::BIAS_BYTES, M_TILE_BYTES = gsym::M_TILE_BYTES, OFF_C = gsym::OFF_C,
                     OFF_CT = gsym::OFF_CT, OFF_M = gsym::OFF_M, OFF_BIAS = gsym::OFF_BIAS,
                     TM_UHI = gsym::TM_UHI, TM_ULO = gsym::TM_ULO, TM_ST = gsym::TM_ST, TM_OUT = gsym::TM_OUT;
  constexpr int CHUNK = 2;             // super-blocks per OUT chunk accumulator
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + gsym::OFF_BAR;
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
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + gsym::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const int half = blockIdx.y;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  const int ksteps = half == 0 ? grad::SYM_SPLIT / 8 : (136 - grad::SYM_SPLIT) / 8;    // 9 / 8
  const int kstep0 = half == 0 ? 0 : (grad::SYM_SPLIT - 64) / 8;                       // box of half 1 starts at column 64
  constexpr uint32_t ROWS_CTA = PAIR ? BK / 2 : BK;
  constexpr uint32_t ATOM_BYTES = ROWS_CTA * 128;                   // M tile atom held by this CTA
  constexpr uint32_t TILE_BYTES = 3 * ATOM_BYTES;
  constexpr uint32_t ATOM_DESC = ATOM_BYTES >> 4;
  constexpr uint32_t CT_ROWS = PAIR ? 16 : 32;
  constexpr uint32_t CT_ATOM_BYTES = CT_ROWS * 128;
  constexpr uint32_t CT_BYTES = 2 * CT_ATOM_BYTES;                  // per hi / lo tile per CTA
  constexpr uint32_t CT_ATOM_DESC = CT_ATOM_BYTES >> 4;
  constexpr uint32_t IDESC_T = make_idesc(PAIR ? 256 : 128, BK);
  constexpr uint32_t IDESC_3 = make_idesc(PAIR ? 256 : 128, 32);
  const int num_chunks = (num_blocks + CHUNK - 1) / CHUNK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), 4); mbar_init(BAR_B_FULL(s), 1);
      mbar_init(BAR_CT_FULL(s), 1); mbar_init(BAR_CT_EMPTY(s), 1);
    }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR_ST_FULL(b), 1); mbar_init(BAR_U_FULL(b), 4 * NPAIR);
      mbar_init(BAR_CH_FULL(b), 1); mbar_init(BAR_CH_FREE(b), 4 * NPAIR);
    }
    mbar_init(BAR_DONE, 1);
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
                   ::"r"(base + gsym::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + gsym::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
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
  float zb = 0.f;
  float zrow[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) zrow[j] = 0.f;
  if (warp >= 2) {
    const int64_t r = row0 + prow;
    if (grp == 0) {
      zb = write_z_tiles(gbase, z, r, n, prow, alpha);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * 16);
      float nrm = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = __ldg(src + q);
        zrow[4 * q] = v.x; zrow[4 * q + 1] = v.y; zrow[4 * q + 2] = v.z; zrow[4 * q + 3] = v.w;
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
      }
      if (grp == 1) zb = -nrm * alpha;
    }
    // packed symmetric U: Ut_p = U_ij + U_ji (i<j), U_ii; this thread converts 40 packed columns
    const float* urow = u + r * (u_packed ? SYM_COLS : NCOL);   // packed: a symmetric U, [N,144]
    const int pbase = half == 0 ? 0 : grad::SYM_SPLIT;
    const int pend = half == 0 ? grad::SYM_SPLIT : 136;
    uint32_t h[32], l[32], h2[8], l2[8];
    int p = pbase + grp * 40;
    int ri = 0, rb = 0;
    while (ri < 15 && p >= rb + (16 - ri)) { rb += 16 - ri; ++ri; }
    int cj = ri + (p - rb);
#pragma unroll
    for (int i = 0; i < 40; ++i) {
      float v = 0.f;
      if (r < n && p < pend) {
        if (u_packed) {
          v = __ldg(urow + p);
          if (cj != ri) v += v;
        } else {
          v = __ldg(urow + ri * 16 + cj);
          if (cj != ri) v += __ldg(urow + cj * 16 + ri);
        }
      }
      const float hi = tf32_rna(v);
      if (i < 32) { h[i] = __float_as_uint(hi); l[i] = __float_as_uint(v - hi); }
      else { h2[i - 32] = __float_as_uint(hi); l2[i - 32] = __float_as_uint(v - hi); }
      ++p; ++cj;
      if (cj == 16) { ++ri; cj = ri; }
    }
    TMEM_ST32(tmem_base + lane_addr + TM_UHI + grp * 40, h);
    TMEM_ST8(tmem_base + lane_addr + TM_UHI + grp * 40 + 32, h2);
    TMEM_ST32(tmem_base + lane_addr + TM_ULO + grp * 40, l);
    TMEM_ST8(tmem_base + lane_addr + TM_ULO + grp * 40 + 32, l2);
    tmem_wait_st();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_TS(d, a, b, id, acc) do { if (PAIR) mma_ts_pair(d, a, b, id, acc); else mma_ts(d, a, b, id, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (warp == 0) {
    // =========================================================== TMA producer (warp-converged)
    for (int j = 0; j < num_blocks; ++j) {
      const int cs = j % C_STAGES;
      mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE_BYTES);
        const uint32_t dst = base + OFF_C + cs * C_TILE_BYTES;
        if (PAIR) {
          tma_load_2d_pair(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32 * (int)rank);
        } else {
          tma_load_2d(dst, &tm_cstack, BAR_C_FULL(cs), 0, j * BK);
          tma_load_2d(dst + C_TILE_BYTES / 2, &tm_cstack, BAR_C_FULL(cs), 0, j * BK + 32);
        }
        mbar_expect_tx(BAR_B_FULL(cs), BIAS_BYTES);
        bulk_load_1d(base + OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_B_FULL(cs));
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int it = 2 * j + h, ms = it % M_STAGES;
        mbar_wait(BAR_M_EMPTY(ms), ((it / M_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * TILE_BYTES);
          const CUtensorMap* map = h == 0 ? &tm_mn_hi : &tm_mn_lo;
          const uint32_t dst = base + OFF_M + ms * M_TILE_BYTES;
          const int row = j * BK + (PAIR ? 32 * (int)rank : 0);
          if (PAIR) tma_load_3d_pair(dst, map, BAR_M_FULL(ms), 0, row, half * 2);
          else tma_load_3d(dst, map, BAR_M_FULL(ms), 0, row, half * 2);
        }
        __syncwarp();
      }
      // Ct tiles (hi, lo): [32 (pair: 16) rows x 64 centroids] as two 32-centroid atoms each
      mbar_wait(BAR_CT_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_CT_FULL(cs), NPAIR * 2 * CT_BYTES);
        const uint32_t dst = base + OFF_CT + cs * 2 * CT_TILE_BYTES;
        const int row = PAIR ? 16 * (int)rank : 0;
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
  } else if (warp == 1) {
    // =========================================================== MMA issuer (warp-converged; pair: leader only)
    if (leader) {
      const uint64_t a1_desc = make_desc_sw128(base + OFF_A1);
      const uint64_t a2_desc = make_desc_sw128(base + OFF_A2);
      // [GEMM1, T-GEMM] of super-block j into S|T buffer j&1
      auto issue_st = [&](int j) {
        const int cs = j % C_STAGES, sb = j & 1;
        const int ms_hi = (2 * j) % M_STAGES, ms_lo = (2 * j + 1) % M_STAGES;
        mbar_wait(BAR_C_FULL(cs), (j / C_STAGES) & 1);
        mbar_wait(BAR_M_FULL(ms_hi), ((2 * j) / M_STAGES) & 1);
        tc_fence_after();
        const uint32_t s_t = tmem_base + TM_ST + sb * 128;
        const uint32_t t_t = s_t + 64;
        const uint64_t bh = make_desc_sw128(base + OFF_M + ms_hi * M_TILE_BYTES);
        const uint64_t bl = make_desc_sw128(base + OFF_M + ms_lo * M_TILE_BYTES);
        if (elect_one()) {
          issue_gemm1<PAIR>(s_t, a1_desc, a2_desc, make_desc_sw128(base + OFF_C + cs * C_TILE_BYTES));
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_UHI + 8 * kk, bh + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, kk > 0);
          }
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_ULO + 8 * kk, bh + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, 1);
          }
          COMMIT(BAR_M_EMPTY(ms_hi));
        }
        __syncwarp();
        mbar_wait(BAR_M_FULL(ms_lo), ((2 * j + 1) / M_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) {
            const int st = kstep0 + kk;
            MMA_TS(t_t, tmem_base + TM_UHI + 8 * kk, bl + (st >> 2) * ATOM_DESC + 2 * (st & 3), IDESC_T, 1);
          }
          COMMIT(BAR_M_EMPTY(ms_lo));
          COMMIT(BAR_ST_FULL(sb));
        }
        __syncwarp();
      };
      issue_st(0);
      for (int j = 0; j < num_blocks; ++j) {
        // S|T(j+1) first: its buffer was released by GEMM3(j-1), already queued ahead in the pipe
        if (j + 1 < num_blocks) issue_st(j + 1);
        const int cs = j % C_STAGES, sb = j & 1;
        const int chunk = j / CHUNK;
        const int first = (j % CHUNK) == 0;
        if (first && chunk >= 2) mbar_wait(BAR_CH_FREE(chunk & 1), ((chunk >> 1) - 1) & 1);
        mbar_wait(BAR_CT_FULL(cs), (j / C_STAGES) & 1);
        mbar_wait(BAR_U_FULL(sb), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t u_hi = tmem_base + TM_ST + sb * 128;
        const uint32_t u_lo = u_hi + 64;
        const uint32_t acc = tmem_base + TM_OUT + (chunk & 1) * 32;
        const uint64_t ch = make_desc_sw128(base + OFF_CT + cs * 2 * CT_TILE_BYTES);
        const uint64_t cl = make_desc_sw128(base + OFF_CT + cs * 2 * CT_TILE_BYTES + CT_TILE_BYTES);
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
          if ((j % CHUNK) == CHUNK - 1 || j == num_blocks - 1) COMMIT(BAR_CH_FULL(chunk & 1));
        }
        __syncwarp();
      }
      if (elect_one()) COMMIT(BAR_DONE);
      __syncwarp();
    }
  } else {
    // =========================================================== exp groups (one thread per point)
    const float two_alpha = 2.f * alpha;
    float tot[32];                       // this group's share of OUT: chunks of parity grp
#pragma unroll
    for (int e = 0; e < 32; ++e) tot[e] = 0.f;
    auto fold_chunk = [&](int c, bool signal) {
      uint32_t a[32];
      TMEM_LD32(tmem_base + lane_addr + TM_OUT + (c & 1) * 32, a);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) tot[i] += __uint_as_float(a[i]);
      if (signal) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE(c & 1)); else mbar_arrive(BAR_CH_FREE(c & 1)); }
      }
    };
    int next_chunk = grp;                // chunks c with (c & 1) == grp belong to this group
    for (int j = grp; j < num_blocks; j += 2) {
      const int cs = j % C_STAGES, sb = j & 1;
      const uint32_t st = tmem_base + lane_addr + TM_ST + sb * 128;
      mbar_wait(BAR_B_FULL(cs), (j / C_STAGES) & 1);
      mbar_wait(BAR_ST_FULL(sb), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {
        uint32_t sv[32], tv[32];
        TMEM_LD32(st + rnd * 32, sv);
        TMEM_LD32(st + 64 + rnd * 32, tv);
        const float4* bias4 = reinterpret_cast<const float4*>(gbase + OFF_BIAS + cs * BIAS_BYTES) + rnd * 8;
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = bias4[q];
          const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * q + e;
            const float w = ex2_approx(fmaf(__uint_as_float(sv[i]), two_alpha, b4[e] + zb));
            const float uv = w * __uint_as_float(tv[i]);
            const uint32_t uh = __float_as_uint(uv) & 0xFFFFE000u;
            sv[i] = uh;
            tv[i] = __float_as_uint(uv - __uint_as_float(uh));
          }
        }
        TMEM_ST32(st + rnd * 32, sv);
        TMEM_ST32(st + 64 + rnd * 32, tv);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_U_FULL(sb)); else mbar_arrive(BAR_U_FULL(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
      // fold this group's chunks whose last super-block is <= j-1 (their GEMM3 is queued ahead of
      // everything this group waits for next, so the wait is short)
      while (next_chunk < num_chunks && min((next_chunk + 1) * CHUNK - 1, num_blocks - 1) <= j - 1) {
        mbar_wait(BAR_CH_FULL(next_chunk & 1), (next_chunk >> 1) & 1);
        tc_fence_after();
        fold_chunk(next_chunk, true);
        next_chunk += 2;
      }
    }
    mbar_wait(BAR_DONE, 0);
    tc_fence_after();
    while (next_chunk < num_chunks) { fold_chunk(next_chunk, false); next_chunk += 2; }
    // ---------------------------------------------------------- combine the two groups, then the halves
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
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float ge = tot[e] + red[prow * RED_LD + e];
          atomicAdd(out + r * 16 + e, scale * (ge - zrow[e] * su));   // exactly two addends per element
        }
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
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int make_map_2d(PFN_encodeTiled enc, CUtensorMap* map, float* ptr, uint64_t inner, uint64_t outer,
                       uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 4;
  }
  return 0;
}

// natural table [Kpad, 256] viewed as [8 column atoms][Kpad][32]: one box = 4 atoms x 64 rows
static int make_map_atoms(PFN_encodeTiled enc, CUtensorMap* map, float* ptr, uint64_t Kpad, uint32_t row_len = 256,
                          uint32_t box_rows = tc::BK, uint32_t box_atoms = 4) {
  cuuint64_t dims[3] = {32, Kpad, row_len / 32};
  cuuint64_t strides[2] = {row_len * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, box_atoms};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3d) failed with CUresult " + std::to_string((int)r));
    return 4;
  }
  return 0;
}

int tc_build_descriptors(rlvae_tables* t) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  RLVAE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  RLVAE_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  const uint64_t Kpad = (uint64_t)t->Kpad;
  if (int rc = make_map_2d(enc, &t->tm_cstack, t->cstack, 32, Kpad, 32, 32)) return rc;   // 32 centroid rows per box
  // M^T tiles are fetched one 32-centroid swizzle atom (128 B rows) at a time
  if (int rc = make_map_2d(enc, &t->tm_mt_hi, t->Mt_hi, Kpad, tc::NCOL, 32, tc::NHALF)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mt_lo, t->Mt_lo, Kpad, tc::NCOL, 32, tc::NHALF)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mt2_hi, t->Mt_hi, Kpad, tc::NCOL, 32, tc::NHALF / 2)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mt2_lo, t->Mt_lo, Kpad, tc::NCOL, 32, tc::NHALF / 2)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mn_hi, t->Mn_hi, Kpad)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mn_lo, t->Mn_lo, Kpad)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mn2_hi, t->Mn_hi, Kpad, 256, tc::BK / 2, 4)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mn2_lo, t->Mn_lo, Kpad, 256, tc::BK / 2, 4)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct_hi, t->ct_hi, Kpad, 32, 32, 32)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct_lo, t->ct_lo, Kpad, 32, 32, 32)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct2_hi, t->ct_hi, Kpad, 32, 32, 16)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct2_lo, t->ct_lo, Kpad, 32, 32, 16)) return rc;
  // split-fp16 gradient kernel: only the 16 c^T rows (N = 16), pair: 8 rows per CTA
  if (int rc = make_map_2d(enc, &t->tm_ct16_hi, t->ct_hi, Kpad, 32, 32, 16)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct16_lo, t->ct_lo, Kpad, 32, 32, 16)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct8_hi, t->ct_hi, Kpad, 32, 32, 8)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct8_lo, t->ct_lo, Kpad, 32, 32, 8)) return rc;
  return 0;
}

// The split-fp16 gradient kernel contracts with the centred centroids (c - shift)^T [16, Kpad]: its
// epilogue forms sum_k coef_k (c_k - z) as sum_k coef_k c~_k - z~ sum_k coef_k, which would cancel
// badly for a latent cloud far from the origin if c and z were used as given.
int tc_build_ct_centred_descriptors(rlvae_tables* t) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  RLVAE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  RLVAE_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  const uint64_t Kpad = (uint64_t)t->Kpad;
  if (int rc = make_map_2d(enc, &t->tm_ct16_hi, t->ctc_hi, Kpad, 16, 32, 16)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct16_lo, t->ctc_lo, Kpad, 16, 32, 16)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct8_hi, t->ctc_hi, Kpad, 16, 32, 8)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_ct8_lo, t->ctc_lo, Kpad, 16, 32, 8)) return rc;
  if (t->bt_hi != nullptr && t->bt_lo != nullptr) {     // pythae table: same shape, same boxes
    if (int rc = make_map_2d(enc, &t->tm_bt16_hi, t->bt_hi, Kpad, 16, 32, 16)) return rc;
    if (int rc = make_map_2d(enc, &t->tm_bt16_lo, t->bt_lo, Kpad, 16, 32, 16)) return rc;
    if (int rc = make_map_2d(enc, &t->tm_bt8_hi, t->bt_hi, Kpad, 16, 32, 8)) return rc;
    if (int rc = make_map_2d(enc, &t->tm_bt8_lo, t->bt_lo, Kpad, 16, 32, 8)) return rc;
  }
  return 0;
}

int tc_build_sym_descriptors(rlvae_tables* t) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  RLVAE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  RLVAE_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  const uint64_t Kpad = (uint64_t)t->Kpad;
  // packed-transposed tables [144, Kpad]; a box is one 32-centroid atom x 80 packed rows (rows past
  // 144 are out of bounds and arrive as zeros)
  if (int rc = make_map_2d(enc, &t->tm_mts_hi, t->Mts_hi, Kpad, tc::SYM_COLS, 32, tc::SYM_H0)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mts_lo, t->Mts_lo, Kpad, tc::SYM_COLS, 32, tc::SYM_H0)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mts2_hi, t->Mts_hi, Kpad, tc::SYM_COLS, 32, tc::SYM_H0 / 2)) return rc;
  if (int rc = make_map_2d(enc, &t->tm_mts2_lo, t->Mts_lo, Kpad, tc::SYM_COLS, 32, tc::SYM_H0 / 2)) return rc;
  // packed natural tables [Kpad, 160] of the gradient kernel: boxes of 3 column atoms
  if (int rc = make_map_atoms(enc, &t->tm_mns_hi, t->Mns_hi, Kpad, kSymNatCols, tc::BK, 3)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mns_lo, t->Mns_lo, Kpad, kSymNatCols, tc::BK, 3)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mns2_hi, t->Mns_hi, Kpad, kSymNatCols, tc::BK / 2, 3)) return rc;
  if (int rc = make_map_atoms(enc, &t->tm_mns2_lo, t->Mns_lo, Kpad, kSymNatCols, tc::BK / 2, 3)) return rc;
  return 0;
}

// CTA pairs (cta_group::2) are the default; RLVAE_TC_PAIR=0 selects the single-CTA kernels.
static bool use_pairs() {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_PAIR");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();   // initialised once, thread-safe (C++11 magic static)
  return v == 1;
}

template <bool SYM, bool PAIR>
static int launch_fwd(const CUtensorMap& c, const CUtensorMap& hi, const CUtensorMap& lo, const rlvae_tables* t,
                      const float* z, int64_t n, float* out, cudaStream_t s) {
  auto kern = tc::inverse_metric_tc_kernel<SYM, PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::fwd::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;               // clusters of two point tiles
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 2, 1);
  cfg.blockDim = dim3(tc::fwd::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::fwd::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float alpha = 1.4426950408889634f / t->T2;
  const float* cbias = t->cbias;
  const int nb = t->Kpad / tc::BK;
  const float lambda = t->lambda;
  static int chunk_blocks = 0;       // super-blocks per tensor-core accumulation chunk (RLVAE_TC_CHUNK)
  if (chunk_blocks == 0) {
    const char* e = getenv("RLVAE_TC_CHUNK");
    chunk_blocks = (e != nullptr && atoi(e) >= 1 && atoi(e) <= 16) ? atoi(e) : tc::CHUNK_BLOCKS;
  }
  const int cb = chunk_blocks;
  RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, c, hi, lo, z, cbias, n, nb, cb, alpha, lambda, out));
  return 0;
}

int launch_inverse_metric_tc_sym(const rlvae_tables* t, const float* z, int64_t n, float* packed,
                                 cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->tensor_capable && t->symmetric && t->Mts_hi != nullptr,
                "symmetric tensor path needs latent_dim == 16 and symmetric tables");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed) & 15) == 0,
                "tensor path needs 16-byte aligned z and output");
  return use_pairs() ? launch_fwd<true, true>(t->tm_cstack, t->tm_mts2_hi, t->tm_mts2_lo, t, z, n, packed, s)
                     : launch_fwd<true, false>(t->tm_cstack, t->tm_mts_hi, t->tm_mts_lo, t, z, n, packed, s);
}

int launch_inverse_metric_tc(const rlvae_tables* t, const float* z, int64_t n, float* ginv,
                             cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->tensor_capable, "tensor path needs latent_dim == 16");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(ginv) & 15) == 0,
                "tensor path needs 16-byte aligned z and output");
  return use_pairs() ? launch_fwd<false, true>(t->tm_cstack, t->tm_mt2_hi, t->tm_mt2_lo, t, z, n, ginv, s)
                     : launch_fwd<false, false>(t->tm_cstack, t->tm_mt_hi, t->tm_mt_lo, t, z, n, ginv, s);
}

template <bool SYM, bool PAIR>
static int launch_grad(const CUtensorMap& c, const CUtensorMap& hi, const CUtensorMap& lo, const rlvae_tables* t,
                       const float* z, const float* u, int64_t n, float scale, float* out, cudaStream_t s,
                       int u_packed) {
  auto kern = tc::metric_grad_tc_kernel<SYM, PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::grad::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 2, 1);
  cfg.blockDim = dim3(tc::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::grad::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float alpha = 1.4426950408889634f / t->T2;
  const float* cnat = t->c;
  const float* cbias = t->cbias;
  const int nb = t->Kpad / tc::BK;
  RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, c, hi, lo, z, u, cnat, cbias, n, nb, alpha, scale, out, u_packed));
  return 0;
}

template <bool PAIR>
static int launch_grad_sym(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale,
                           float* out, cudaStream_t s, int u_packed) {
  auto kern = tc::metric_grad_sym_kernel<PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::gsym::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, 2, 1);
  cfg.blockDim = dim3(tc::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::gsym::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float alpha = 1.4426950408889634f / t->T2;
  const float* cbias = t->cbias;
  const int nb = t->Kpad / tc::BK;
  if (PAIR) {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_cstack, t->tm_mns2_hi, t->tm_mns2_lo, t->tm_ct2_hi,
                                     t->tm_ct2_lo, z, u, cbias, n, nb, alpha, scale, out, u_packed));
  } else {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_cstack, t->tm_mns_hi, t->tm_mns_lo, t->tm_ct_hi,
                                     t->tm_ct_lo, z, u, cbias, n, nb, alpha, scale, out, u_packed));
  }
  return 0;
}

// RLVAE_TC_GRAD = h16 (default: split-fp16 T GEMM, one CTA per 128 points) | tf32 (3xTF32, tensor-core
// final contraction) | fma (3xTF32, final contraction on the FMA pipe)
static int grad_variant() {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_GRAD");
    return (e == nullptr) ? 2 : (e[0] == 'f' ? 0 : (e[0] == 't' ? 1 : 2));
  }();   // initialised once, thread-safe (C++11 magic static)
  return v;
}

int launch_metric_grad_tc(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale,
                          float* out, cudaStream_t s, int u_packed) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 16 && t->tensor_capable, "tensor path needs latent_dim == 16");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0,
                "tensor path needs 16-byte aligned z and u");
  const bool sym = t->symmetric && t->Mns_hi != nullptr;
  RLVAE_REQUIRE(sym || !u_packed, "packed U needs symmetric tables");
  if (sym && grad_variant() == 2 && t->Mnh_hi != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    return launch_metric_grad_h16(t, z, u, n, scale, out, s, u_packed);
  RLVAE_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * 16, s));   // the two column halves add
  if (sym && grad_variant() >= 1) {
    return use_pairs() ? launch_grad_sym<true>(t, z, u, n, scale, out, s, u_packed)
                       : launch_grad_sym<false>(t, z, u, n, scale, out, s, u_packed);
  }
  if (sym) {
    return use_pairs() ? launch_grad<true, true>(t->tm_cstack, t->tm_mns2_hi, t->tm_mns2_lo, t, z, u, n, scale, out, s, u_packed)
                       : launch_grad<true, false>(t->tm_cstack, t->tm_mns_hi, t->tm_mns_lo, t, z, u, n, scale, out, s, u_packed);
  }
  return use_pairs() ? launch_grad<false, true>(t->tm_cstack, t->tm_mn2_hi, t->tm_mn2_lo, t, z, u, n, scale, out, s, 0)
                     : launch_grad<false, false>(t->tm_cstack, t->tm_mn_hi, t->tm_mn_lo, t, z, u, n, scale, out, s, 0);
}

}  // namespace rlvae
