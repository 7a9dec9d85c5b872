// Split-fp16 tcgen05 forward kernel for latent_dim == 64, symmetric tables (sm_100a):
// BASELINE.json configs[4] ("large-metric stress": d = 64, K = 50k, tables streamed via TMA).
//
//   G^{-1}[n] = sum_k exp(-||z_n - c_k||^2 / T^2) M_k + lambda I      ref src/models/components/metric_tensor.py:115-135
//
// Same skeleton as inverse_metric_h16_kernel (rlvae_tc16.cu), with the two changes d = 64 forces:
//  * the 64 x 65 / 2 = 2080 packed columns do not fit one CTA's TMEM, so blockIdx.y selects a tile
//    of 128 packed columns (17 tiles, 2176 padded columns); every column tile recomputes the
//    distance GEMM and the exp stage for its 128 points;
//  * that makes the distance GEMM a third of the work, so it runs on kind::f16 as well:
//    z' = 2^ez z (per point), c' = 2^ec c (per table), both split hi + lo in fp16,
//    S' = z'_hi.c'_hi + z'_hi.c'_lo + z'_lo.c'_hi  (12 MMAs, K = 16 each, A operand = z' in TMEM),
//    S = 2^-(ez+ec) S' is folded into the exponent's FFMA.  Error class: 2^-22 relative to
//    |z||c|, the same as the 3xTF32 split of the d = 16 kernels.
// Per 64-centroid super-block and column tile: GEMM1 12 x 32 + GEMM2 12 x 64 = 1152 tensor cycles.
// The per-point 64 x 64 inverse / log det stays a separate kernel (spd64_*_kernel, one warp per matrix;
// pivoting fallback batched_inverse_kernel<64>): a 2080-entry matrix does not fit one thread's registers.
//
//   TMEM: [0,192) three S/P buffers, [192,320) / [320,448) chunk accumulators (N = 128),
//         [448,480) z'_hi, [480,512) z'_lo (64 dims as packed fp16).
//   Output: packed [N, 2176] fp32 (entry 64 i - i (i-1)/2 + (j - i) = element (i <= j), WITHOUT lambda;
//           unpack_sym64_kernel expands to [N,64,64] and adds lambda on the diagonal).
#include <type_traits>

#include "rlvae_tc_common.cuh"

namespace rlvae {
namespace tc {
namespace h64 {

constexpr int D = 64;
constexpr int THREADS = 512;
constexpr int C_STAGES = 3;
constexpr int SP_BUFS = 3;
constexpr int M_STAGES = 3;
constexpr int AHEAD = 3;
constexpr int NT = 128;                                   // packed columns per CTA = MMA N of GEMM2
constexpr int NPACK = 2080;                               // 64 * 65 / 2
constexpr int NPAD = 2176;                                // 17 tiles of 128
constexpr uint32_t C_TILE64 = 2 * BK * 128;               // [64 centroids x (hi atom | lo atom)] fp16
constexpr uint32_t M_HALF_BYTES = NT * 128;               // [128 rows x 64 centroids] fp16 (pair: 64 rows used)
constexpr uint32_t M_TILE_BYTES = 2 * M_HALF_BYTES;       // hi tile, then lo tile
constexpr uint32_t OFF_C = 0;
constexpr uint32_t OFF_M = OFF_C + C_STAGES * C_TILE64;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;
constexpr int NUM_BARS = 3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 4;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
constexpr int OUT_LD = 132;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TILE_M * OUT_LD * 4 <= M_STAGES * M_TILE_BYTES, "epilogue staging must fit the M ring");
constexpr uint32_t TM_SP = 0, TM_ACC = 192, TM_ZHI = 448, TM_ZLO = 480;
constexpr float P_SHIFT = 14.f;

}  // namespace h64

template <bool PAIR>
__global__ void __launch_bounds__(h64::THREADS, 1)
inverse_metric_h64_kernel(const __grid_constant__ CUtensorMap tm_c64,
                          const __grid_constant__ CUtensorMap tm_mh_hi,
                          const __grid_constant__ CUtensorMap tm_mh_lo,
                          const float* __restrict__ z, const float* __restrict__ cbias, int64_t n,
                          int num_blocks, float alpha /* log2(e)/T^2 */, float c_unscale /* 2^-ec */,
                          float out_scale /* 2^-(14+e) */, const float* __restrict__ cshift /* [64] table centre */,
                          float* __restrict__ out /* [N, 2176] */) {
  constexpr int C_STAGES = h64::C_STAGES, SP_BUFS = h64::SP_BUFS, M_STAGES = h64::M_STAGES, AHEAD = h64::AHEAD,
                NT = h64::NT, OUT_LD = h64::OUT_LD, D = h64::D;
  constexpr uint32_t M_TILE_BYTES = h64::M_TILE_BYTES, M_HALF_BYTES = h64::M_HALF_BYTES, C_TILE64 = h64::C_TILE64,
                     TM_SP = h64::TM_SP, TM_ACC = h64::TM_ACC, TM_ZHI = h64::TM_ZHI, TM_ZLO = h64::TM_ZLO;
  constexpr float P_SHIFT = h64::P_SHIFT;
  constexpr int CB = 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

  const uint32_t bar0 = base + h64::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_BIAS_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };
  auto BAR_M_FULL = [&](int s) { return bar0 + 8u * (3 * C_STAGES + s); };
  auto BAR_M_EMPTY = [&](int s) { return bar0 + 8u * (3 * C_STAGES + M_STAGES + s); };
  auto BAR_S_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + b); };
  auto BAR_P_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + SP_BUFS + b); };
  auto BAR_CH_FULL = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + b); };
  auto BAR_CH_FREE = [&](int b) { return bar0 + 8u * (3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + h64::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const int col_tile = blockIdx.y;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  constexpr int ROWS_BOX = NT / NPAIR;                    // B-tile rows held by this CTA
  constexpr uint32_t TILE_BYTES = ROWS_BOX * 128;
  constexpr int C_ROWS = BK / NPAIR;                      // centroid rows of a C tile held by this CTA
  constexpr uint32_t C_ATOM_BYTES = C_ROWS * 128;
  constexpr uint32_t C_ATOM_DESC = C_ATOM_BYTES >> 4;
  constexpr uint32_t IDESC_G1 = make_idesc_f16(PAIR ? 256 : 128, BK);
  constexpr uint32_t IDESC_G2 = make_idesc_f16(PAIR ? 256 : 128, NT);
  const int row_cta = col_tile * NT + (PAIR ? (int)rank * ROWS_BOX : 0);
  const int num_chunks = (num_blocks + CB - 1) / CB;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), 4); mbar_init(BAR_BIAS_FULL(s), 1);
    }
    for (int s = 0; s < SP_BUFS; ++s) { mbar_init(BAR_S_FULL(s), 1); mbar_init(BAR_P_FULL(s), 4 * NPAIR); }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR_CH_FULL(b), 1); mbar_init(BAR_CH_FREE(b), 4 * NPAIR); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mh_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mh_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + h64::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + h64::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }

  const int quarter = warp & 3;
  const int prow = quarter * 32 + lane;
  tc_fence_before();
  __syncthreads();            // TMEM base published (the exp threads store z into TMEM below)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  float zb = 0.f, s_scale = 0.f;   // exponent = S' * s_scale + bias_k + zb
  if (wg == 1 || wg == 2) {
    const int64_t r = row0 + prow;
    float nrm = 0.f, zmax = 0.f;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * D);
#pragma unroll 4
      for (int q = 0; q < D / 4; ++q) {
        float4 v = __ldg(src + q);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);   // expanded form about the table centre
        v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
        zmax = fmaxf(zmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
    }
    // z' = 2^ez z with max|z'| in [2^13, 2^14)
    int ez = 0;
    if (zmax > 0.f && zmax < 3.0e38f) {
      const int ex = (int)((__float_as_uint(zmax) >> 23) & 0xffu) - 126;    // zmax = f 2^ex, f in [0.5, 1)
      ez = 14 - ex;
      ez = ez > 50 ? 50 : (ez < -50 ? -50 : ez);
    }
    const float zsc = __uint_as_float((uint32_t)(ez + 127) << 23);
    zb = -nrm * alpha + P_SHIFT;
    s_scale = 2.f * alpha * c_unscale * __uint_as_float((uint32_t)(127 - ez) << 23);
    if (wg == 1) {                   // exp group A writes the A operand of GEMM1 (split fp16) into TMEM
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      uint32_t hi[32], lo[32];
      const float4* src = reinterpret_cast<const float4*>(z + r * D);
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) {
          v = __ldg(src + q);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
          v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
        }
        split_pair(v.x * zsc, v.y * zsc, hi[2 * q], lo[2 * q]);
        split_pair(v.z * zsc, v.w * zsc, hi[2 * q + 1], lo[2 * q + 1]);
      }
      TMEM_ST32(tmem_base + lane_addr + TM_ZHI, hi);
      TMEM_ST32(tmem_base + lane_addr + TM_ZLO, lo);
      tmem_wait_st();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_G1(d, a, b, acc) do { if (PAIR) mma_ts_f16_pair(d, a, b, IDESC_G1, acc); else mma_ts_f16(d, a, b, IDESC_G1, acc); } while (0)
#define MMA_H(d, a, b, acc) do { if (PAIR) mma_ts_f16_pair(d, a, b, IDESC_G2, acc); else mma_ts_f16(d, a, b, IDESC_G2, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (wg == 0) {
    reg_dec<72>();
    if (warp == 0) {
      // =========================================================== TMA producer 1: centroid tiles + bias
      for (int j = 0; j < num_blocks; ++j) {
        const int cs = j % C_STAGES;
        mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE64);
          const uint32_t dst = base + h64::OFF_C + cs * C_TILE64;
          // box = [64 fp16] x [C_ROWS centroid rows] x [2 atoms: hi, lo]
          if (PAIR) tma_load_3d_pair(dst, &tm_c64, BAR_C_FULL(cs), 0, j * BK + C_ROWS * (int)rank, 0);
          else tma_load_3d(dst, &tm_c64, BAR_C_FULL(cs), 0, j * BK, 0);
          mbar_expect_tx(BAR_BIAS_FULL(cs), BIAS_BYTES);
          bulk_load_1d(base + h64::OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_BIAS_FULL(cs));
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // =========================================================== TMA producer 2: table tiles of this column tile
      for (int jm = 0; jm < num_blocks; ++jm) {
        const int ms = jm % M_STAGES;
        mbar_wait(BAR_M_EMPTY(ms), ((jm / M_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * 2 * TILE_BYTES);
          const uint32_t dst = base + h64::OFF_M + ms * M_TILE_BYTES;
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
      // =========================================================== MMA issuer (6x unrolled, see rlvae_tc16.cu)
      const uint64_t c_desc0 = make_desc_sw128(base + h64::OFF_C);
      const uint64_t m_desc0 = make_desc_sw128(base + h64::OFF_M);
      auto gemm1 = [&](auto CSc, auto SBc) {
        constexpr int cs = decltype(CSc)::value, sb = decltype(SBc)::value;
        if (elect_one()) {
          const uint32_t d = tmem_base + TM_SP + sb * 64;
          const uint64_t bh = c_desc0 + ((cs * C_TILE64) >> 4);       // hi atom; lo atom C_ATOM_DESC further
          const uint64_t bl = bh + C_ATOM_DESC;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_G1(d, tmem_base + TM_ZHI + 8 * kk, bh + 2 * kk, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_G1(d, tmem_base + TM_ZHI + 8 * kk, bl + 2 * kk, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_G1(d, tmem_base + TM_ZLO + 8 * kk, bh + 2 * kk, 1);
          COMMIT(BAR_S_FULL(sb));
        }
        __syncwarp();
      };
      uint32_t free_phase = 0;
      auto block = [&](auto Jc, const int j, const uint32_t qodd /* (j / 6) & 1 */) {
        constexpr int J = decltype(Jc)::value;
        constexpr int first = (J % CB) == 0, sb = J % SP_BUFS;
        const uint32_t ab = ((J / CB) & 1) ^ qodd;
        constexpr int ms = J % M_STAGES;
        if (first && j >= 2 * CB) {
          mbar_wait(BAR_CH_FREE(ab), (free_phase >> ab) & 1u);
          free_phase ^= 1u << ab;
        }
        tc_fence_after();
        const uint32_t p = tmem_base + TM_SP + sb * 64;
        const uint32_t acc = tmem_base + TM_ACC + ab * NT;
        const uint64_t bh = m_desc0 + ((ms * M_TILE_BYTES) >> 4);
        const uint64_t bl = bh + (M_HALF_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_H(acc, p + (kk >> 1) * 32 + (kk & 1) * 8, bh + 2 * kk, !(first && kk == 0));
        }
        __syncwarp();
        if (j + 1 < num_blocks) mbar_wait(BAR_M_FULL((J + 1) % M_STAGES), ((J + 1) / M_STAGES) & 1);
        if (j + AHEAD < num_blocks) mbar_wait(BAR_C_FULL((J + AHEAD) % C_STAGES), ((J + AHEAD) / C_STAGES) & 1);
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
        if (j + 1 < num_blocks) mbar_wait(BAR_P_FULL((J + 1) % SP_BUFS), ((J + 1) / SP_BUFS) & 1);
      };
      if (0 < num_blocks) { mbar_wait(BAR_C_FULL(0), 0); tc_fence_after(); gemm1(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{}); }
      if (1 < num_blocks) { mbar_wait(BAR_C_FULL(1), 0); tc_fence_after(); gemm1(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{}); }
      if (2 < num_blocks) { mbar_wait(BAR_C_FULL(2), 0); tc_fence_after(); gemm1(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{}); }
      mbar_wait(BAR_M_FULL(0), 0);
      mbar_wait(BAR_P_FULL(0), 0);
      static_assert(6 % CB == 0 && 6 % SP_BUFS == 0 && 6 % C_STAGES == 0 && 6 % M_STAGES == 0 && AHEAD == 3,
                    "the 6x unrolled issue loop assumes these periods");
      uint32_t qodd = 0;
      for (int j0 = 0; j0 < num_blocks; j0 += 6, qodd ^= 1u) {
#define RLVAE_BLK(J) if (j0 + J < num_blocks) block(std::integral_constant<int, J>{}, j0 + J, qodd);
        RLVAE_BLK(0) RLVAE_BLK(1) RLVAE_BLK(2) RLVAE_BLK(3) RLVAE_BLK(4) RLVAE_BLK(5)
#undef RLVAE_BLK
      }
    }
  } else if (wg == 1 || wg == 2) {
    // =========================================================== exp groups (one thread per point)
    reg_dec<104>();
    const int grp = wg - 1;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    for (int j = grp; j < num_blocks; j += 2) {
      const int cs = j % C_STAGES, sb = j % SP_BUFS;
      const uint32_t sp = tmem_base + lane_addr + TM_SP + sb * 64;
      mbar_wait(BAR_BIAS_FULL(cs), (j / C_STAGES) & 1);
      mbar_wait(BAR_S_FULL(sb), (j / SP_BUFS) & 1);
      tc_fence_after();
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {
        uint32_t s[32], ph[16], pl[16];
        TMEM_LD32(sp + rnd * 32, s);
        const float4* bias4 = reinterpret_cast<const float4*>(gbase + h64::OFF_BIAS + cs * BIAS_BYTES) + rnd * 8;
        tmem_wait_ld();
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
    }
  } else {
    // =========================================================== fold group: fp32 running total + output
    reg_inc<232>();
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float total[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) total[i] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int ab = c & 1;
      mbar_wait(BAR_CH_FULL(ab), (c >> 1) & 1);
      tc_fence_after();
      const uint32_t src = tmem_base + lane_addr + TM_ACC + ab * NT;
#pragma unroll
      for (int cb = 0; cb < NT / 32; ++cb) {
        uint32_t a[32];
        TMEM_LD32(src + cb * 32, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) total[cb * 32 + i] += __uint_as_float(a[i]);
      }
      if (c + 2 < num_chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE(ab)); else mbar_arrive(BAR_CH_FREE(ab)); }
      }
    }
    // ---------------------------------------------------------- epilogue (all TMA / MMA work is complete)
    float* stage = reinterpret_cast<float*>(gbase + h64::OFF_M);
    const int t = threadIdx.x - 384;
    const int64_t rows_here = (n - row0 < TILE_M) ? (n - row0) : TILE_M;
#pragma unroll
    for (int q = 0; q < NT / 4; ++q)
      *reinterpret_cast<float4*>(stage + prow * OUT_LD + q * 4) =
          make_float4(total[4 * q] * out_scale, total[4 * q + 1] * out_scale, total[4 * q + 2] * out_scale,
                      total[4 * q + 3] * out_scale);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float* dst = out + row0 * h64::NPAD + col_tile * NT;
    for (int i = t; i < (int)rows_here * (NT / 4); i += 128) {
      const int r = i / (NT / 4), c4 = i - r * (NT / 4);
      *reinterpret_cast<float4*>(dst + (int64_t)r * h64::NPAD + c4 * 4) =
          *reinterpret_cast<const float4*>(stage + r * OUT_LD + c4 * 4);
    }
  }
#undef MMA_G1
#undef MMA_H
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
// Gradient / backward kernel for latent_dim == 64, symmetric tables (BASELINE.json configs[4]):
//   out[n,:] = scale * sum_k w_nk <U_n, M_k> (c_k - z_n)            (rlvae_metric_grad, K4)
// The 2080 packed columns of <U, M_k> are linear in the column range, so -- like the forward kernel --
// blockIdx.y selects a tile of 128 packed columns and the CTA produces the PARTIAL result of its tile
//   out^(c)[n,:] = scale * sum_k w_nk t^(c)_nk (c_k - z_n),   t^(c)_nk = sum_{p in tile c} Ut_np M_kp
// into partial[c][n][:]; reduce_partials64_kernel adds the 17 tiles in a fixed order (deterministic).
// Every column tile recomputes the distance GEMM and the exp stage for its 128 points (as the forward
// kernel does).  Per 64-centroid super-block, all on kind::f16 with the split hi + lo operands:
//   GEMM1  S[128 x 64]  = Z'.C'^T           12 MMAs, A = z' (hi | lo) resident in TMEM
//   T-GEMM T[128 x 64]  = U'.M'^T           24 MMAs (K = 128 columns), A = U' (hi | lo) resident in TMEM,
//                                           B = natural table tiles [64 centroids x 128 columns]
//   exp    u = w t; the two threads that own a point (one per exp group, 32 centroids each) agree on one power
//          of two per block through shared memory, u' = 2^s u has its largest element in [2^13, 2^14), split
//          hi | lo fp16, written over S (the layout of the forward kernel's P).  A fixed scale would lose the
//          RELATIVE accuracy of points whose weights are uniformly tiny (far from every centroid).
//   GEMM3  OUT[128 x 64] = u'.C'             12 MMAs, B = (c - shift)^T tiles [64 dims x 64 centroids]; a fresh
//          accumulation per block, folded as 2^-s OUT into fp32 registers (group g: output dims [32g, 32g+32))
// = 384 + 768 + 384 tensor cycles (A from shared memory cost 48 instead of 32 cycles per GEMM1 MMA -- measured
// 126 -> 120 ms per 2^17 points at K = 50k -- so z' lives in TMEM and OUT has ONE chunk accumulator: its fold
// runs under the next block's T-GEMM anyway).  u' <= 2^14 because |t'| <= 128 * 2^28; both fp16 operands of GEMM3 are exact
// hi + lo sums down to 2^-25 absolute (2^-39 of the largest u').
// TMEM: [0,64) U'_hi, [64,128) U'_lo, [128,384) two (S | T) buffers, [384,448) the OUT chunk accumulator,
//       [448,480) z'_hi, [480,512) z'_lo.
// Warps: 0 TMA (C, bias, C^T), 1 MMA issuer (GEMM1 + T), 2-5 / 6-9 exp groups (32 centroids each of every
// super-block), 10 TMA (table tiles), 11 MMA issuer (GEMM3).
// ==========================================================================================
namespace g64 {
constexpr int D = 64;
constexpr int THREADS = 384;
constexpr int C_STAGES = 4;
constexpr int M_STAGES = 2;
constexpr int NT = 128;                                   // packed columns per CTA = K of the T GEMM
constexpr int KSTEPS = NT / 16;
constexpr int NTILES = h64::NPAD / NT;                    // 17
constexpr uint32_t C_TILE64 = h64::C_TILE64;              // [64 centroids x (hi atom | lo atom)]
constexpr uint32_t CT_HALF = D * 128;                     // [64 dims x 64 centroids] fp16 (pair: 32 rows used)
constexpr uint32_t CT_TILE = 2 * CT_HALF;                 // hi, lo
constexpr uint32_t M_HALF = 2 * BK * 128;                 // 2 column atoms (64 fp16 each) x 64 centroid rows
constexpr uint32_t M_TILE = 2 * M_HALF;                   // hi, lo
constexpr uint32_t OFF_C = 0;
constexpr uint32_t OFF_CT = OFF_C + C_STAGES * C_TILE64;
constexpr uint32_t OFF_M = OFF_CT + C_STAGES * CT_TILE;
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE;
constexpr uint32_t OFF_UMAX = OFF_BIAS + C_STAGES * BIAS_BYTES;   // [2 block parities][2 groups][128 points] block maxima
constexpr uint32_t OFF_BAR = OFF_UMAX + 2 * 2 * TILE_M * 4;
constexpr int NUM_BARS = 5 * C_STAGES + 2 * M_STAGES + 11;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_UHI = 0, TM_ULO = 64, TM_ST = 128, TM_OUT = 384, TM_ZHI = 448, TM_ZLO = 480;
}  // namespace g64

// U [N,64,64] (any matrix) -> packed Ut [N,2176] fp32 (Ut_p = U_ij + U_ji for the packed column p = (i < j), U_ii on
// the diagonal, zeros in the padding columns) + the per-point exponent eU with max|2^eU Ut| in [2^13, 2^14).
// One CTA per point: the matrix goes through shared memory (coalesced 128-bit loads; row stride 65 so the
// transposed reads hit distinct banks), the packed row leaves coalesced -- the gradient kernel's prologue then reads
// 256 contiguous bytes per thread and column tile instead of 128 scattered words (7.5 -> ~1 ms per 2^17 points).
__global__ void __launch_bounds__(256)
pack_u64_kernel(const float* __restrict__ u, int64_t n, float* __restrict__ up, int* __restrict__ eu) {
  __shared__ float sm[64 * 65];
  __shared__ float wmax[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(u + p * 4096);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i4 = tid + 256 * q;                 // float4 index: row i4 / 16, columns 4 (i4 % 16) ..
      const float4 v = __ldg(src + i4);
      float* dst = sm + (i4 >> 4) * 65 + ((i4 & 15) << 2);
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    float m = 0.f;
    float* dst = up + p * h64::NPAD;
    for (int c = tid; c < h64::NPAD; c += 256) {
      float v = 0.f;
      if (c < h64::NPACK) {
        // packed column c -> (i, j), i <= j: base(i) = i (129 - i) / 2 <= c < base(i + 1)
        int i = (int)((129.f - sqrtf(16641.f - 8.f * (float)c)) * 0.5f);
        i = i < 0 ? 0 : (i > 63 ? 63 : i);
        while (i < 63 && ((i + 1) * (128 - i)) / 2 <= c) ++i;
        while (i > 0 && (i * (129 - i)) / 2 > c) --i;
        const int j = i + (c - (i * (129 - i)) / 2);
        v = sm[i * 65 + j];
        if (j != i) v += sm[j * 65 + i];
      }
      dst[c] = v;
      m = fmaxf(m, fabsf(v));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) wmax[warp] = m;
    __syncthreads();                                // (also: every thread is done reading sm)
    if (tid == 0) {
      for (int w = 1; w < 8; ++w) m = fmaxf(m, wmax[w]);
      int e = 0;
      if (m > 0.f && m < 3.0e38f) {
        const int ex = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126;    // m = f 2^ex, f in [0.5, 1)
        e = 14 - ex;
        e = e > 50 ? 50 : (e < -50 ? -50 : e);
      }
      eu[p] = e;
    }
  }
}

__global__ void reduce_partials64_kernel(const float* __restrict__ partial, int64_t n, int ntiles,
                                         float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index into [n, 64]
  if (i >= n * 16) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < ntiles; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (int64_t)c * n * 64) + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

template <bool PAIR>
__global__ void __launch_bounds__(g64::THREADS, 1)
metric_grad_h64_kernel(const __grid_constant__ CUtensorMap tm_c64,
                       const __grid_constant__ CUtensorMap tm_mn_hi,
                       const __grid_constant__ CUtensorMap tm_mn_lo,
                       const __grid_constant__ CUtensorMap tm_ct_hi,
                       const __grid_constant__ CUtensorMap tm_ct_lo,
                       const float* __restrict__ z, const float* __restrict__ up /* [N,2176] packed Ut (pack_u64_kernel) */,
                       const int* __restrict__ eu /* [N] per-point exponent of U' */,
                       const float* __restrict__ cbias, int64_t n, int num_blocks, float alpha,
                       float scale /* includes 2^-eM */, float c_unscale /* 2^-ec */,
                       const float* __restrict__ cshift /* [64] */, float* __restrict__ partial /* [17, N, 64] */) {
  constexpr int C_STAGES = g64::C_STAGES, M_STAGES = g64::M_STAGES, KSTEPS = g64::KSTEPS, D = g64::D, NT = g64::NT;
  constexpr uint32_t C_TILE64 = g64::C_TILE64, CT_HALF = g64::CT_HALF, CT_TILE = g64::CT_TILE, M_HALF = g64::M_HALF,
                     M_TILE = g64::M_TILE, OFF_C = g64::OFF_C, OFF_CT = g64::OFF_CT,
                     OFF_M = g64::OFF_M, OFF_BIAS = g64::OFF_BIAS,
                     TM_UHI = g64::TM_UHI, TM_ULO = g64::TM_ULO, TM_ST = g64::TM_ST, TM_OUT = g64::TM_OUT,
                     TM_ZHI = g64::TM_ZHI, TM_ZLO = g64::TM_ZLO;
  constexpr int CHUNK = 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + g64::OFF_BAR;
  auto BAR_C_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_C_EMPTY = [&](int s) { return bar0 + 8u * (C_STAGES + s); };
  auto BAR_B_FULL = [&](int s) { return bar0 + 8u * (2 * C_STAGES + s); };
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
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + g64::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
  const int col_tile = blockIdx.y;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NPAIR = PAIR ? 2 : 1;
  constexpr int C_ROWS = BK / NPAIR;                        // centroid rows of a C / table tile held by this CTA
  constexpr uint32_t C_ATOM_DESC = (C_ROWS * 128) >> 4;
  constexpr uint32_t M_ATOM_BYTES = C_ROWS * 128;            // one column atom of the table tile in this CTA
  constexpr uint32_t M_ATOM_DESC = M_ATOM_BYTES >> 4;
  constexpr int CT_ROWS = D / NPAIR;                        // rows (latent dims) of a C^T tile held by this CTA
  constexpr uint32_t IDESC_64 = make_idesc_f16(PAIR ? 256 : 128, BK);     // N = 64 for all three GEMMs

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(BAR_C_FULL(s), 1); mbar_init(BAR_C_EMPTY(s), 8); mbar_init(BAR_B_FULL(s), 1);
      mbar_init(BAR_CT_FULL(s), 1); mbar_init(BAR_CT_EMPTY(s), 1);
    }
    for (int s = 0; s < M_STAGES; ++s) { mbar_init(BAR_M_FULL(s), 1); mbar_init(BAR_M_EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR_ST_FULL(b), 1); mbar_init(BAR_U_FULL(b), 8 * NPAIR);
      mbar_init(BAR_CH_FULL(b), 1); mbar_init(BAR_CH_FREE(b), 8 * NPAIR);     // both exp groups fold every block
      mbar_init(BAR_G3_DONE(b), 1);
    }
    mbar_init(BAR_DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_mn_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ct_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ct_lo) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + g64::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(base + g64::OFF_TMEM_PTR), "r"(TMEM_COLS) : "memory");
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
  float zb = 0.f, s_scale = 0.f, u_unscale = 1.f;
  if (warp >= 2 && warp < 10) {
    const int64_t r = row0 + prow;
    // ---- z~ = z - shift: exponent constants for both groups, z' = 2^ez z~ (hi | lo fp16) tiles by group 0
    float nrm = 0.f, zmax = 0.f;
    if (r < n) {
      const float4* src = reinterpret_cast<const float4*>(z + r * D);
#pragma unroll 4
      for (int q = 0; q < D / 4; ++q) {
        float4 v = __ldg(src + q);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
        v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
        zmax = fmaxf(zmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
    }
    int ez = 0;
    if (zmax > 0.f && zmax < 3.0e38f) {
      const int ex = (int)((__float_as_uint(zmax) >> 23) & 0xffu) - 126;
      ez = 14 - ex;
      ez = ez > 50 ? 50 : (ez < -50 ? -50 : ez);
    }
    zb = -nrm * alpha;
    s_scale = 2.f * alpha * c_unscale * __uint_as_float((uint32_t)(127 - ez) << 23);
    if (grp == 0) {                  // the A operand of GEMM1 (split fp16) goes into TMEM
      const float zsc = __uint_as_float((uint32_t)(ez + 127) << 23);
      const float4* src = reinterpret_cast<const float4*>(z + r * D);
      uint32_t zh[32], zl[32];
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) {
          v = __ldg(src + q);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift) + q);
          v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
        }
        split_pair(v.x * zsc, v.y * zsc, zh[2 * q], zl[2 * q]);
        split_pair(v.z * zsc, v.w * zsc, zh[2 * q + 1], zl[2 * q + 1]);
      }
      TMEM_ST32(tmem_base + lane_addr + TM_ZHI, zh);
      TMEM_ST32(tmem_base + lane_addr + TM_ZLO, zl);
    }
    // ---- U' = 2^eU Ut for this column tile: group g converts packed columns [64 g, 64 g + 64) of the tile
    int e_u = 0;
    if (r < n) e_u = __ldg(eu + r);
    const float usc = __uint_as_float((uint32_t)(e_u + 127) << 23);
    u_unscale = __uint_as_float((uint32_t)(127 - e_u) << 23);
    {
      const float4* urow = reinterpret_cast<const float4*>(up + r * h64::NPAD + col_tile * NT + grp * 64);
      uint32_t h[32], l[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) v = __ldg(urow + i);
        split_pair(v.x * usc, v.y * usc, h[2 * i], l[2 * i]);
        split_pair(v.z * usc, v.w * usc, h[2 * i + 1], l[2 * i + 1]);
      }
      TMEM_ST32(tmem_base + lane_addr + TM_UHI + grp * 32, h);
      TMEM_ST32(tmem_base + lane_addr + TM_ULO + grp * 32, l);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#define MMA_TS64(d, a, b, acc) do { if (PAIR) mma_ts_f16_pair(d, a, b, IDESC_64, acc); else mma_ts_f16(d, a, b, IDESC_64, acc); } while (0)
#define COMMIT(bar) do { if (PAIR) tc_commit_pair(bar); else tc_commit(bar); } while (0)

  if (warp == 0) {
    // =========================================================== TMA producer 1: centroid tiles + bias, C^T tiles
    for (int j = 0; j < num_blocks; ++j) {
      const int cs = j % C_STAGES;
      mbar_wait(BAR_C_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_C_FULL(cs), C_TILE64);
        const uint32_t dst = base + OFF_C + cs * C_TILE64;
        if (PAIR) tma_load_3d_pair(dst, &tm_c64, BAR_C_FULL(cs), 0, j * BK + C_ROWS * (int)rank, 0);
        else tma_load_3d(dst, &tm_c64, BAR_C_FULL(cs), 0, j * BK, 0);
        mbar_expect_tx(BAR_B_FULL(cs), BIAS_BYTES);
        bulk_load_1d(base + OFF_BIAS + cs * BIAS_BYTES, cbias + (int64_t)j * BK, BIAS_BYTES, BAR_B_FULL(cs));
      }
      __syncwarp();
      mbar_wait(BAR_CT_EMPTY(cs), ((j / C_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_CT_FULL(cs), NPAIR * 2 * CT_ROWS * 128);
        const uint32_t dst = base + OFF_CT + cs * CT_TILE;
        const int row = PAIR ? CT_ROWS * (int)rank : 0;
        if (PAIR) {
          tma_load_2d_pair(dst, &tm_ct_hi, BAR_CT_FULL(cs), j * BK, row);
          tma_load_2d_pair(dst + CT_HALF, &tm_ct_lo, BAR_CT_FULL(cs), j * BK, row);
        } else {
          tma_load_2d(dst, &tm_ct_hi, BAR_CT_FULL(cs), j * BK, row);
          tma_load_2d(dst + CT_HALF, &tm_ct_lo, BAR_CT_FULL(cs), j * BK, row);
        }
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // =========================================================== TMA producer 2: natural table tiles of this column tile
    for (int j = 0; j < num_blocks; ++j) {
      const int ms = j % M_STAGES;
      mbar_wait(BAR_M_EMPTY(ms), ((j / M_STAGES) & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(BAR_M_FULL(ms), NPAIR * 2 * 2 * M_ATOM_BYTES);
        const uint32_t dst = base + OFF_M + ms * M_TILE;
        const int row = j * BK + (PAIR ? C_ROWS * (int)rank : 0);
        if (PAIR) {
          tma_load_3d_pair(dst, &tm_mn_hi, BAR_M_FULL(ms), 0, row, 2 * col_tile);
          tma_load_3d_pair(dst + M_HALF, &tm_mn_lo, BAR_M_FULL(ms), 0, row, 2 * col_tile);
        } else {
          tma_load_3d(dst, &tm_mn_hi, BAR_M_FULL(ms), 0, row, 2 * col_tile);
          tma_load_3d(dst + M_HALF, &tm_mn_lo, BAR_M_FULL(ms), 0, row, 2 * col_tile);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer 1 (pair: leader only): GEMM1 + T-GEMM
    if (leader) {
      static_assert(C_STAGES == 4 && M_STAGES == 2 && CHUNK == 2, "the 4x unrolled issue loops assume these periods");
      const uint64_t c_desc0 = make_desc_sw128(base + OFF_C);
      const uint64_t m_desc0 = make_desc_sw128(base + OFF_M);
      auto issue_st = [&](auto Jc, const int j, const uint32_t qodd /* (j / 4) & 1 */) {
        constexpr int J = decltype(Jc)::value;
        constexpr int cs = J % C_STAGES, sb = J & 1, ms = J % M_STAGES;
        mbar_wait(BAR_C_FULL(cs), qodd);
        mbar_wait(BAR_M_FULL(ms), (J / M_STAGES) & 1);
        if (j >= 2) mbar_wait(BAR_G3_DONE(sb), ((J >> 1) + 1) & 1);     // GEMM3(j-2) has consumed this buffer
        tc_fence_after();
        const uint32_t s_t = tmem_base + TM_ST + sb * 128;
        const uint32_t t_t = s_t + 64;
        const uint64_t ch = c_desc0 + ((cs * C_TILE64) >> 4);         // hi atom; lo atom C_ATOM_DESC further
        const uint64_t cl = ch + C_ATOM_DESC;
        const uint64_t mh = m_desc0 + ((ms * M_TILE) >> 4);
        const uint64_t ml = mh + (M_HALF >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_TS64(s_t, tmem_base + TM_ZHI + 8 * kk, ch + 2 * kk, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_TS64(s_t, tmem_base + TM_ZHI + 8 * kk, cl + 2 * kk, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) MMA_TS64(s_t, tmem_base + TM_ZLO + 8 * kk, ch + 2 * kk, 1);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_TS64(t_t, tmem_base + TM_UHI + 8 * kk, mh + (kk >> 2) * M_ATOM_DESC + 2 * (kk & 3), kk > 0);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_TS64(t_t, tmem_base + TM_ULO + 8 * kk, mh + (kk >> 2) * M_ATOM_DESC + 2 * (kk & 3), 1);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            MMA_TS64(t_t, tmem_base + TM_UHI + 8 * kk, ml + (kk >> 2) * M_ATOM_DESC + 2 * (kk & 3), 1);
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
      auto issue_g3 = [&](auto Jc, const int j, const uint32_t qodd) {
        constexpr int J = decltype(Jc)::value;
        constexpr int cs = J % C_STAGES, sb = J & 1;
        if (j >= 1) mbar_wait(BAR_CH_FREE(0), (J + 1) & 1);                  // both groups have folded block j - 1 out of OUT
        mbar_wait(BAR_CT_FULL(cs), qodd);
        mbar_wait(BAR_U_FULL(sb), (J >> 1) & 1);
        tc_fence_after();
        const uint32_t up = tmem_base + TM_ST + sb * 128;     // k-step kk: u'_hi at (kk>>1)*32 + (kk&1)*8, u'_lo 16 further
        const uint32_t acc = tmem_base + TM_OUT;
        const uint64_t th = ct_desc0 + ((cs * CT_TILE) >> 4);
        const uint64_t tl = th + (CT_HALF >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_TS64(acc, up + (kk >> 1) * 32 + (kk & 1) * 8, th + 2 * kk, kk > 0);       // a fresh sum per block
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_TS64(acc, up + (kk >> 1) * 32 + 16 + (kk & 1) * 8, th + 2 * kk, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            MMA_TS64(acc, up + (kk >> 1) * 32 + (kk & 1) * 8, tl + 2 * kk, 1);
          COMMIT(BAR_CT_EMPTY(cs));
          COMMIT(BAR_G3_DONE(sb));
          COMMIT(BAR_CH_FULL(0));
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
    // =========================================================== exp groups (one thread per point, 32 centroids of every block)
    float tot[32];                       // output dims [32 grp, 32 grp + 32) of sum_blocks 2^-s OUT
    float su = 0.f;                      // this group's share of sum_k u
#pragma unroll
    for (int e = 0; e < 32; ++e) tot[e] = 0.f;
    float* umax_sm = reinterpret_cast<float*>(gbase + g64::OFF_UMAX);
    // fold block jb: this group's half of the output columns, un-scaled by the block's 2^-s
    auto fold_block = [&](int jb, float unscale, bool signal) {
      mbar_wait(BAR_CH_FULL(0), jb & 1);
      tc_fence_after();
      uint32_t a[32];
      TMEM_LD32(tmem_base + lane_addr + TM_OUT + grp * 32, a);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) tot[i] = fmaf(__uint_as_float(a[i]), unscale, tot[i]);
      if (signal) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(BAR_CH_FREE(0)); else mbar_arrive(BAR_CH_FREE(0)); }
      }
    };
    float unsc_prev = 0.f;               // 2^-s of the previous block (folded one block late: its GEMM3 runs meanwhile)
    for (int j = 0; j < num_blocks; ++j) {
      const int cs = j % C_STAGES, sb = j & 1;
      const uint32_t st = tmem_base + lane_addr + TM_ST + sb * 128;
      mbar_wait(BAR_B_FULL(cs), (j / C_STAGES) & 1);
      mbar_wait(BAR_ST_FULL(sb), (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], tv[32], ph[16], pl[16];
      TMEM_LD32(st + grp * 32, sv);
      TMEM_LD32(st + 64 + grp * 32, tv);
      const float4* bias4 = reinterpret_cast<const float4*>(gbase + OFF_BIAS + cs * BIAS_BYTES) + grp * 8;
      tmem_wait_ld();
      float su_blk = 0.f, umax = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bv = bias4[q];
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * q + e;
          const float w = ex2_approx(fmaf(__uint_as_float(sv[i]), s_scale, b4[e] + zb));
          const float uv = w * __uint_as_float(tv[i]);
          su_blk += uv;
          umax = fmaxf(umax, fabsf(uv));
          tv[i] = __float_as_uint(uv);
        }
      }
      su += su_blk;
      // one scale per (point, block): the larger of the two groups' maxima, exchanged through shared memory
      umax_sm[(sb * 2 + grp) * TILE_M + prow] = umax;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      umax = fmaxf(umax, umax_sm[(sb * 2 + (grp ^ 1)) * TILE_M + prow]);
      int es = 0;
      if (umax > 0.f && umax < 3.0e38f) {
        const int ex = (int)((__float_as_uint(umax) >> 23) & 0xffu) - 126;     // umax = f 2^ex, f in [0.5, 1)
        es = 14 - ex;
        es = es > 100 ? 100 : (es < -100 ? -100 : es);
      }
      const float usc_blk = __uint_as_float((uint32_t)(es + 127) << 23);
      const float unsc_cur = __uint_as_float((uint32_t)(127 - es) << 23);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        split_pair(__uint_as_float(tv[2 * i]) * usc_blk, __uint_as_float(tv[2 * i + 1]) * usc_blk, ph[i], pl[i]);
      TMEM_ST16(st + grp * 32, ph);
      TMEM_ST16(st + grp * 32 + 16, pl);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(BAR_U_FULL(sb)); else mbar_arrive(BAR_U_FULL(sb));
        mbar_arrive(BAR_C_EMPTY(cs));
      }
      if (j > 0) fold_block(j - 1, unsc_prev, true);      // GEMM3(j) starts a fresh sum in the same accumulator
      unsc_prev = unsc_cur;
    }
    if (num_blocks > 0) fold_block(num_blocks - 1, unsc_prev, false);
    mbar_wait(BAR_DONE, 0);
    // ---------------------------------------------------------- sum_k u over both groups, write this group's half of the row
    umax_sm[grp * TILE_M + prow] = su;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      const int64_t r = row0 + prow;
      if (r < n) {
        const float sut = su + umax_sm[(grp ^ 1) * TILE_M + prow];
        const float f = u_unscale * scale;
        float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)col_tile * n + r) * D + grp * 32);
        const float4* zsrc = reinterpret_cast<const float4*>(z + r * D + grp * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 zv = __ldg(zsrc + q);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(cshift + grp * 32) + q);
          zv.x -= sh.x; zv.y -= sh.y; zv.z -= sh.z; zv.w -= sh.w;
          float4 o;
          o.x = (tot[4 * q] * c_unscale - zv.x * sut) * f;          // C^T tiles hold 2^ec (c - shift)
          o.y = (tot[4 * q + 1] * c_unscale - zv.y * sut) * f;
          o.z = (tot[4 * q + 2] * c_unscale - zv.z * sut) * f;
          o.w = (tot[4 * q + 3] * c_unscale - zv.w * sut) * f;
          dst[q] = o;
        }
      }
    }
  }
#undef MMA_TS64
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

// ------------------------------------------------------------------------------------------ tables
__host__ __device__ constexpr int sym64_index(int i, int j) {   // i <= j
  return i * 64 - (i * (i - 1)) / 2 + (j - i);
}

// [Kpad, 128] fp16: row k = [fp16(2^ec c_k) (64) | fp16(2^ec c_k - hi) (64)], plus the exp2 bias
// mean centroid [64] (the expanded distance form is evaluated about it, as for d = 16)
__global__ void centroid_mean64_kernel(const float* __restrict__ c, int K, float* __restrict__ shift) {
  __shared__ double part[4][64];
  const int j = threadIdx.x & 63, slice = threadIdx.x >> 6;       // 256 threads: 4 slices of the K range
  double acc = 0.0;
  for (int k = slice; k < K; k += 4) acc += (double)c[(int64_t)k * 64 + j];
  part[slice][j] = acc;
  __syncthreads();
  if (slice == 0) {
    const float m = (float)((part[0][j] + part[1][j] + part[2][j] + part[3][j]) / (double)K);
    shift[j] = isfinite(m) ? m : 0.f;
  }
}

// stats[0] += sum_k ||c_k - shift||^2, stats[1] = max |c - shift|
__global__ void centred_stats64_kernel(const float* __restrict__ c, const float* __restrict__ shift, int K,
                                       float* __restrict__ stats) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float nrm = 0.f, amax = 0.f;
  for (int j = 0; j < 64; ++j) {
    const float v = c[(int64_t)k * 64 + j] - shift[j];
    nrm = fmaf(v, v, nrm);
    amax = fmaxf(amax, fabsf(v));
  }
  if (isfinite(nrm)) atomicAdd(&stats[0], nrm);
  if (isfinite(amax)) atomicMax(reinterpret_cast<int*>(&stats[1]), __float_as_int(amax));
}

__global__ void pack_c64_kernel(const float* __restrict__ c, const float* __restrict__ shift, int K, int Kpad,
                                float scale, float inv_T2_log2e, __half* __restrict__ c64, float* __restrict__ cbias) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  float nrm = 0.f;
  for (int j = 0; j < 64; ++j) {
    const float v = (k < K) ? c[(int64_t)k * 64 + j] - shift[j] : 0.f;
    nrm = fmaf(v, v, nrm);
    const float sv = v * scale;
    const __half h = __float2half_rn(sv);
    c64[(int64_t)k * 128 + j] = h;
    c64[(int64_t)k * 128 + 64 + j] = __float2half_rn(sv - __half2float(h));
  }
  cbias[k] = (k < K) ? -nrm * inv_T2_log2e : -1.0e30f;
}

// packed-transposed fp16 tables [2176, Kpad]: hi = fp16(scale * M), lo = fp16(scale * M - hi)
__global__ void pack_sym64_h_kernel(const float* __restrict__ M, int Kpad, float scale, __half* __restrict__ hi_t,
                                    __half* __restrict__ lo_t) {
  const int64_t total = (int64_t)tc::h64::NPAD * Kpad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx / Kpad), k = (int)(idx - (int64_t)p * Kpad);
    float v = 0.f;
    if (p < tc::h64::NPACK) {
      int i = 0, base = 0;
      while (p >= base + (64 - i)) { base += 64 - i; ++i; }
      const int j = i + (p - base);
      v = scale * 0.5f * (M[(int64_t)k * 4096 + i * 64 + j] + M[(int64_t)k * 4096 + j * 64 + i]);
    }
    const __half h = __float2half_rn(v);
    hi_t[idx] = h;
    lo_t[idx] = __float2half_rn(v - __half2float(h));
  }
}

// natural fp16 tables [Kpad, 2176] (row = centroid, 2080 packed columns + zeros): B operand of the T GEMM
__global__ void pack_sym64_nat_h_kernel(const float* __restrict__ M, int Kpad, float scale, __half* __restrict__ hi_n,
                                        __half* __restrict__ lo_n) {
  const int64_t total = (int64_t)Kpad * tc::h64::NPAD;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / tc::h64::NPAD), p = (int)(idx - (int64_t)k * tc::h64::NPAD);
    float v = 0.f;
    if (p < tc::h64::NPACK) {
      int i = 0, base = 0;
      while (p >= base + (64 - i)) { base += 64 - i; ++i; }
      const int j = i + (p - base);
      v = scale * 0.5f * (M[(int64_t)k * 4096 + i * 64 + j] + M[(int64_t)k * 4096 + j * 64 + i]);
    }
    const __half h = __float2half_rn(v);
    hi_n[idx] = h;
    lo_n[idx] = __float2half_rn(v - __half2float(h));
  }
}

// (c - shift)^T scaled by 2^ec, split fp16 [64, Kpad] (centroid index contiguous): B operand of GEMM3
__global__ void pack_ct64_kernel(const float* __restrict__ c, const float* __restrict__ shift, int K, int Kpad,
                                 float scale, __half* __restrict__ ct_hi, __half* __restrict__ ct_lo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  for (int j = 0; j < 64; ++j) {
    const float v = (k < K) ? (c[(int64_t)k * 64 + j] - shift[j]) * scale : 0.f;
    const __half h = __float2half_rn(v);
    ct_hi[(int64_t)j * Kpad + k] = h;
    ct_lo[(int64_t)j * Kpad + k] = __float2half_rn(v - __half2float(h));
  }
}

// packed [N, 2176] -> full symmetric [N, 64, 64], + lambda on the diagonal.  One CTA per point: the packed row goes
// through shared memory (coalesced 128-bit loads), the 64 x 64 matrix leaves as coalesced 128-bit stores -- a thread per
// output element gathering from global memory was long-scoreboard bound (1.6 ms per 2^17 points, 2.0 TB/s).
__global__ void __launch_bounds__(256)
unpack_sym64_kernel(const float* __restrict__ packed, int64_t n, float lambda, float* __restrict__ full) {
  __shared__ __align__(16) float sm[tc::h64::NPAD];
  const int tid = threadIdx.x;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(packed + p * tc::h64::NPAD);
    for (int c4 = tid; c4 < tc::h64::NPAD / 4; c4 += 256) reinterpret_cast<float4*>(sm)[c4] = __ldg(src + c4);
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(full + p * 4096);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e4 = tid + 256 * q;                  // float4 index: row e4 / 16, columns 4 (e4 % 16) ..
      const int i = e4 >> 4, j0 = (e4 & 15) << 2;
      float v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = j0 + t;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        v[t] = sm[(lo * (129 - lo)) / 2 + (hi - lo)] + (i == j ? lambda : 0.f);
      }
      dst[e4] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
  }
}

typedef CUresult (*PFN_encodeTiled64)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

static int encode(PFN_encodeTiled64 enc, CUtensorMap* map, void* ptr, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box) {
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (d = 64 tables) failed with CUresult " + std::to_string((int)r));
    return 4;
  }
  return 0;
}

// Build the d = 64 split-fp16 tables and descriptors (symmetric tables only).  Synchronises `s`.
int tc_build_h64_tables(rlvae_tables* t, cudaStream_t s) {
  const int Kpad = t->Kpad, K = t->K;
  // centre of the expanded distance form (mean centroid) + statistics of the centred table
  RLVAE_CUDA_OK(cudaMalloc(&t->cshift, sizeof(float) * 66));
  RLVAE_CUDA_OK(cudaMemsetAsync(t->cshift, 0, sizeof(float) * 66, s));
  {
    const char* ce = getenv("RLVAE_TC_CENTRE");
    if (ce == nullptr || ce[0] != '0') {
      centroid_mean64_kernel<<<1, 256, 0, s>>>(t->c, K, t->cshift);
      RLVAE_LAUNCH_OK();
    }
  }
  centred_stats64_kernel<<<(K + 127) / 128, 128, 0, s>>>(t->c, t->cshift, K, t->cshift + 64);
  RLVAE_LAUNCH_OK();
  float h_st[2] = {0.f, 0.f};
  RLVAE_CUDA_OK(cudaMemcpyAsync(h_st, t->cshift + 64, sizeof(h_st), cudaMemcpyDeviceToHost, s));
  RLVAE_CUDA_OK(cudaStreamSynchronize(s));
  const float cmax = h_st[1];
  t->r2mean_centred = h_st[0] / (float)K;
  if (!(t->m_absmax > 0.f) || !(cmax > 0.f)) return 0;      // degenerate tables: direct path only
  int ex = 0;
  frexpf(t->m_absmax, &ex);
  const int em = 14 - ex;
  frexpf(cmax, &ex);
  const int ec = 14 - ex;
  if (em <= -60 || em >= 60 || ec <= -60 || ec >= 60) return 0;
  RLVAE_CUDA_OK(cudaMalloc(&t->c64h, sizeof(__half) * (size_t)Kpad * 128));
  RLVAE_CUDA_OK(cudaMalloc(&t->cbias, sizeof(float) * (size_t)Kpad));
  RLVAE_CUDA_OK(cudaMalloc(&t->Mh_hi, sizeof(__half) * (size_t)tc::h64::NPAD * Kpad));
  RLVAE_CUDA_OK(cudaMalloc(&t->Mh_lo, sizeof(__half) * (size_t)tc::h64::NPAD * Kpad));
  pack_c64_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(t->c, t->cshift, K, Kpad, ldexpf(1.f, ec), 1.4426950408889634f / t->T2,
                                                     static_cast<__half*>(t->c64h), t->cbias);
  RLVAE_LAUNCH_OK();
  pack_sym64_h_kernel<<<1184, 256, 0, s>>>(t->M, Kpad, ldexpf(1.f, em), static_cast<__half*>(t->Mh_hi),
                                           static_cast<__half*>(t->Mh_lo));
  RLVAE_LAUNCH_OK();
  RLVAE_CUDA_OK(cudaStreamSynchronize(s));
  t->h16_out_scale = ldexpf(1.f, -(14 + em));
  t->h16_m_unscale = ldexpf(1.f, -em);
  t->c64_unscale = ldexpf(1.f, -ec);
  // gradient kernel tables: natural split-fp16 M [Kpad, 2176] and (c - shift)^T [64, Kpad]
  const bool grad_tables =
      cudaMalloc(&t->Mnh_hi, sizeof(__half) * (size_t)tc::h64::NPAD * Kpad) == cudaSuccess &&
      cudaMalloc(&t->Mnh_lo, sizeof(__half) * (size_t)tc::h64::NPAD * Kpad) == cudaSuccess &&
      cudaMalloc(&t->ct64_hi, sizeof(__half) * (size_t)64 * Kpad) == cudaSuccess &&
      cudaMalloc(&t->ct64_lo, sizeof(__half) * (size_t)64 * Kpad) == cudaSuccess;
  if (grad_tables) {
    pack_sym64_nat_h_kernel<<<1184, 256, 0, s>>>(t->M, Kpad, ldexpf(1.f, em), static_cast<__half*>(t->Mnh_hi),
                                                 static_cast<__half*>(t->Mnh_lo));
    RLVAE_LAUNCH_OK();
    pack_ct64_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(t->c, t->cshift, K, Kpad, ldexpf(1.f, ec),
                                                        static_cast<__half*>(t->ct64_hi), static_cast<__half*>(t->ct64_lo));
    RLVAE_LAUNCH_OK();
    RLVAE_CUDA_OK(cudaStreamSynchronize(s));
  } else {
    (void)cudaGetLastError();      // out of memory for the extra tables: the gradient stays on the direct kernel
    if (t->Mnh_hi) cudaFree(t->Mnh_hi);
    if (t->Mnh_lo) cudaFree(t->Mnh_lo);
    if (t->ct64_hi) cudaFree(t->ct64_hi);
    if (t->ct64_lo) cudaFree(t->ct64_lo);
    t->Mnh_hi = t->Mnh_lo = t->ct64_hi = t->ct64_lo = nullptr;
  }

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  RLVAE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  RLVAE_REQUIRE(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  PFN_encodeTiled64 enc = reinterpret_cast<PFN_encodeTiled64>(fn);
  {   // centroid rows [Kpad, 128] fp16 viewed as [2 atoms][Kpad][64]
    cuuint64_t dims[3] = {64, (cuuint64_t)Kpad, 2};
    cuuint64_t strides[2] = {128 * sizeof(__half), 64 * sizeof(__half)};
    cuuint32_t box[3] = {64, tc::BK, 2};
    if (int rc = encode(enc, &t->tm_c64, t->c64h, 3, dims, strides, box)) return rc;
    cuuint32_t box2[3] = {64, tc::BK / 2, 2};
    if (int rc = encode(enc, &t->tm_c64_2, t->c64h, 3, dims, strides, box2)) return rc;
  }
  {   // table tiles [2176, Kpad] fp16: box = 64 centroids x 128 (pair: 64) rows
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)tc::h64::NPAD};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * sizeof(__half)};
    cuuint32_t box[2] = {tc::BK, tc::h64::NT};
    if (int rc = encode(enc, &t->tm_mh_hi, t->Mh_hi, 2, dims, strides, box)) return rc;
    if (int rc = encode(enc, &t->tm_mh_lo, t->Mh_lo, 2, dims, strides, box)) return rc;
    cuuint32_t box2[2] = {tc::BK, tc::h64::NT / 2};
    if (int rc = encode(enc, &t->tm_mh2_hi, t->Mh_hi, 2, dims, strides, box2)) return rc;
    if (int rc = encode(enc, &t->tm_mh2_lo, t->Mh_lo, 2, dims, strides, box2)) return rc;
  }
  if (t->Mnh_hi != nullptr) {
    {   // natural tables [Kpad, 2176] fp16 viewed as [34 column atoms][Kpad][64]: box = 2 atoms x 64 (pair: 32) rows
      cuuint64_t dims[3] = {64, (cuuint64_t)Kpad, (cuuint64_t)(tc::h64::NPAD / 64)};
      cuuint64_t strides[2] = {(cuuint64_t)tc::h64::NPAD * sizeof(__half), 64 * sizeof(__half)};
      cuuint32_t box[3] = {64, tc::BK, 2};
      if (int rc = encode(enc, &t->tm_mnh_hi, t->Mnh_hi, 3, dims, strides, box)) return rc;
      if (int rc = encode(enc, &t->tm_mnh_lo, t->Mnh_lo, 3, dims, strides, box)) return rc;
      cuuint32_t box2[3] = {64, tc::BK / 2, 2};
      if (int rc = encode(enc, &t->tm_mnh2_hi, t->Mnh_hi, 3, dims, strides, box2)) return rc;
      if (int rc = encode(enc, &t->tm_mnh2_lo, t->Mnh_lo, 3, dims, strides, box2)) return rc;
    }
    {   // (c - shift)^T [64, Kpad] fp16: box = 64 centroids x 64 (pair: 32) rows
      cuuint64_t dims[2] = {(cuuint64_t)Kpad, 64};
      cuuint64_t strides[1] = {(cuuint64_t)Kpad * sizeof(__half)};
      cuuint32_t box[2] = {tc::BK, 64};
      if (int rc = encode(enc, &t->tm_ct64_hi, t->ct64_hi, 2, dims, strides, box)) return rc;
      if (int rc = encode(enc, &t->tm_ct64_lo, t->ct64_lo, 2, dims, strides, box)) return rc;
      cuuint32_t box2[2] = {tc::BK, 32};
      if (int rc = encode(enc, &t->tm_ct64_2_hi, t->ct64_hi, 2, dims, strides, box2)) return rc;
      if (int rc = encode(enc, &t->tm_ct64_2_lo, t->ct64_lo, 2, dims, strides, box2)) return rc;
    }
  }
  return 0;
}

static bool h64_use_pairs() {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_PAIR");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();   // initialised once, thread-safe (C++11 magic static)
  return v == 1;
}

template <bool PAIR>
static int launch_h64(const rlvae_tables* t, const float* z, int64_t n, float* packed, cudaStream_t s) {
  auto kern = tc::inverse_metric_h64_kernel<PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::h64::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, tc::h64::NPAD / tc::h64::NT, 1);
  cfg.blockDim = dim3(tc::h64::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::h64::SMEM_BYTES;
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
  const float cu = t->c64_unscale, os = t->h16_out_scale;
  if (PAIR) {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c64_2, t->tm_mh2_hi, t->tm_mh2_lo, z, cbias, n, nb, alpha,
                                       cu, os, t->cshift, packed));
  } else {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c64, t->tm_mh_hi, t->tm_mh_lo, z, cbias, n, nb, alpha,
                                       cu, os, t->cshift, packed));
  }
  return 0;
}

// d = 64, symmetric tables: z -> full G^{-1} [N,64,64] through the packed [N,2176] scratch
int launch_inverse_metric_h64(const rlvae_tables* t, const float* z, int64_t n, float* ginv, float* packed_scratch,
                              cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->d == 64 && t->symmetric && t->c64h != nullptr, "d = 64 tensor path needs symmetric tables");
  RLVAE_REQUIRE(packed_scratch != nullptr, "d = 64 tensor path needs the packed scratch buffer");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed_scratch) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(ginv) & 15) == 0,
                "tensor path needs 16-byte aligned z, G^-1 buffer and scratch");
  if (int rc = h64_use_pairs() ? launch_h64<true>(t, z, n, packed_scratch, s)
                               : launch_h64<false>(t, z, n, packed_scratch, s))
    return rc;
  unpack_sym64_kernel<<<(unsigned)(n < 148 * 64 ? n : 148 * 64), 256, 0, s>>>(packed_scratch, n, t->lambda, ginv);
  RLVAE_LAUNCH_OK();
  return 0;
}

template <bool PAIR>
static int launch_g64(const rlvae_tables* t, const float* z, const float* u, const int* eu, int64_t n, float scale,
                      float* partial, cudaStream_t s) {
  auto kern = tc::metric_grad_h64_kernel<PAIR>;
  RLVAE_OPT_IN_SMEM(kern, (int)tc::g64::SMEM_BYTES);
  unsigned tiles = (unsigned)((n + tc::TILE_M - 1) / tc::TILE_M);
  if (PAIR) tiles = (tiles + 1) & ~1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles, tc::g64::NTILES, 1);
  cfg.blockDim = dim3(tc::g64::THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc::g64::SMEM_BYTES;
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
  const float sc = scale * t->h16_m_unscale;
  const float cu = t->c64_unscale;
  if (PAIR) {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c64_2, t->tm_mnh2_hi, t->tm_mnh2_lo, t->tm_ct64_2_hi,
                                       t->tm_ct64_2_lo, z, u, eu, cbias, n, nb, alpha, sc, cu, t->cshift, partial));
  } else {
    RLVAE_LAUNCH_EX(cudaLaunchKernelEx(&cfg, kern, t->tm_c64, t->tm_mnh_hi, t->tm_mnh_lo, t->tm_ct64_hi,
                                       t->tm_ct64_lo, z, u, eu, cbias, n, nb, alpha, sc, cu, t->cshift, partial));
  }
  return 0;
}

int64_t metric_grad_h64_scratch_floats(int64_t n) {
  return n * (int64_t)(tc::g64::NTILES * 64) + ((n + 3) & ~(int64_t)3) +      // partial tiles + per-point exponents
         n * (int64_t)tc::h64::NPAD;                                           // + packed Ut
}

bool metric_grad_h64_available(const rlvae_tables* t) {
  return t->d == 64 && t->symmetric && t->c64h != nullptr && t->Mnh_hi != nullptr;
}

// d = 64, symmetric tables: (scale) * sum_k w_k <U, M_k> (c_k - z), u = [N,64,64] (any U);
// scratch: metric_grad_h64_scratch_floats(n) floats (16-byte aligned)
int launch_metric_grad_h64(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale, float* out,
                           float* scratch, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(metric_grad_h64_available(t), "d = 64 tensor gradient needs symmetric tables");
  RLVAE_REQUIRE(scratch != nullptr, "d = 64 tensor gradient needs a workspace");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
                "tensor path needs 16-byte aligned z, u, out and workspace");
  float* partial = scratch;
  int* eu = reinterpret_cast<int*>(scratch + n * (int64_t)(tc::g64::NTILES * 64));
  float* up = scratch + n * (int64_t)(tc::g64::NTILES * 64) + ((n + 3) & ~(int64_t)3);
  const unsigned g1 = (unsigned)(n < 148 * 64 ? n : 148 * 64);
  tc::pack_u64_kernel<<<g1, 256, 0, s>>>(u, n, up, eu);
  RLVAE_LAUNCH_OK();
  if (int rc = h64_use_pairs() ? launch_g64<true>(t, z, up, eu, n, scale, partial, s)
                               : launch_g64<false>(t, z, up, eu, n, scale, partial, s))
    return rc;
  const int64_t total4 = n * 16;
  tc::reduce_partials64_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(partial, n, tc::g64::NTILES, out);
  RLVAE_LAUNCH_OK();
  return 0;
}

}  // namespace rlvae
