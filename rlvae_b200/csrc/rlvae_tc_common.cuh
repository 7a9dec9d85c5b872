// Shared constants and PTX wrappers of the tcgen05 / TMEM / TMA kernels (rlvae_tc.cu, rlvae_tc16.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>

#include "rlvae_internal.h"

namespace rlvae {
namespace tc {

constexpr int TILE_M = 128;          // points per CTA
constexpr int BK = 64;               // centroids per super-block (two 32-wide swizzle atoms)
constexpr int NCOL = 256;            // d*d
constexpr int NHALF = 128;           // output columns per CTA
constexpr int CHUNK_BLOCKS = 2;      // super-blocks accumulated on the tensor core before an fp32 fold
constexpr int C_STAGES = 3;
constexpr int SP_BUFS = 2;           // S/P TMEM buffers
constexpr int M_STAGES = 5;
constexpr int THREADS = 320;         // TMA warp, MMA warp, two exp warpgroups of 4 warps

constexpr uint32_t A_BYTES = TILE_M * 128;            // one 128 x 32 fp32 operand tile
constexpr uint32_t C_TILE_BYTES = BK * 128;           // 64 centroid rows of [hi|lo]
constexpr uint32_t M_ATOM_BYTES = NHALF * 128;        // 128 rows x 32 centroids fp32
constexpr uint32_t M_TILE_BYTES = 2 * M_ATOM_BYTES;   // one stage = both atoms of a super-block
constexpr uint32_t BIAS_BYTES = BK * 4;

// shared memory map (offsets from a 1024-aligned base)
constexpr uint32_t OFF_A1 = 0;                                    // [z_hi | z_hi]
constexpr uint32_t OFF_A2 = OFF_A1 + A_BYTES;                     // [z_lo | 0   ]
constexpr uint32_t OFF_C = OFF_A2 + A_BYTES;                      // C ring
constexpr uint32_t OFF_M = OFF_C + C_STAGES * C_TILE_BYTES;       // M ring
constexpr uint32_t OFF_BIAS = OFF_M + M_STAGES * M_TILE_BYTES;    // bias ring
constexpr uint32_t OFF_BAR = OFF_BIAS + C_STAGES * BIAS_BYTES;    // mbarriers
constexpr int NUM_BARS = 3 * C_STAGES + 2 * M_STAGES + 2 * SP_BUFS + 2 + 1 + 2;
constexpr uint32_t OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;         // + alignment slack
constexpr int OUT_LD = 132;                                       // epilogue staging row (floats)
static_assert(TILE_M * OUT_LD * 4 <= M_STAGES * M_TILE_BYTES, "epilogue staging must fit the M ring");
static_assert(144 == kSymCols, "packed row length");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_CH = 0;        // + buf*128 : chunk accumulator (2 buffers)
constexpr uint32_t TM_SP = 256;      // + buf*128 : S/P_hi (64) ; + 64 : P_lo (64)

// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a=b=tf32, K-major both, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t IDESC_G1 = make_idesc(128, BK);      // N = 64
constexpr uint32_t IDESC_G2 = make_idesc(128, NHALF);
// Symmetric tables: only the 136 entries i <= j of every M_k (packed row-major upper triangle,
// padded to 144) are accumulated.  Column half 0 owns packed columns [0,80), half 1 [80,144).
constexpr int SYM_COLS = 144;        // packed row length in HBM (136 real + 8 zero)
constexpr int SYM_H0 = 80;           // columns of half 0 (MMA N = 80), half 1 has 64 (N = 64)
constexpr uint32_t SYM_ATOM_BYTES = SYM_H0 * 128;
constexpr int SYM_FOLD = 48;         // running-total columns per exp thread (2 groups x 48 >= 80)
constexpr int SYM_OUT_LD = 100;
__host__ __device__ constexpr int sym_index(int i, int j) {   // i <= j
  return i * 16 - (i * (i - 1)) / 2 + (j - i);
}

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
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
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  The pair = two CTAs of a cluster on one TPC: the MMA
// spans M = 256 (128 TMEM lanes in each CTA) and each CTA supplies HALF of the B tile from its own
// shared memory, so per-SM TMA ingest and B-operand smem reads are halved.  The leader (even rank)
// issues; barriers the leader waits on are signalled from the peer through shared::cluster.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // works from either CTA of the pair
  // default semantics (.release.cta): what is published lives in this SM's TMEM and is ordered by
  // tcgen05.wait/fence; a cluster-scope release here cost ~900 cycles per arrive (measured)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                                 int c1) {           // completes on the LEADER's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                                 int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {       // arrives on both CTAs' barrier
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle, 128-byte rows, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
#define TMEM_LD32(taddr, r)                                                                          \
  asm volatile(                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"  \
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),          \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),      \
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),   \
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),   \
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                           \
      : "r"(taddr) : "memory")
#define TMEM_LD16(taddr, r)                                                                          \
  asm volatile(                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                      \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                              \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),          \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),      \
        "=r"(r[14]), "=r"(r[15])                                                                     \
      : "r"(taddr) : "memory")
#define TMEM_ST8(taddr, r)                                                                           \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"               \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), \
                 "r"(r[7]) : "memory")
#define TMEM_ST32(taddr, r)                                                                          \
  asm volatile(                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25," \
      "%26,%27,%28,%29,%30,%31,%32};"                                                                \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),     \
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), \
        "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),          \
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),          \
        "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// One lane of a converged warp.  Issuing tcgen05.mma / TMA under `if (lane == 0)` makes ptxas wrap
// every UTCHMMA in a per-lane ELECT loop (~45 cycles per MMA, measured); under elect.sync the
// issue cost drops to the hardware floor (N/2 cycles for M=128 kind::tf32).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------ shared pieces
// Z operand tiles for GEMM1 (written by one thread per point), returns zb = -alpha*||z||^2.
__device__ __forceinline__ float write_z_tiles(uint8_t* gbase, const float* __restrict__ z, int64_t r,
                                               int64_t n, int prow, float alpha) {
  float zv[16];
  if (r < n) {
    const float4* src = reinterpret_cast<const float4*>(z + r * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = __ldg(src + q);
      zv[4 * q] = v.x; zv[4 * q + 1] = v.y; zv[4 * q + 2] = v.z; zv[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) zv[j] = 0.f;
  }
  float nrm = 0.f, hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    nrm = fmaf(zv[j], zv[j], nrm);
    hi[j] = tf32_rna(zv[j]);
    lo[j] = zv[j] - hi[j];
  }
  // 128-byte swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
  uint8_t* a1 = gbase + OFF_A1 + prow * 128;
  uint8_t* a2 = gbase + OFF_A2 + prow * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int q = c & 3;  // which 4 of the 16 dims
    const float4 vh = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    const float4 vl = (c < 4) ? make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3])
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    const int pc = (c ^ (prow & 7)) * 16;
    *reinterpret_cast<float4*>(a1 + pc) = vh;   // [z_hi | z_hi]
    *reinterpret_cast<float4*>(a2 + pc) = vl;   // [z_lo | 0]
  }
  return -nrm * alpha;
}

// GEMM1: S[128 x 32] = (z_hi|z_hi).(c_hi|c_lo) + z_lo.c_hi   (6 tcgen05.mma, K = 8 each)
template <bool PAIR = false>
__device__ __forceinline__ void issue_gemm1(uint32_t d_tmem, uint64_t a1_desc, uint64_t a2_desc,
                                            uint64_t b_desc) {
  constexpr uint32_t id = make_idesc(PAIR ? 256 : 128, BK);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (PAIR) mma_ss_pair(d_tmem, a1_desc + 2 * k, b_desc + 2 * k, id, k > 0);
    else mma_ss(d_tmem, a1_desc + 2 * k, b_desc + 2 * k, id, k > 0);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (PAIR) mma_ss_pair(d_tmem, a2_desc + 2 * k, b_desc + 2 * k, id, 1);
    else mma_ss(d_tmem, a2_desc + 2 * k, b_desc + 2 * k, id, 1);
  }
}

// instruction descriptor for kind::f16: c = f32, a = b = f16, K-major both
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts_f16_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::f16
__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ss_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
#define TMEM_ST16(taddr, r)                                                                          \
  asm volatile(                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                    \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),     \
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), \
        "r"(r[15]) : "memory")

// (a, b) -> packed fp16 pair {hi half = a, lo half = b} and the fp32 residuals of both
__device__ __forceinline__ void split_pair(float even, float odd, uint32_t& hi2, uint32_t& lo2) {
  uint32_t h;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(odd), "f"(even));   // .hi = odd, .lo = even element
  float he, ho;
  asm("{\n\t.reg .f16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}"
      : "=f"(he), "=f"(ho) : "r"(h));
  uint32_t l;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(odd - ho), "f"(even - he));
  hi2 = h;
  lo2 = l;
}

// packed fp32x2 arithmetic (two lanes per 64-bit register pair)
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ void ffma2_acc(float2& acc, float2 a, float2 b) {   // acc += a * b
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
}
// ||z - c||^2 for one centroid row (16 fp32 in shared memory, the same address for every lane:
// broadcast loads) and the negated point held as 8 float2: exact differences, like the reference
__device__ __forceinline__ float dist2_row16(const float4* __restrict__ crow, const float2 (&nz)[8]) {
  // 16 packed subtractions + 16 packed FMAs = 32 FMA-pipe cycles per (point, centroid): the exact mode is
  // bound by the FP32 pipe (floor 9.3 ms per 2^20 x 10k), not by latency -- more accumulators only add
  // register pressure (measured: 24.5 ms with four of them vs 18.5 ms with one)
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 c = crow[q];
    const float2 d0 = fadd2(make_float2(c.x, c.y), nz[2 * q]);
    const float2 d1 = fadd2(make_float2(c.z, c.w), nz[2 * q + 1]);
    ffma2_acc(acc, d0, d0);
    ffma2_acc(acc, d1, d1);
  }
  return acc.x + acc.y;
}

// the same from global memory (read-only path)
__device__ __forceinline__ float dist2_row16_ldg(const float4* __restrict__ crow, const float2 (&nz)[8]) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 c = __ldg(crow + q);
    const float2 d0 = fadd2(make_float2(c.x, c.y), nz[2 * q]);
    const float2 d1 = fadd2(make_float2(c.z, c.w), nz[2 * q + 1]);
    ffma2_acc(acc, d0, d0);
    ffma2_acc(acc, d1, d1);
  }
  return acc.x + acc.y;
}

// Per-lane scratch row of 32 words in shared memory for the HYBRID refinement: word i of lane l is component
// i % 4 of the float4 at [plane i / 4][lane l] -- the 128-bit spills / reloads of a warp touch 512 contiguous
// bytes per plane (conflict free), and a lane can address ITS word i with a run-time i, which a register array
// cannot (that costs a 32-way select chain per access).  HYB_ROW_BYTES per warp.
constexpr uint32_t HYB_ROW_BYTES = 8 * 32 * 16;
__device__ __forceinline__ void hyb_row_spill(float* row, const uint32_t (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    reinterpret_cast<uint4*>(row)[q * 32] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void hyb_row_reload(const float* row, uint32_t (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 t = reinterpret_cast<const uint4*>(row)[q * 32];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ int hyb_row_word(int i) { return (i >> 2) * 128 + (i & 3); }

// HYBRID weight mode: ex[i] (i < 32, float bits) are this thread's exponents for 32 consecutive
// centroids from the expanded form; the ones flagged in `live` are replaced by
// neg_alpha * ||z - c_i||^2 + shift from exact differences (crows: the 32 natural centroid rows; SMEM_ROWS: staged in
// shared memory by the producer warp -- the forward kernel, where per-lane rows from L2 made the exp warps latency
// bound -- otherwise in global memory, read-only path).
// Must be called by the whole warp.
// Sparse case (the usual one at small T): every lane works on its own flagged centroid at the same
// time, so a round costs one distance however many lanes need one; the exponents sit in the lane's
// shared-memory row (`row` = this lane's float4 of plane 0) while the rounds run, so a refined value is ONE store.
// Dense case: a uniform sweep with broadcast loads, like the exact mode.
template <bool SMEM_ROWS>
__device__ __forceinline__ void refine_exponents(uint32_t (&ex)[32], uint32_t live, const float4* __restrict__ crows,
                                                 const float2 (&nz)[8], float neg_alpha, float shift, float* row) {
  if (!__any_sync(0xffffffffu, live != 0u)) return;
  const int total = __reduce_add_sync(0xffffffffu, __popc(live));
  if (total > 160) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float v = fmaf((SMEM_ROWS ? dist2_row16(crows + i * 4, nz) : dist2_row16_ldg(crows + i * 4, nz)), neg_alpha, shift);
      if ((live >> i) & 1u) ex[i] = __float_as_uint(v);
    }
    return;
  }
  hyb_row_spill(row, ex);
  while (__any_sync(0xffffffffu, live != 0u)) {
    if (live != 0u) {
      const int b = __ffs(live) - 1;
      live &= live - 1u;
      row[hyb_row_word(b)] = fmaf((SMEM_ROWS ? dist2_row16(crows + b * 4, nz) : dist2_row16_ldg(crows + b * 4, nz)), neg_alpha, shift);
    }
  }
  hyb_row_reload(row, ex);
}

template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

}  // namespace tc
}  // namespace rlvae
