// Per-point d x d linear algebra ("K3"): batched inverse / log|det| / sign / diag(inverse)
// and cholesky(A + jitter I) @ eps.   HBM-bound: 4*d*d bytes in (+ the same out when the
// inverse is stored) per point; the arithmetic stays in registers with warp shuffles.
//
//   batched_inverse : replaces torch.linalg.inv / slogdet / det at
//                     ref src/models/components/metric_tensor.py:152,175 and
//                     ref src/models/samplers/hmc_sampler.py:28
//   chol_apply      : replaces torch.linalg.cholesky + einsum('bij,bj->bi') at
//                     ref src/models/samplers/riemannian_sampler.py:83-84,156-157,200-201,271-272
//
// Layout: one matrix per LANES = min(d, 32) lanes of a warp, lane l owns rows l, l+32 (d = 64).
// Global <-> register traffic is staged through shared memory so that every global access is a
// contiguous, coalesced run of the 4*d*d-byte matrix.
#include <cmath>

#include "rlvae_internal.h"

namespace rlvae {

template <int D>
struct PP {
  static constexpr int LANES = D < 32 ? D : 32;
  static constexpr int RPL = D / LANES;                 // rows per lane
  static constexpr int THREADS = D == 64 ? 64 : 128;
  static constexpr int MATS = THREADS / LANES;          // matrices per CTA
  static constexpr int LD = D + 1;                      // padded smem row
  static_assert(D % LANES == 0, "latent_dim must be a power of two <= 64");
};

// In-place Gauss-Jordan with implicit partial (row) pivoting.
// After the sweep the lane that served as pivot at step j holds row j of A^{-1}; its slot m
// holds column perm[m] (perm[m] = pivot row chosen at step m).
__host__ __device__ constexpr int sym16_index(int i, int j) {   // packed upper triangle, i <= j
  return i * 16 - (i * (i - 1)) / 2 + (j - i);
}

// PACKED (D == 16 only): the input is the symmetric packed layout [N, 144] of the tensor kernel
// `list` / `count` (optional): process only the matrices list[0 .. *count) -- the fallback pass behind
// sym16_cholesky_kernel -- with a grid-stride loop; otherwise matrix = blockIdx.x * MATS + group.
// `packed_out` (PACKED only): the inverse in the packed symmetric layout (upper triangle).
template <int D, bool PACKED = false>
__global__ void __launch_bounds__(PP<D>::THREADS)
batched_inverse_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ inv,
                       float* __restrict__ logabsdet, float* __restrict__ sign,
                       float* __restrict__ diag_inv, int transpose_inv,
                       const int* __restrict__ list = nullptr, const int* __restrict__ count = nullptr,
                       float* __restrict__ packed_out = nullptr, float lad_scale = 1.f, int dr = D) {
  // dr <= D: the real matrices are dr x dr (any latent_dim, e.g. 10, 12, 20); they are embedded as
  // diag(A, I_{D-dr}), which leaves the inverse block, log|det| and the sign unchanged (the identity
  // rows are only ever chosen as pivots for their own columns, with pivot 1).
  using P = PP<D>;
  constexpr int LANES = P::LANES, RPL = P::RPL, LD = P::LD;
  __shared__ float stage[P::MATS * D * LD];

  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t limit = (list != nullptr) ? (int64_t)*count : n;
  for (int64_t slot = (int64_t)blockIdx.x * P::MATS + grp; slot < limit; slot += (int64_t)gridDim.x * P::MATS) {
  const int64_t mat = (list != nullptr) ? (int64_t)list[slot] : slot;
  const bool live = true;
  const int64_t msafe = mat;
  float* sm = stage + grp * D * LD;
  // lanes of one matrix always sit inside one warp
  const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                       : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));

  if (PACKED) {
    const float* src = a + msafe * kSymCols;
    for (int i = lane; i < D * D; i += LANES) {
      const int rr = i / D, cc = i % D;
      sm[rr * LD + cc] = src[rr <= cc ? sym16_index(rr, cc) : sym16_index(cc, rr)];
    }
  } else {
    const float* src = a + msafe * dr * dr;
    if (D >= 8 && dr == D && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // 128-bit loads, several in flight: a scalar loop waits out one global-memory latency per element
      // (64 x 64: 128 dependent round trips per lane -- part of the d = 64 kernel time: 20 -> 13-18 ms per 2^17 matrices)
      const float4* src4 = reinterpret_cast<const float4*>(src);
#pragma unroll 8
      for (int i = lane; i < D * D / 4; i += LANES) {
        const float4 v = __ldg(src4 + i);
        const int rr = (4 * i) / D, cc = (4 * i) % D;
        float* dst = sm + rr * LD + cc;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
    } else {
      for (int i = lane; i < D * D; i += LANES) {
        const int rr = i / D, cc = i % D;
        sm[rr * LD + cc] = (rr < dr && cc < dr) ? src[rr * dr + cc] : (rr == cc ? 1.f : 0.f);
      }
    }
  }
  __syncwarp(gmask);

  float r[RPL][D];
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int m = 0; m < D; ++m) r[i][m] = sm[(lane + LANES * i) * LD + m];
  __syncwarp(gmask);

  bool used[RPL];
  int step_of[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) { used[i] = false; step_of[i] = 0; }
  unsigned permw[(D + 3) / 4];
#pragma unroll
  for (int i = 0; i < (D + 3) / 4; ++i) permw[i] = 0u;

  float lad = 0.f, sgn = 1.f;
  unsigned parity = 0u;
  bool singular = false;   // an exact zero pivot: sign 0, log|det| -inf (torch.linalg.slogdet)

#pragma unroll
  for (int j = 0; j < D; ++j) {
    // --- pivot search over unused rows: max |a[row][j]|, ties -> lowest row
    float bv = -1.f;
    int br = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      float v = used[i] ? -1.f : fabsf(r[i][j]);
      int row = lane + LANES * i;
      if (v > bv || (v == bv && row < br)) { bv = v; br = row; }
    }
#pragma unroll
    for (int off = LANES / 2; off >= 1; off >>= 1) {
      float ov = __shfl_xor_sync(gmask, bv, off, LANES);
      int orow = __shfl_xor_sync(gmask, br, off, LANES);
      if (ov > bv || (ov == bv && orow < br)) { bv = ov; br = orow; }
    }
    const int pl = br % LANES, pi = br / LANES;
    // permutation parity: number of still-unused rows with a smaller index than the pivot
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      unsigned b = __ballot_sync(gmask, !used[i] && (lane + LANES * i) < br) & gmask;
      parity ^= (unsigned)__popc(b) & 1u;
    }
    float mine = r[0][j];
#pragma unroll
    for (int i = 1; i < RPL; ++i) if (pi == i) mine = r[i][j];
    const float piv = __shfl_sync(gmask, mine, pl, LANES);
    lad += logf(fabsf(piv));
    if (piv < 0.f) sgn = -sgn;
    if (piv == 0.f) singular = true;
    const float pinv = 1.f / piv;
    permw[j / 4] |= (unsigned)br << (8 * (j % 4));

    // --- owner scales its pivot row in place
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      if (lane == pl && pi == i) {
#pragma unroll
        for (int m = 0; m < D; ++m) r[i][m] = (m == j) ? pinv : r[i][m] * pinv;
        used[i] = true;
        step_of[i] = j;
      }
    }
    // --- broadcast the scaled pivot row, eliminate column j everywhere else
    float pr[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
      float v = r[0][m];
#pragma unroll
      for (int i = 1; i < RPL; ++i) if (pi == i) v = r[i][m];
      pr[m] = __shfl_sync(gmask, v, pl, LANES);
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      if (!(lane == pl && pi == i)) {
        const float f = r[i][j];
        r[i][j] = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) r[i][m] = fmaf(-f, pr[m], r[i][m]);
      }
    }
  }

  // --- un-permute through shared memory: row step_of[i], column perm[m]
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int m = 0; m < D; ++m) {
      const int col = (permw[m / 4] >> (8 * (m % 4))) & 0xff;
      if (transpose_inv) sm[col * LD + step_of[i]] = r[i][m];
      else sm[step_of[i] * LD + col] = r[i][m];
    }
  __syncwarp(gmask);
  if (live) {
    if (inv != nullptr) {
      float* dst = inv + mat * dr * dr;
      if (D >= 8 && dr == D && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        float4* dst4 = reinterpret_cast<float4*>(dst);
#pragma unroll 8
        for (int i = lane; i < D * D / 4; i += LANES) {
          const float* sp = sm + ((4 * i) / D) * LD + (4 * i) % D;
          dst4[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
        }
      } else {
        for (int i = lane; i < dr * dr; i += LANES) dst[i] = sm[(i / dr) * LD + (i % dr)];
      }
    }
    if (diag_inv != nullptr)
      for (int i = lane; i < dr; i += LANES) diag_inv[mat * dr + i] = sm[i * LD + i];
    if (PACKED && packed_out != nullptr) {
      float* dst = packed_out + mat * kSymCols;
      for (int i = lane; i < D * D; i += LANES) {
        const int rr = i / D, cc = i % D;
        if (rr <= cc) dst[sym16_index(rr, cc)] = sm[rr * LD + cc];
      }
    }
    if (lane == 0) {
      if (logabsdet != nullptr) logabsdet[mat] = lad_scale * (singular ? -INFINITY : lad);
      if (sign != nullptr) sign[mat] = singular ? 0.f : (parity ? -sgn : sgn);
    }
  }
  __syncwarp(gmask);    // the staging rows are reused by the next slot of this group
  }
}

int launch_batched_inverse(const float* a, int64_t n, int d, float* inv, float* logabsdet,
                           float* sign, float* diag_inv, int transpose_inv, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(d >= 1 && d <= kMaxLatentDim, "batched_inverse: latent_dim must be in [1,64]");
  int dp = 1;                       // next power of two: the kernel embeds A as diag(A, I)
  while (dp < d) dp <<= 1;
  switch (dp) {
#define CASE(D)                                                                             \
  case D: {                                                                                 \
    unsigned grid = (unsigned)((n + PP<D>::MATS - 1) / PP<D>::MATS);                        \
    batched_inverse_kernel<D><<<grid, PP<D>::THREADS, 0, s>>>(a, n, inv, logabsdet, sign,  \
                                                               diag_inv, transpose_inv,     \
                                                               nullptr, nullptr, nullptr,   \
                                                               1.f, d);                     \
  } break;
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32) CASE(64)
#undef CASE
    default: RLVAE_REQUIRE(false, "batched_inverse: latent_dim must be in [1,64]");
  }
  RLVAE_LAUNCH_OK();
  return 0;
}

int launch_batched_inverse_packed16(const float* a_packed, int64_t n, float* inv, float* logabsdet,
                                    float* sign, float* diag_inv, int transpose_inv, cudaStream_t s) {
  if (n == 0) return 0;
  unsigned grid = (unsigned)((n + PP<16>::MATS - 1) / PP<16>::MATS);
  batched_inverse_kernel<16, true><<<grid, PP<16>::THREADS, 0, s>>>(a_packed, n, inv, logabsdet, sign,
                                                                     diag_inv, transpose_inv);
  RLVAE_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Symmetric positive-definite fast path (d == 16, packed [N,144] input from the symmetric tensor
// kernel): ONE THREAD PER MATRIX, the 136 packed entries in registers, everything statically indexed.
//   Cholesky A = L L^T  ->  L^{-1} in place  ->  G = L^{-T} L^{-1} in place (LAPACK potrf/trtri/lauum)
// ~2.2k FMAs per matrix and no shuffles (the 16-lane Gauss-Jordan above spends ~27k lane-instructions
// per matrix), so the kernel is bound by its HBM traffic: 576 B in + up to 576 + 64 + 8 B out per point.
// Global traffic is staged through shared memory (row stride 148 floats: 16-byte aligned and
// conflict-free for 128-bit accesses) so that every global access is a coalesced float4.
// A matrix whose Cholesky pivot is not > 0 (not positive definite -- the loader only warns about
// that, ref src/models/components/metric_loader.py:211-214) is appended to `fail_list`; the pivoting
// Gauss-Jordan kernel then redoes exactly those matrices (launch_sym16_inverse), which keeps the
// torch.linalg.inv / slogdet semantics for every input.
// ------------------------------------------------------------------------------------------------
namespace sym16 {
constexpr int THREADS = 64;
constexpr int LD = 148;
}
#define SYM_L(r, c) a[sym16_index((c), (r))]   /* lower-triangular entry (r >= c) */

__global__ void __launch_bounds__(sym16::THREADS)
sym16_cholesky_kernel(const float* __restrict__ a_packed, int64_t n, float* __restrict__ g_packed,
                      float* __restrict__ logabsdet, float lad_scale, float* __restrict__ sign,
                      float* __restrict__ diag_g, int* __restrict__ fail_count,
                      int* __restrict__ fail_list) {
  constexpr int LD = sym16::LD, TH = sym16::THREADS;
  __shared__ __align__(16) float stage[TH * LD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TH;
  const int rows = (int)((n - m0 < TH) ? (n - m0) : TH);
  {
    const float4* src = reinterpret_cast<const float4*>(a_packed + m0 * kSymCols);
    for (int i = tid; i < rows * 36; i += TH) {
      const int r = i / 36, c = i - r * 36;
      *reinterpret_cast<float4*>(stage + r * LD + 4 * c) = __ldg(src + i);
    }
  }
  __syncthreads();
  const bool live = tid < rows;
  float a[136];
#pragma unroll
  for (int q = 0; q < 34; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(stage + tid * LD + 4 * q);
    a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
  }
  if (!live) {          // keep the idle lanes finite
#pragma unroll
    for (int i = 0; i < 136; ++i) a[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[sym16_index(i, i)] = 1.f;
  }
  // ---- Cholesky (left-looking, column by column); rd[j] = 1 / L_jj
  float rd[16];
  float lad = 0.f;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float d = SYM_L(j, j);
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-SYM_L(j, k), SYM_L(j, k), d);
    ok = ok && (d > 0.f);
    lad += logf(d);
    // 1/sqrt(d) by MUFU.RSQ + one Newton step (< 1 ulp), L_jj = d / sqrt(d): this is the serial part of
    // the factorisation (16 dependent columns), sqrtf + an IEEE division were a third of its latency
    float inv = rsqrtf(d);
    inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);
    const float ljj = d * inv;
    rd[j] = inv;
    SYM_L(j, j) = ljj;
#pragma unroll
    for (int i = j + 1; i < 16; ++i) {
      float sacc = SYM_L(i, j);
#pragma unroll
      for (int k = 0; k < j; ++k) sacc = fmaf(-SYM_L(i, k), SYM_L(j, k), sacc);
      SYM_L(i, j) = sacc * inv;
    }
  }
  // ---- L^{-1} in place, column by column (columns > j still hold L)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    SYM_L(j, j) = rd[j];
#pragma unroll
    for (int i = j + 1; i < 16; ++i) {
      float sacc = 0.f;
#pragma unroll
      for (int k = j; k < i; ++k) sacc = fmaf(SYM_L(i, k), SYM_L(k, j), sacc);
      SYM_L(i, j) = -sacc * rd[i];
    }
  }
  if (diag_g != nullptr && live) {
    float dg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float sacc = 0.f;
#pragma unroll
      for (int k = i; k < 16; ++k) sacc = fmaf(SYM_L(k, i), SYM_L(k, i), sacc);
      dg[i] = sacc;
    }
    float4* dst = reinterpret_cast<float4*>(diag_g + (m0 + tid) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
  }
  if (live) {
    if (logabsdet != nullptr) logabsdet[m0 + tid] = lad_scale * lad;
    if (sign != nullptr) sign[m0 + tid] = 1.f;
    if (!ok) fail_list[atomicAdd(fail_count, 1)] = (int)(m0 + tid);
  }
  if (g_packed == nullptr) return;
  // ---- G = L^{-T} L^{-1} in place, row by row (row i is dead once G_i. is formed)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float sacc = 0.f;
#pragma unroll
      for (int k = i; k < 16; ++k) sacc = fmaf(SYM_L(k, i), SYM_L(k, j), sacc);
      SYM_L(i, j) = sacc;
    }
  }
#pragma unroll
  for (int q = 0; q < 34; ++q)
    *reinterpret_cast<float4*>(stage + tid * LD + 4 * q) = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
  *reinterpret_cast<float4*>(stage + tid * LD + 136) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(stage + tid * LD + 140) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(g_packed + m0 * kSymCols);
  for (int i = tid; i < rows * 36; i += TH) {
    const int r = i / 36, c = i - r * 36;
    dst[i] = *reinterpret_cast<const float4*>(stage + r * LD + 4 * c);
  }
}
#undef SYM_L

// Packed symmetric G^{-1} [N,144] -> any of { packed G [N,144], lad_scale * log|det G^{-1}|, sign,
// diag(G) }.  `fail_ws` holds 1 + n ints (counter, then the list of non-positive-definite matrices).
int launch_sym16_inverse(const float* a_packed, int64_t n, float* g_packed, float* logabsdet,
                         float lad_scale, float* sign, float* diag_g, int* fail_ws, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(fail_ws != nullptr, "sym16_inverse: workspace required");
  RLVAE_REQUIRE(n < (int64_t)1 << 31, "sym16_inverse: batch too large for the 32-bit fallback list");
  RLVAE_CUDA_OK(cudaMemsetAsync(fail_ws, 0, sizeof(int), s));
  const unsigned grid = (unsigned)((n + sym16::THREADS - 1) / sym16::THREADS);
  sym16_cholesky_kernel<<<grid, sym16::THREADS, 0, s>>>(a_packed, n, g_packed, logabsdet, lad_scale, sign,
                                                         diag_g, fail_ws, fail_ws + 1);
  RLVAE_LAUNCH_OK();
  return launch_sym16_fallback(a_packed, n, g_packed, logabsdet, lad_scale, sign, diag_g, fail_ws, s);
}

// The pivoting pass over the matrices a Cholesky kernel rejected (fail_ws = counter + list); usually
// the list is empty and every CTA exits after reading the counter.
// expanded [N,16,16] rows of the listed matrices from their packed rows (after the pivoting pass)
__global__ void unpack_sym16_list_kernel(const float* __restrict__ packed, const int* __restrict__ list,
                                         const int* __restrict__ count, float* __restrict__ full) {
  const int cnt = *count;
  for (int slot = blockIdx.x; slot < cnt; slot += gridDim.x) {
    const int64_t p = list[slot];
    const int e = threadIdx.x, i = e >> 4, j = e & 15;
    full[p * 256 + e] = packed[p * kSymCols + (i <= j ? sym16_index(i, j) : sym16_index(j, i))];
  }
}

// Packed G^{-1} rows of the LISTED points recomputed from the natural fp32 tables with exact
// differences (the arithmetic of the direct kernel).  Used when the fused tensor kernel was told not
// to store packed G^{-1} (tables certified positive semi-definite: a Cholesky failure can then only
// come from rounding, so paying 576 B per point for the fallback's input is waste); one CTA per
// listed point, thread p < 136 owns packed entry p.
__global__ void __launch_bounds__(160)
recompute_packed_rows_kernel(const float* __restrict__ z, const float* __restrict__ c,
                             const float* __restrict__ M, int K, float inv_T2, float lambda,
                             const int* __restrict__ list, const int* __restrict__ count,
                             float* __restrict__ a_packed) {
  __shared__ float w[160];
  __shared__ float zs[16];
  const int cnt = *count;
  const int tid = threadIdx.x;
  int pi = 0, pj = 0;                         // packed entry tid = (pi <= pj)
  if (tid < 136) {
    int base = 0;
    while (tid >= base + (16 - pi)) { base += 16 - pi; ++pi; }
    pj = pi + (tid - base);
  }
  for (int slot = blockIdx.x; slot < cnt; slot += gridDim.x) {
    const int64_t p = list[slot];
    __syncthreads();
    if (tid < 16) zs[tid] = z[p * 16 + tid];
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += 160) {
      __syncthreads();
      float wk = 0.f;
      if (k0 + tid < K) {
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float df = c[(int64_t)(k0 + tid) * 16 + j] - zs[j];
          sq = fmaf(df, df, sq);
        }
        wk = expf(-sq * inv_T2);
      }
      w[tid] = wk;
      __syncthreads();
      if (tid < 136) {
        const int kmax = min(160, K - k0);
        for (int kk = 0; kk < kmax; ++kk) {
          const float* m = M + (int64_t)(k0 + kk) * 256;
          acc = fmaf(w[kk], 0.5f * (m[pi * 16 + pj] + m[pj * 16 + pi]), acc);
        }
      }
    }
    if (tid < 136) a_packed[p * kSymCols + tid] = acc + (pi == pj ? lambda : 0.f);
    else if (tid < 144) a_packed[p * kSymCols + tid] = 0.f;
  }
}

// `recompute_from` (optional): the producer did not store packed G^{-1}; the rows of the listed points
// are first recomputed into a_packed (which must still be a valid [N,144] scratch) from these tables.
int launch_sym16_fallback(const float* a_packed, int64_t n, float* g_packed, float* logabsdet,
                          float lad_scale, float* sign, float* diag_g, int* fail_ws, cudaStream_t s,
                          float* g_full, const rlvae_tables* recompute_from, const float* z) {
  if (n == 0) return 0;
  const int64_t groups = (n + PP<16>::MATS - 1) / PP<16>::MATS;
  const unsigned fgrid = (unsigned)(groups < 1184 ? groups : 1184);
  if (recompute_from != nullptr) {
    const rlvae_tables* t = recompute_from;
    recompute_packed_rows_kernel<<<(unsigned)(n < 592 ? n : 592), 160, 0, s>>>(
        z, t->c, t->M, t->K, 1.f / t->T2, t->lambda, fail_ws + 1, fail_ws, const_cast<float*>(a_packed));
    RLVAE_LAUNCH_OK();
  }
  batched_inverse_kernel<16, true><<<fgrid, PP<16>::THREADS, 0, s>>>(
      a_packed, n, nullptr, logabsdet, sign, diag_g, 0, fail_ws + 1, fail_ws, g_packed, lad_scale);
  RLVAE_LAUNCH_OK();
  if (g_full != nullptr && g_packed != nullptr) {
    unpack_sym16_list_kernel<<<296, 256, 0, s>>>(g_packed, fail_ws + 1, fail_ws, g_full);
    RLVAE_LAUNCH_OK();
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Eigenvalues of symmetric 16 x 16 matrices (per-step spectrum / condition number of G^{-1}, G for
// the flow-analysis consumers: ref src/visualizations/flow_analysis.py:104-126,
// src/visualizations/manifold.py:79-101, src/models/modular_rlvae.py:434-457, which call
// torch.linalg.eigvals / det per batch).  ONE THREAD PER MATRIX: cyclic Jacobi on the 136 packed
// entries held in registers (every index static), sweeps until the whole warp has converged, then a
// sorting network; ascending order like torch.linalg.eigvalsh.  Input: packed [N,144] (PACKED) or
// the upper triangle of a full [N,16,16].  HBM-bound target: 576 (or 1024) B in, 64 B out per matrix.
// ------------------------------------------------------------------------------------------------
#define SYM_U(r, c) a[sym16_index((r), (c))]   /* r <= c */

// One Jacobi rotation (P, Q), P < Q, on the packed upper triangle; every index is a template constant.
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(float (&a)[136]) {
  constexpr int ipq = sym16_index(P, Q), ipp = sym16_index(P, P), iqq = sym16_index(Q, Q);
  const float apq = a[ipq], app = a[ipp], aqq = a[iqq];
  // rotation angle (Rutishauser); a negligible a_pq gives the identity rotation
  const bool skip = fabsf(apq) <= 1e-30f || apq * apq <= 1e-18f * fabsf(app * aqq);
  const float theta = 0.5f * (aqq - app) / (skip ? 1.f : apq);
  float t = 1.f / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
  t = theta < 0.f ? -t : t;
  t = skip ? 0.f : t;
  const float c = rsqrtf(fmaf(t, t, 1.f));
  const float sn = t * c;
  a[ipp] = fmaf(-t, apq, app);
  a[iqq] = fmaf(t, apq, aqq);
  a[ipq] = skip ? apq : 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (k != P && k != Q) {
      const int ikp = (k < P) ? sym16_index(k, P) : sym16_index(P, k);
      const int ikq = (k < Q) ? sym16_index(k, Q) : sym16_index(Q, k);
      const float x = a[ikp], y = a[ikq];
      a[ikp] = fmaf(c, x, -sn * y);
      a[ikq] = fmaf(sn, x, c * y);
    }
  }
}
template <int P, int Q>
struct JacobiSweep {
  static __device__ __forceinline__ void run(float (&a)[136]) {
    jacobi_rotate<P, Q>(a);
    if constexpr (Q + 1 < 16) JacobiSweep<P, Q + 1>::run(a);
    else if constexpr (P + 2 < 16) JacobiSweep<P + 1, P + 2>::run(a);
  }
};

template <bool PACKED>
__global__ void __launch_bounds__(sym16::THREADS)
sym16_eigvalsh_kernel(const float* __restrict__ src_all, int64_t n, float* __restrict__ eig) {
  constexpr int LD = sym16::LD, TH = sym16::THREADS;
  __shared__ __align__(16) float stage[TH * LD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TH;
  const int rows = (int)((n - m0 < TH) ? (n - m0) : TH);
  float a[136];
  if (PACKED) {
    const float4* src = reinterpret_cast<const float4*>(src_all + m0 * kSymCols);
    for (int i = tid; i < rows * 36; i += TH) {
      const int r = i / 36, c = i - r * 36;
      *reinterpret_cast<float4*>(stage + r * LD + 4 * c) = __ldg(src + i);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 34; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(stage + tid * LD + 4 * q);
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
  } else {
    // full [N,16,16]: stage 256 floats per matrix in four 64-column passes, keep the upper triangle
    const float4* src = reinterpret_cast<const float4*>(src_all + m0 * 256);
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {            // rows 4*pass .. 4*pass+3 of every matrix
      __syncthreads();
      for (int i = tid; i < rows * 16; i += TH) {
        const int r = i >> 4, c = i & 15;
        *reinterpret_cast<float4*>(stage + r * LD + 4 * c) = __ldg(src + (int64_t)r * 64 + pass * 16 + c);
      }
      __syncthreads();
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int cc = 0; cc < 16; ++cc)
          if (cc >= 4 * pass + rr) SYM_U(4 * pass + rr, cc) = stage[tid * LD + rr * 16 + cc];
    }
  }
  const bool live = tid < rows;
  if (!live) {
#pragma unroll
    for (int i = 0; i < 136; ++i) a[i] = 0.f;
  }
  float fro = 0.f;
#pragma unroll
  for (int i = 0; i < 136; ++i) fro = fmaf(a[i], a[i], fro);
  const float tol = 1e-14f * fro;          // off-diagonal sum of squares target (relative 1e-7 in norm)
  for (int sweep = 0; sweep < 12; ++sweep) {
    float off = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p)
#pragma unroll
      for (int q = p + 1; q < 16; ++q) off = fmaf(SYM_U(p, q), SYM_U(p, q), off);
    if (__all_sync(0xffffffffu, !(off > tol))) break;
    JacobiSweep<0, 1>::run(a);
  }
  float ev[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ev[i] = SYM_U(i, i);
  // odd-even transposition sort, ascending
#pragma unroll
  for (int round = 0; round < 16; ++round) {
#pragma unroll
    for (int i = (round & 1); i + 1 < 16; i += 2) {
      const float lo = fminf(ev[i], ev[i + 1]), hi = fmaxf(ev[i], ev[i + 1]);
      ev[i] = lo; ev[i + 1] = hi;
    }
  }
  if (live) {
    float4* dst = reinterpret_cast<float4*>(eig + (m0 + tid) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(ev[4 * q], ev[4 * q + 1], ev[4 * q + 2], ev[4 * q + 3]);
  }
}
#undef SYM_U

int launch_sym16_eigvalsh(const float* a, int64_t n, int packed, float* eig, cudaStream_t s) {
  if (n == 0) return 0;
  const unsigned grid = (unsigned)((n + sym16::THREADS - 1) / sym16::THREADS);
  if (packed) sym16_eigvalsh_kernel<true><<<grid, sym16::THREADS, 0, s>>>(a, n, eig);
  else sym16_eigvalsh_kernel<false><<<grid, sym16::THREADS, 0, s>>>(a, n, eig);
  RLVAE_LAUNCH_OK();
  return 0;
}

__global__ void unpack_sym16_kernel(const float* __restrict__ packed, int64_t n, float* __restrict__ full) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * 256) return;
  const int64_t p = gid >> 8;
  const int e = (int)(gid & 255), i = e >> 4, j = e & 15;
  full[gid] = packed[p * kSymCols + (i <= j ? sym16_index(i, j) : sym16_index(j, i))];
}

int launch_unpack_sym16(const float* a_packed, int64_t n, float* full, cudaStream_t s) {
  if (n == 0) return 0;
  const int64_t total = n * 256;
  unpack_sym16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a_packed, n, full);
  RLVAE_LAUNCH_OK();
  return 0;
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------- SPD 64 x 64
// Symmetric positive definite 64 x 64 matrices (G^{-1} of symmetric tables at latent_dim 64): log det and,
// optionally, the inverse WITHOUT pivoting, one warp per matrix, rows lane / lane + 32 in registers.
//   * symmetry gives coalesced global accesses with no shared-memory staging (row i of A == column i: lane l
//     reads / writes element (k, l) of every row k), so the kernel keeps 12 warps per SM instead of the staged
//     Gauss-Jordan's 33 KB per two matrices;
//   * the Schur complement of an SPD matrix stays symmetric, so the pivot ROW every elimination step needs is
//     the pivot COLUMN the lanes already hold: two shared-memory stores per lane + broadcast reads replace the
//     64 shuffles per step of batched_inverse_kernel, and there is no pivot search.
//   spd64_logdet_kernel: Gaussian elimination of the lower triangle only (2512 FMAs per lane), log det = sum of
//     the log pivots.  spd64_inverse_kernel: the symmetric sweep operator (pivot d: a_jj <- -1/d, a_ij <- a_ij/d,
//     a_ik <- a_ik - a_ij a_jk / d) applied to all 64 pivots leaves -A^{-1}.
// A pivot that is not > 0 (not positive definite, or NaN) appends the matrix to `fail` (fail[0] = count,
// fail[1..] = indices); the caller re-runs those through the pivoting Gauss-Jordan (torch.linalg.inv /
// slogdet semantics, ref src/models/components/metric_tensor.py:152,175).
// ---- log det only: elimination of the lower triangle, fully unrolled (4.2 k instructions per matrix)
__global__ void __launch_bounds__(128, 3)
spd64_logdet_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ logabsdet, float lad_scale,
                    int* __restrict__ fail) {
  __shared__ __align__(16) float colbuf[4][2][64];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t mat = (int64_t)blockIdx.x * 4 + wrp;
  if (mat >= n) return;                                    // (warp-uniform; no block-wide barrier below)
  const float* src = a + mat * 4096;
  float r0[32], r1[64];                                    // row lane needs columns <= 31 only
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    if (k < 32) r0[k] = __ldg(src + k * 64 + lane);        // A[lane][k] == A[k][lane]
    r1[k] = __ldg(src + k * 64 + 32 + lane);
  }
  float lad = 0.f, prod = 1.f;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    float* col = colbuf[wrp][j & 1];
    if (j < 32) col[lane] = r0[j];
    col[lane + 32] = r1[j];
    __syncwarp();
    const float d = col[j];
    if (!(d > 0.f)) bad = true;
    const float dinv = 1.f / d;
    prod *= d;
    if ((j & 7) == 7) { lad += logf(prod); prod = 1.f; }
    // rows <= j are finished: their multipliers are garbage, but they only touch entries above the diagonal,
    // which nothing reads
    const float f0 = (j < 31) ? r0[j] * dinv : 0.f, f1 = r1[j] * dinv;
#pragma unroll
    for (int k4 = ((j + 1) & ~3); k4 < 64; k4 += 4) {
      const float4 c = *reinterpret_cast<const float4*>(col + k4);
      const float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k4 + q;
        if (k <= j) continue;
        if (k < 32) r0[k] = fmaf(-f0, cv[q], r0[k]);
        r1[k] = fmaf(-f1, cv[q], r1[k]);
      }
    }
  }
  if (bad) {
    if (lane == 0) {
      const int slot = atomicAdd(fail, 1);
      fail[1 + slot] = (int)mat;
    }
    return;
  }
  if (lane == 0) logabsdet[mat] = lad_scale * lad;
}

// ---- inverse (+ log det, sign, diagonal): the symmetric sweep.  Fully unrolled it is 14.5 k instructions (232 KB of
// straight-line code per matrix) and bound by instruction FETCH (ncu: `no_instruction` 4.2 stalls per issue), so the
// pivots go in groups of 8: inside a group the pivot column is register u (compile time), between groups the
// column registers rotate by 8 -- register m holds column (m + j0) mod 64, and the shared-memory column is written
// at the rotated position, so its reads stay compile-time 128-bit loads.  SLOT = the row slot that holds the pivot
// rows of this half of the sweep (rows 0..31: slot 0), which keeps the pivot-row select out of the other slot.
template <int SLOT>
__device__ __forceinline__ void spd64_sweep_group(float (&r0)[64], float (&r1)[64], float* colw, int j0, int lane,
                                                  float& lad, bool& bad) {
  float prod = 1.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float* col = colw + (u & 1) * 64;
    col[(lane - j0) & 63] = r0[u];
    col[(lane + 32 - j0) & 63] = r1[u];
    __syncwarp();
    const float d = col[u];                                // row j0 + u sits at position u
    if (!(d > 0.f)) bad = true;
    const float dinv = 1.f / d;
    prod *= d;
    const bool own = lane == ((j0 + u) & 31);              // this lane's SLOT row is the pivot row
    const float f0 = r0[u] * dinv, f1 = r1[u] * dinv;
    const float x0 = (SLOT == 0 && own) ? dinv : -f0, x1 = (SLOT == 1 && own) ? dinv : -f1;
#pragma unroll
    for (int m4 = 0; m4 < 64; m4 += 4) {
      const float4 c = *reinterpret_cast<const float4*>(col + m4);
      const float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m4 + q;
        if (m == u) continue;
        // the pivot row becomes col / d, every other row r - (r_j / d) col
        if (SLOT == 0) {
          r0[m] = fmaf(x0, cv[q], own ? 0.f : r0[m]);
          r1[m] = fmaf(x1, cv[q], r1[m]);
        } else {
          r0[m] = fmaf(x0, cv[q], r0[m]);
          r1[m] = fmaf(x1, cv[q], own ? 0.f : r1[m]);
        }
      }
    }
    r0[u] = (SLOT == 0 && own) ? -dinv : f0;
    r1[u] = (SLOT == 1 && own) ? -dinv : f1;
  }
  lad += logf(prod);
  // rotate the column registers by one group
  float t0[8], t1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { t0[i] = r0[i]; t1[i] = r1[i]; }
#pragma unroll
  for (int m = 0; m < 56; ++m) { r0[m] = r0[m + 8]; r1[m] = r1[m + 8]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) { r0[56 + i] = t0[i]; r1[56 + i] = t1[i]; }
}

__global__ void __launch_bounds__(128, 2)
spd64_inverse_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ inv, float* __restrict__ logabsdet,
                     float* __restrict__ sign, float* __restrict__ diag_inv, float lad_scale, int* __restrict__ fail) {
  __shared__ __align__(16) float colbuf[4][2][64];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t mat = (int64_t)blockIdx.x * 4 + wrp;
  if (mat >= n) return;
  const float* src = a + mat * 4096;
  float r0[64], r1[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    r0[k] = __ldg(src + k * 64 + lane);                    // A[lane][k] == A[k][lane]
    r1[k] = __ldg(src + k * 64 + 32 + lane);
  }
  float lad = 0.f;
  bool bad = false;
#pragma unroll 1
  for (int j0 = 0; j0 < 32; j0 += 8) spd64_sweep_group<0>(r0, r1, &colbuf[wrp][0][0], j0, lane, lad, bad);
#pragma unroll 1
  for (int j0 = 32; j0 < 64; j0 += 8) spd64_sweep_group<1>(r0, r1, &colbuf[wrp][0][0], j0, lane, lad, bad);
  // (eight rotations by 8: the registers are back in natural column order; they hold -A^{-1})
  if (bad) {
    if (lane == 0) {
      const int slot = atomicAdd(fail, 1);
      fail[1 + slot] = (int)mat;
    }
    return;
  }
  if (inv != nullptr) {
    float* dst = inv + mat * 4096;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      dst[k * 64 + lane] = -r0[k];
      dst[k * 64 + 32 + lane] = -r1[k];
    }
  }
  if (diag_inv != nullptr) {
    float* dd = diag_inv + mat * 64;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (lane == k) { dd[k] = -r0[k]; dd[k + 32] = -r1[k + 32]; }
    }
  }
  if (lane == 0) {
    if (logabsdet != nullptr) logabsdet[mat] = lad_scale * lad;
    if (sign != nullptr) sign[mat] = 1.f;
  }
}

// 64 x 64 SPD batch: inv (optional, [n,64,64]), lad_scale * log det, sign (= 1), diag of the inverse (optional);
// `fail_ws`: 1 + n ints.  Matrices that are not positive definite fall through to the pivoting Gauss-Jordan.
int launch_spd64(const float* a, int64_t n, float* inv, float* logabsdet, float lad_scale, int* fail_ws,
                 cudaStream_t s, float* sign, float* diag_inv) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(n < ((int64_t)1 << 31), "spd64: batch too large for the fallback list");
  RLVAE_CUDA_OK(cudaMemsetAsync(fail_ws, 0, sizeof(int), s));
  const unsigned grid = (unsigned)((n + 3) / 4);
  if (inv != nullptr || diag_inv != nullptr)
    spd64_inverse_kernel<<<grid, 128, 0, s>>>(a, n, inv, logabsdet, sign, diag_inv, lad_scale, fail_ws);
  else {
    spd64_logdet_kernel<<<grid, 128, 0, s>>>(a, n, logabsdet, lad_scale, fail_ws);
    if (sign != nullptr) {
      RLVAE_LAUNCH_OK();
      fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sign, n, 1.f);
    }
  }
  RLVAE_LAUNCH_OK();
  const unsigned fgrid = (unsigned)(n / 2 < 296 ? (n + 1) / 2 : 296);
  batched_inverse_kernel<64><<<fgrid, PP<64>::THREADS, 0, s>>>(a, n, inv, logabsdet, sign, diag_inv, 0, fail_ws + 1,
                                                               fail_ws, nullptr, lad_scale, 64);
  RLVAE_LAUNCH_OK();
  return 0;
}

// out = L @ eps, L = cholesky(A + jitter I) (lower; only the lower triangle of A is read,
// like LAPACK potrf('L') behind torch.linalg.cholesky).  Right-looking, column by column.
template <int D>
__global__ void __launch_bounds__(PP<D>::THREADS)
chol_apply_kernel(const float* __restrict__ a, const float* __restrict__ eps, int64_t n,
                  float jitter, float* __restrict__ out, int32_t* __restrict__ status, int dr) {
  // dr <= D: real dimension; A is embedded as diag(A, I), eps padded with zeros
  using P = PP<D>;
  constexpr int LANES = P::LANES, RPL = P::RPL, LD = P::LD;
  __shared__ float stage[P::MATS * D * LD];
  __shared__ float epss[P::MATS * D];

  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t mat = (int64_t)blockIdx.x * P::MATS + grp;
  const bool live = mat < n;
  const int64_t msafe = live ? mat : (n - 1);
  float* sm = stage + grp * D * LD;
  float* es = epss + grp * D;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                       : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));

  const float* src = a + msafe * dr * dr;
  for (int i = lane; i < D * D; i += LANES) {
    const int rr = i / D, cc = i % D;
    sm[rr * LD + cc] = (rr < dr && cc < dr) ? src[rr * dr + cc] : (rr == cc ? 1.f : 0.f);
  }
  for (int i = lane; i < D; i += LANES) es[i] = (i < dr) ? eps[msafe * dr + i] : 0.f;
  __syncwarp(gmask);

  float r[RPL][D];
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int m = 0; m < D; ++m) {
      const int row = lane + LANES * i;
      r[i][m] = sm[row * LD + m] + ((m == row && row < dr) ? jitter : 0.f);
    }

  bool bad = false;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const float djj = __shfl_sync(gmask, r[j / LANES][j], j % LANES, LANES);
    if (!(djj > 0.f)) bad = true;
    const float ljj = sqrtf(djj);
    const float linv = 1.f / ljj;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int row = lane + LANES * i;
      r[i][j] = (row == j) ? ljj : r[i][j] * linv;
    }
#pragma unroll
    for (int m = j + 1; m < D; ++m) {
      const float lmj = __shfl_sync(gmask, r[m / LANES][j], m % LANES, LANES);
#pragma unroll
      for (int i = 0; i < RPL; ++i) r[i][m] = fmaf(-r[i][j], lmj, r[i][m]);
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int row = lane + LANES * i;
      float y = 0.f;
#pragma unroll
      for (int m = 0; m < D; ++m)
        if (m <= row) y = fmaf(r[i][m], es[m], y);
      if (row < dr) out[mat * dr + row] = bad ? __int_as_float(0x7fc00000) : y;
    }
    if (status != nullptr && lane == 0) status[mat] = bad ? 1 : 0;
  }
}

int launch_chol_apply(const float* a, const float* eps, int64_t n, int d, float jitter, float* out,
                      int32_t* status, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(d >= 1 && d <= kMaxLatentDim, "chol_apply: latent_dim must be in [1,64]");
  int dp = 1;
  while (dp < d) dp <<= 1;
  switch (dp) {
#define CASE(D)                                                                             \
  case D: {                                                                                 \
    unsigned grid = (unsigned)((n + PP<D>::MATS - 1) / PP<D>::MATS);                        \
    chol_apply_kernel<D><<<grid, PP<D>::THREADS, 0, s>>>(a, eps, n, jitter, out, status, d); \
  } break;
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32) CASE(64)
#undef CASE
    default: RLVAE_REQUIRE(false, "chol_apply: latent_dim must be in [1,64]");
  }
  RLVAE_LAUNCH_OK();
  return 0;
}

}  // namespace rlvae
