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
template <int D, bool PACKED = false>
__global__ void __launch_bounds__(PP<D>::THREADS)
batched_inverse_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ inv,
                       float* __restrict__ logabsdet, float* __restrict__ sign,
                       float* __restrict__ diag_inv, int transpose_inv) {
  using P = PP<D>;
  constexpr int LANES = P::LANES, RPL = P::RPL, LD = P::LD;
  __shared__ float stage[P::MATS * D * LD];

  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t mat = (int64_t)blockIdx.x * P::MATS + grp;
  const bool live = mat < n;
  const int64_t msafe = live ? mat : (n - 1);
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
    const float* src = a + msafe * D * D;
    for (int i = lane; i < D * D; i += LANES) sm[(i / D) * LD + (i % D)] = src[i];
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
      float* dst = inv + mat * D * D;
      for (int i = lane; i < D * D; i += LANES) dst[i] = sm[(i / D) * LD + (i % D)];
    }
    if (diag_inv != nullptr)
      for (int i = lane; i < D; i += LANES) diag_inv[mat * D + i] = sm[i * LD + i];
    if (lane == 0) {
      if (logabsdet != nullptr) logabsdet[mat] = singular ? -INFINITY : lad;
      if (sign != nullptr) sign[mat] = singular ? 0.f : (parity ? -sgn : sgn);
    }
  }
}

int launch_batched_inverse(const float* a, int64_t n, int d, float* inv, float* logabsdet,
                           float* sign, float* diag_inv, int transpose_inv, cudaStream_t s) {
  if (n == 0) return 0;
  switch (d) {
#define CASE(D)                                                                             \
  case D: {                                                                                 \
    unsigned grid = (unsigned)((n + PP<D>::MATS - 1) / PP<D>::MATS);                        \
    batched_inverse_kernel<D><<<grid, PP<D>::THREADS, 0, s>>>(a, n, inv, logabsdet, sign,  \
                                                               diag_inv, transpose_inv);    \
  } break;
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
    default: RLVAE_REQUIRE(false, "batched_inverse: latent_dim must be 1,2,4,8,16 or 32");
  }
  RLVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_batched_inverse_packed16(const float* a_packed, int64_t n, float* inv, float* logabsdet,
                                    float* sign, float* diag_inv, int transpose_inv, cudaStream_t s) {
  if (n == 0) return 0;
  unsigned grid = (unsigned)((n + PP<16>::MATS - 1) / PP<16>::MATS);
  batched_inverse_kernel<16, true><<<grid, PP<16>::THREADS, 0, s>>>(a_packed, n, inv, logabsdet, sign,
                                                                     diag_inv, transpose_inv);
  RLVAE_CUDA_OK(cudaGetLastError());
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
  RLVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

// out = L @ eps, L = cholesky(A + jitter I) (lower; only the lower triangle of A is read,
// like LAPACK potrf('L') behind torch.linalg.cholesky).  Right-looking, column by column.
template <int D>
__global__ void __launch_bounds__(PP<D>::THREADS)
chol_apply_kernel(const float* __restrict__ a, const float* __restrict__ eps, int64_t n,
                  float jitter, float* __restrict__ out, int32_t* __restrict__ status) {
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

  const float* src = a + msafe * D * D;
  for (int i = lane; i < D * D; i += LANES) sm[(i / D) * LD + (i % D)] = src[i];
  for (int i = lane; i < D; i += LANES) es[i] = eps[msafe * D + i];
  __syncwarp(gmask);

  float r[RPL][D];
#pragma unroll
  for (int i = 0; i < RPL; ++i)
#pragma unroll
    for (int m = 0; m < D; ++m) {
      const int row = lane + LANES * i;
      r[i][m] = sm[row * LD + m] + ((m == row) ? jitter : 0.f);
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
      out[mat * D + row] = bad ? __int_as_float(0x7fc00000) : y;
    }
    if (status != nullptr && lane == 0) status[mat] = bad ? 1 : 0;
  }
}

int launch_chol_apply(const float* a, const float* eps, int64_t n, int d, float jitter, float* out,
                      int32_t* status, cudaStream_t s) {
  if (n == 0) return 0;
  switch (d) {
#define CASE(D)                                                                             \
  case D: {                                                                                 \
    unsigned grid = (unsigned)((n + PP<D>::MATS - 1) / PP<D>::MATS);                        \
    chol_apply_kernel<D><<<grid, PP<D>::THREADS, 0, s>>>(a, eps, n, jitter, out, status);  \
  } break;
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
    default: RLVAE_REQUIRE(false, "chol_apply: latent_dim must be 1,2,4,8,16 or 32");
  }
  RLVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rlvae
