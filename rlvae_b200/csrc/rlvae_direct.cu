// CUDA-core ("direct") kernels of the metric hot path: fp32 FFMA, direct ||z-c||^2
// differences exactly like the reference (no expanded-norm cancellation), any
// latent_dim <= 64.  They are the robust path (small temperature, d != 16, small K)
// and the numerical cross-check of the tcgen05 path in rlvae_tc.cu.
//
//   inverse_metric_direct : ref src/models/components/metric_tensor.py:115-135
//   metric_grad_direct    : closed-form backward of the same expression
//                           (oracle.metric_backward)
//   nearest2              : ref src/models/samplers/riemannian_sampler.py:58-67
#include <cfloat>

#include "rlvae_internal.h"

namespace rlvae {

// ------------------------------------------------------------------------------------------
// G^{-1}[n, :] = sum_k w_nk * Mtab[k, :] + lambda * I      (an SGEMM with on-the-fly weights)
// tile: 64 points x 128 table columns per CTA, 32 centroids per smem stage, 4x8 per thread.
// `ncols` is the table row length (d*d, or d*d+d for the pythae-augmented table);
// lambda is added on the diagonal of the leading d x d block only.
// ------------------------------------------------------------------------------------------
constexpr int DM_BM = 64, DM_BN = 128, DM_KC = 32, DM_THREADS = 256;

__global__ void __launch_bounds__(DM_THREADS)
inverse_metric_direct_kernel(const float* __restrict__ z, const float* __restrict__ c,
                             const float* __restrict__ mtab, int64_t n, int K, int d, int ncols,
                             float T2, float lambda, float* __restrict__ out) {
  extern __shared__ float smem[];
  float* zs = smem;                       // [d][BM]   (transposed: conflict-free over points)
  float* cs = zs + d * DM_BM;             // [KC][d]
  float* ws = cs + DM_KC * d;             // [KC][BM]
  float* ms = ws + DM_KC * DM_BM;         // [KC][BN]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * DM_BM;
  const int col0 = blockIdx.y * DM_BN;

  for (int i = tid; i < DM_BM * d; i += DM_THREADS) {
    int p = i / d, j = i - p * d;
    int64_t r = row0 + p;
    zs[j * DM_BM + p] = (r < n) ? z[r * d + j] : 0.f;
  }

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += DM_KC) {
    __syncthreads();  // previous stage fully consumed (also orders the zs fill)
    for (int i = tid; i < DM_KC * d; i += DM_THREADS) {
      int kk = i / d;
      cs[i] = (k0 + kk < K) ? c[(int64_t)k0 * d + i] : 0.f;
    }
    for (int i = tid; i < DM_KC * DM_BN; i += DM_THREADS) {
      int kk = i / DM_BN, col = i - kk * DM_BN;
      int gc = col0 + col;
      ms[i] = (k0 + kk < K && gc < ncols) ? mtab[(int64_t)(k0 + kk) * ncols + gc] : 0.f;
    }
    __syncthreads();
    {
      const int p = tid & (DM_BM - 1);
#pragma unroll
      for (int i = 0; i < DM_KC / 4; ++i) {
        const int kk = (tid >> 6) + 4 * i;
        float sq = 0.f;
        for (int j = 0; j < d; ++j) {
          float df = cs[kk * d + j] - zs[j * DM_BM + p];
          sq = fmaf(df, df, sq);
        }
        ws[kk * DM_BM + p] = (k0 + kk < K) ? expf(-sq / T2) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < DM_KC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&ws[kk * DM_BM + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&ms[kk * DM_BN + tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ms[kk * DM_BN + 64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  const int dd = d * d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = col0 + h * 64 + tx * 4 + j;
        if (gc < ncols) {
          float v = acc[i][h * 4 + j];
          if (gc < dd && (gc / d) == (gc % d)) v += lambda;
          out[r * ncols + gc] = v;
        }
      }
    }
  }
}

static size_t dm_smem_bytes(int d) {
  return sizeof(float) * ((size_t)d * DM_BM + (size_t)DM_KC * d + DM_KC * DM_BM + DM_KC * DM_BN);
}

int launch_inverse_metric_direct(const rlvae_tables* t, const float* z, int64_t n, float* ginv,
                                 cudaStream_t s) {
  if (n == 0) return 0;
  const int ncols = t->d * t->d;
  dim3 grid((unsigned)((n + DM_BM - 1) / DM_BM), (unsigned)((ncols + DM_BN - 1) / DM_BN));
  size_t smem = dm_smem_bytes(t->d);
  RLVAE_OPT_IN_SMEM(inverse_metric_direct_kernel, (int)dm_smem_bytes(kMaxLatentDim));
  inverse_metric_direct_kernel<<<grid, DM_THREADS, smem, s>>>(z, t->c, t->M, n, t->K, t->d, ncols,
                                                             t->T2, t->lambda, ginv);
  RLVAE_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// out[n,:] = scale * sum_k w_nk <U_n, M_k> (c_k - z_n)
// per CTA: 64 points; per stage 64 centroids: T = U (64 x d^2) . M^T (d^2 x 64) as a 4x4
// register-tiled GEMM, then u = w*t through smem and a second small contraction over the
// stage's centroids.
// ------------------------------------------------------------------------------------------
constexpr int DG_BM = 64, DG_KC = 64, DG_QC = 16, DG_THREADS = 256, DG_PAD = 1;

__global__ void __launch_bounds__(DG_THREADS)
metric_grad_direct_kernel(const float* __restrict__ z, const float* __restrict__ u,
                          const float* __restrict__ c, const float* __restrict__ mtab, int64_t n,
                          int K, int d, float T2, float scale, float* __restrict__ out) {
  extern __shared__ float smem[];
  float* zs = smem;                                   // [d][BM]
  float* cs = zs + d * DG_BM;                         // [KC][d+1]
  float* us = cs + DG_KC * (d + 1);                   // [QC][BM+1]
  float* ms = us + DG_QC * (DG_BM + DG_PAD);          // [QC][KC+1]
  float* ut = ms + DG_QC * (DG_KC + DG_PAD);          // [BM][KC+1]   u = w * t

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * DG_BM;
  const int dd = d * d;

  for (int i = tid; i < DG_BM * d; i += DG_THREADS) {
    int p = i / d, j = i - p * d;
    int64_t r = row0 + p;
    zs[j * DG_BM + p] = (r < n) ? z[r * d + j] : 0.f;
  }

  // second-phase ownership: point gp, dims gj0 + 4*i
  const int gp = tid >> 2, gj0 = tid & 3;
  float gacc[kMaxLatentDim / 4];
#pragma unroll
  for (int i = 0; i < kMaxLatentDim / 4; ++i) gacc[i] = 0.f;

  for (int k0 = 0; k0 < K; k0 += DG_KC) {
    __syncthreads();
    for (int i = tid; i < DG_KC * d; i += DG_THREADS) {
      int kk = i / d, j = i - kk * d;
      cs[kk * (d + 1) + j] = (k0 + kk < K) ? c[(int64_t)(k0 + kk) * d + j] : 0.f;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int q0 = 0; q0 < dd; q0 += DG_QC) {
      __syncthreads();
      for (int i = tid; i < DG_BM * DG_QC; i += DG_THREADS) {
        int p = i / DG_QC, q = i - p * DG_QC;
        int64_t r = row0 + p;
        us[q * (DG_BM + DG_PAD) + p] = (r < n && q0 + q < dd) ? u[r * dd + q0 + q] : 0.f;
      }
      for (int i = tid; i < DG_KC * DG_QC; i += DG_THREADS) {
        int kk = i / DG_QC, q = i - kk * DG_QC;
        ms[q * (DG_KC + DG_PAD) + kk] =
            (k0 + kk < K && q0 + q < dd) ? mtab[(int64_t)(k0 + kk) * dd + q0 + q] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < DG_QC; ++q) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = us[q * (DG_BM + DG_PAD) + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = ms[q * (DG_KC + DG_PAD) + tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    // weights for this thread's 4x4 (point, centroid) pairs, then u = w * t
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = tx * 4 + j;
        float sq = 0.f;
        for (int e = 0; e < d; ++e) {
          float df = cs[kk * (d + 1) + e] - zs[e * DG_BM + p];
          sq = fmaf(df, df, sq);
        }
        float w = (k0 + kk < K) ? expf(-sq / T2) : 0.f;
        ut[p * (DG_KC + DG_PAD) + kk] = w * acc[i][j];
      }
    }
    __syncthreads();
    for (int kk = 0; kk < DG_KC; ++kk) {
      const float uv = ut[gp * (DG_KC + DG_PAD) + kk];
#pragma unroll
      for (int i = 0; i < kMaxLatentDim / 4; ++i) {
        const int j = gj0 + 4 * i;
        if (j < d) gacc[i] = fmaf(uv, cs[kk * (d + 1) + j] - zs[j * DG_BM + gp], gacc[i]);
      }
    }
  }
  const int64_t r = row0 + gp;
  if (r < n) {
#pragma unroll
    for (int i = 0; i < kMaxLatentDim / 4; ++i) {
      const int j = gj0 + 4 * i;
      if (j < d) out[r * d + j] = scale * gacc[i];
    }
  }
}

static size_t dg_smem_bytes(int d) {
  return sizeof(float) * ((size_t)d * DG_BM + (size_t)DG_KC * (d + 1) +
                          DG_QC * (DG_BM + DG_PAD) + DG_QC * (DG_KC + DG_PAD) +
                          DG_BM * (DG_KC + DG_PAD));
}

int launch_metric_grad_direct(const rlvae_tables* t, const float* z, const float* u, int64_t n,
                              float scale, float* out, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_OPT_IN_SMEM(metric_grad_direct_kernel, (int)dg_smem_bytes(kMaxLatentDim));
  dim3 grid((unsigned)((n + DG_BM - 1) / DG_BM));
  metric_grad_direct_kernel<<<grid, DG_THREADS, dg_smem_bytes(t->d), s>>>(
      z, u, t->c, t->M, n, t->K, t->d, t->T2, scale, out);
  RLVAE_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// variant C (pythae): out = (1/T^2) G^T v,  v = sum_k w_k M_k^T (c_k - z)
//   = sum_k w_k b_k - (G^{-1} - lambda I)^T z   with b_k = M_k^T c_k  folded into an
// augmented table [M_k | b_k] so the same weighted-sum kernel produces both terms.
// The augmented table is built lazily by the caller (tables struct owns it).
// ------------------------------------------------------------------------------------------
__global__ void pythae_finish_kernel(const float* __restrict__ aug, const float* __restrict__ z,
                                     const float* __restrict__ g, int64_t n, int d,
                                     float T2, float* __restrict__ out) {
  // one thread per (point, output dim)
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * d) return;
  const int64_t p = gid / d;
  const int j = (int)(gid - p * d);
  const int ncols = d * d + d;
  const float* row = aug + p * ncols;
  const float* zp = z + p * d;
  const float* gp = g + p * d * d;
  // v_i = B_i - sum_e S[e][i] z_e, S = sum_k w_k M_k (accumulated WITHOUT lambda: fl(S_ii + lambda) - lambda
  // would lose S_ii wherever the weights are small) ; out_j = (1/T2) sum_i G[i][j] v_i
  float o = 0.f;
  for (int i = 0; i < d; ++i) {
    float v = row[d * d + i];
    for (int e = 0; e < d; ++e) v = fmaf(-row[e * d + i], zp[e], v);
    o = fmaf(gp[i * d + j], v, o);
  }
  out[gid] = o / T2;
}

__global__ void build_aug_table_kernel(const float* __restrict__ c, const float* __restrict__ M,
                                       int K, int d, float* __restrict__ aug) {
  const int k = blockIdx.x;
  if (k >= K) return;
  const int dd = d * d, ncols = dd + d;
  for (int i = threadIdx.x; i < dd; i += blockDim.x) aug[(int64_t)k * ncols + i] = M[(int64_t)k * dd + i];
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float b = 0.f;  // b_k[j] = sum_i M_k[i][j] c_k[i]
    for (int i = 0; i < d; ++i) b = fmaf(M[(int64_t)k * dd + i * d + j], c[(int64_t)k * d + i], b);
    aug[(int64_t)k * ncols + dd + j] = b;
  }
}

// The augmented table [K, d*d+d] is a derived cache on the tables handle (built on first use, freed
// with the handle); the [N, d*d+d] scratch comes from the caller (rlvae_metric_grad_pythae_workspace),
// so nothing here is shared between devices, handles or streams.
int launch_metric_grad_pythae(const rlvae_tables* t, const float* z, const float* g, int64_t n,
                              float* out, float* scratch, cudaStream_t s) {
  if (n == 0) return 0;
  const int d = t->d, ncols = d * d + d;
  RLVAE_REQUIRE(scratch != nullptr, "metric_grad_pythae: workspace required");
  if (t->pythae_aug == nullptr) {
    float* aug = nullptr;
    RLVAE_CUDA_OK(cudaMalloc(&aug, sizeof(float) * (size_t)t->K * ncols));
    build_aug_table_kernel<<<t->K, 128, 0, s>>>(t->c, t->M, t->K, d, aug);
    RLVAE_LAUNCH_OK();
    RLVAE_CUDA_OK(cudaStreamSynchronize(s));      // other streams may use the cached table next
    t->pythae_aug = aug;
  }
  dim3 grid((unsigned)((n + DM_BM - 1) / DM_BM), (unsigned)((ncols + DM_BN - 1) / DM_BN));
  RLVAE_OPT_IN_SMEM(inverse_metric_direct_kernel, (int)dm_smem_bytes(kMaxLatentDim));
  inverse_metric_direct_kernel<<<grid, DM_THREADS, dm_smem_bytes(d), s>>>(
      z, t->c, t->pythae_aug, n, t->K, d, ncols, t->T2, 0.f /* no lambda: see pythae_finish_kernel */, scratch);
  RLVAE_LAUNCH_OK();
  const int64_t total = n * d;
  pythae_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(scratch, z, g, n, d, t->T2, out);
  RLVAE_LAUNCH_OK();
  return 0;
}

void pythae_cache_release(const rlvae_tables* t) {
  if (t->pythae_aug) cudaFree(t->pythae_aug);
  t->pythae_aug = nullptr;
}

// ------------------------------------------------------------------------------------------
// two nearest centroids (Euclidean), one thread per point, centroids staged through smem.
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
nearest2_kernel(const float* __restrict__ mu, const float* __restrict__ c, int64_t n, int K,
                int64_t* __restrict__ idx, float* __restrict__ dist, int dr) {
  // dr <= D: real latent_dim (rows of mu and c are dr floats); the extra coordinates are zero on both sides
  constexpr int KC = 64;
  __shared__ float cs[KC * D];
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float zr[D];
#pragma unroll
  for (int j = 0; j < D; ++j) zr[j] = (p < n && j < dr) ? mu[p * dr + j] : 0.f;
  float b0 = FLT_MAX, b1 = FLT_MAX;
  int i0 = 0, i1 = 0;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < KC * D; i += blockDim.x) {
      const int kk = i / D, j = i - kk * D;
      cs[i] = (k0 + kk < K && j < dr) ? c[(int64_t)(k0 + kk) * dr + j] : 0.f;
    }
    __syncthreads();
    const int kmax = min(KC, K - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        float df = zr[j] - cs[kk * D + j];
        sq = fmaf(df, df, sq);
      }
      if (sq < b0) { b1 = b0; i1 = i0; b0 = sq; i0 = k0 + kk; }
      else if (sq < b1) { b1 = sq; i1 = k0 + kk; }
    }
  }
  if (p < n) {
    idx[p * 2 + 0] = i0; idx[p * 2 + 1] = i1;
    dist[p * 2 + 0] = sqrtf(b0); dist[p * 2 + 1] = sqrtf(b1);
  }
}

int launch_nearest2(const rlvae_tables* t, const float* mu, int64_t n, int64_t* idx, float* dist,
                    cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(t->K >= 2, "nearest2 needs at least two centroids");
  unsigned grid = (unsigned)((n + 127) / 128);
  int dp = 1;                      // any latent_dim <= 64: zero-padded to the next power of two
  while (dp < t->d) dp <<= 1;
  switch (dp) {
#define CASE(D) case D: nearest2_kernel<D><<<grid, 128, 0, s>>>(mu, t->c, n, t->K, idx, dist, t->d); break;
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32) CASE(64)
#undef CASE
    default: RLVAE_REQUIRE(false, "nearest2: latent_dim must be in [1,64]");
  }
  RLVAE_LAUNCH_OK();
  return 0;
}

}  // namespace rlvae
