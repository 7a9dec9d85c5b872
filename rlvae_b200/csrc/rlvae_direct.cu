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
//   ref src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:160-187
// The difference c_k - z is formed PER CENTROID, as the reference does.  (Splitting the sum into
// sum_k w_k M_k^T c_k - (sum_k w_k M_k)^T z -- two table contractions, which is what the tensor path does -- loses
// |c| / |c_k - z| digits next to a centroid, multiplied by the condition number of G^{-1} when G is applied: measured
// 1e-2 relative at T = 0.1, lambda = 1e-3 against 4e-5 for the reference's own fp32 arithmetic.  This kernel is the
// accurate form: the whole CUDA-core path, and the fallback of the tensor path for the points its error bound flags.)
// One CTA owns P points: per tile of 32 centroids the threads first form w_k (c_k - z) for every (point, centroid)
// in shared memory, then thread e = (r, q) multiplies M_k[r][q] (one coalesced row of M per centroid, reused by the
// P points) into its per-point partial of v_q; the partials are reduced over r at the end and G^T is applied.
// ------------------------------------------------------------------------------------------
constexpr int PX_THREADS = 256;
constexpr int PX_KT = 32;

__device__ __forceinline__ int packed16_index(int r, int cidx) {
  const int lo = r < cidx ? r : cidx, hi = r < cidx ? cidx : r;
  return lo * 16 - (lo * (lo - 1)) / 2 + (hi - lo);
}

template <int P, int SLOTS>
__global__ void __launch_bounds__(PX_THREADS)
pythae_exact_kernel(const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ M, int K,
                    int Kpad, int d, float T2, const float* __restrict__ g, int g_is_packed,
                    const int* __restrict__ list, const int* __restrict__ count, int64_t n,
                    float* __restrict__ partial /* [n, gridDim.y, d] when the centroids are split over blockIdx.y */,
                    float cut /* > 0: skip centroid tiles whose weights are all below exp(-cut / T^2) of the point's
                                 largest weight (cut = 50 ln 2 T^2: 2^-50, far below fp32 resolution of the sum) */,
                    float* __restrict__ out) {
  extern __shared__ float px_smem[];
  const int dd = d * d;
  float* zs = px_smem;                          // [P][d]
  float* wd = zs + P * d;                       // [KT][d][P]   w_k (c_k - z), the P points contiguous
  float* red = wd + P * PX_KT * d;              // [P][dd]      partials of v before the reduction over r
  float* vs = red + P * dd;                     // [P][d]
  __shared__ int64_t ids[P];
  __shared__ int dmin_s[P];                     // smallest squared distance of each point to a centroid (float bits)
  const int tid = threadIdx.x;
  const int64_t total = (list != nullptr) ? (int64_t)*count : n;
  // centroid range of this CTA (whole tiles of PX_KT; Kpad is a multiple of PX_KT)
  const int tiles_k = Kpad / PX_KT;
  const int per = (tiles_k + (int)gridDim.y - 1) / (int)gridDim.y;
  const int kbeg = (int)blockIdx.y * per * PX_KT;
  const int kend = min(Kpad, kbeg + per * PX_KT);
  for (int64_t tile = blockIdx.x; tile * P < total; tile += gridDim.x) {
    __syncthreads();
    if (tid < P) {
      const int64_t i = tile * P + tid;
      ids[tid] = (i < total) ? (list != nullptr ? (int64_t)list[i] : i) : -1;
    }
    __syncthreads();
    for (int i = tid; i < P * d; i += PX_THREADS) {
      const int p = i / d, j = i - p * d;
      zs[i] = ids[p] >= 0 ? z[ids[p] * d + j] : 0.f;
    }
    float acc[P][SLOTS];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) acc[p][sl] = 0.f;
    if (tid < P) dmin_s[tid] = 0x7f800000;       // +inf
    __syncthreads();
    if (cut > 0.f) {
      // pass 0: distance of every point to its nearest centroid.  At small temperatures all but a few centroids
      // carry weights that vanish against the largest one; whole tiles of them are then skipped below.
      float dmin = __int_as_float(0x7f800000);
      const int p = tid % P;                     // (PX_THREADS is a multiple of P: the same point in every round)
      for (int i = tid; i < P * Kpad; i += PX_THREADS) {
        const int k = i / P;
        if (k >= K) break;
        const float* crow = c + (int64_t)k * d;
        float sq = 0.f;
        if (SLOTS == 1 && d == 16) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(crow) + q4);
            float df = cv.x - zs[p * 16 + 4 * q4 + 0]; sq = fmaf(df, df, sq);
            df = cv.y - zs[p * 16 + 4 * q4 + 1]; sq = fmaf(df, df, sq);
            df = cv.z - zs[p * 16 + 4 * q4 + 2]; sq = fmaf(df, df, sq);
            df = cv.w - zs[p * 16 + 4 * q4 + 3]; sq = fmaf(df, df, sq);
          }
        } else {
          for (int j = 0; j < d; ++j) {
            const float df = crow[j] - zs[p * d + j];
            sq = fmaf(df, df, sq);
          }
        }
        dmin = fminf(dmin, sq);                  // (a NaN distance is ignored here and counts as live below)
      }
      atomicMin(&dmin_s[p], __float_as_int(dmin));     // sq >= 0: the int order is the float order
    }
    for (int k0 = kbeg; k0 < kend; k0 += PX_KT) {
      __syncthreads();                           // zs / dmin_s ready / previous tile's wd consumed
      int live = 0;
      for (int i = tid; i < P * PX_KT; i += PX_THREADS) {
        const int p = i % P, kk = i / P;
        const int k = k0 + kk;
        const float* crow = c + (int64_t)k * d;
        float sq = 0.f;
        if (SLOTS == 1 && d == 16) {             // the common latent_dim: one 64-byte row, differences stay in registers
          float df[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(crow) + q4);
            df[4 * q4 + 0] = cv.x - zs[p * 16 + 4 * q4 + 0];
            df[4 * q4 + 1] = cv.y - zs[p * 16 + 4 * q4 + 1];
            df[4 * q4 + 2] = cv.z - zs[p * 16 + 4 * q4 + 2];
            df[4 * q4 + 3] = cv.w - zs[p * 16 + 4 * q4 + 3];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) sq = fmaf(df[j], df[j], sq);
          const float nrm = sqrtf(sq);           // ref :170-176: exp(-norm(c - z)^2 / T^2)
          const float w = (k < K && ids[p] >= 0) ? expf(-(nrm * nrm) / T2) : 0.f;
          live |= (k < K && ids[p] >= 0 && !(sq >= __int_as_float(dmin_s[p]) + cut)) ? 1 : 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) wd[(kk * 16 + j) * P + p] = w * df[j];
        } else {
          for (int j = 0; j < d; ++j) {
            const float df = crow[j] - zs[p * d + j];
            sq = fmaf(df, df, sq);
          }
          const float nrm = sqrtf(sq);
          const float w = (k < K && ids[p] >= 0) ? expf(-(nrm * nrm) / T2) : 0.f;
          live |= (k < K && ids[p] >= 0 && !(sq >= __int_as_float(dmin_s[p]) + cut)) ? 1 : 0;
          for (int j = 0; j < d; ++j) wd[(kk * d + j) * P + p] = w * (crow[j] - zs[p * d + j]);
        }
      }
      if (!__syncthreads_or(cut > 0.f ? live : 1)) continue;     // every weight of this tile vanishes: nothing to add
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        const int e = tid + sl * PX_THREADS;
        if (e < dd) {
          const int r = e / d;
          const float* mrow = M + (int64_t)k0 * dd + e;
#pragma unroll
          for (int kb = 0; kb < PX_KT; kb += 16) {
          float m[16];
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) m[kk] = __ldg(mrow + (int64_t)(kb + kk) * dd);     // 16 loads in flight
#pragma unroll
          for (int kq = 0; kq < 16; ++kq) {
            const int kk = kb + kq;
            const float* wrow = wd + (kk * d + r) * P;
            if (P % 4 == 0) {
#pragma unroll
              for (int p4 = 0; p4 < P / 4; ++p4) {
                const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * p4);
                acc[4 * p4 + 0][sl] = fmaf(m[kq], wv.x, acc[4 * p4 + 0][sl]);
                acc[4 * p4 + 1][sl] = fmaf(m[kq], wv.y, acc[4 * p4 + 1][sl]);
                acc[4 * p4 + 2][sl] = fmaf(m[kq], wv.z, acc[4 * p4 + 2][sl]);
                acc[4 * p4 + 3][sl] = fmaf(m[kq], wv.w, acc[4 * p4 + 3][sl]);
              }
            } else {
#pragma unroll
              for (int p = 0; p < P; ++p) acc[p][sl] = fmaf(m[kq], wrow[p], acc[p][sl]);
            }
          }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int e = tid + sl * PX_THREADS;
      if (e < dd) {
#pragma unroll
        for (int p = 0; p < P; ++p) red[p * dd + e] = acc[p][sl];
      }
    }
    __syncthreads();
    for (int i = tid; i < P * d; i += PX_THREADS) {       // v_q = sum_r M[r][q] (c - z)_r
      const int p = i / d, q = i - p * d;
      float v = 0.f;
      for (int r = 0; r < d; ++r) v += red[p * dd + r * d + q];
      vs[i] = v;
      if (gridDim.y > 1 && ids[p] >= 0) partial[(ids[p] * gridDim.y + blockIdx.y) * d + q] = v;
    }
    if (gridDim.y > 1) continue;                          // pythae_apply_g_kernel sums the splits and applies G
    __syncthreads();
    for (int i = tid; i < P * d; i += PX_THREADS) {       // out_j = (1/T^2) sum_i G[i][j] v_i
      const int p = i / d, j = i - p * d;
      if (ids[p] < 0) continue;
      float o = 0.f;
      if (g_is_packed) {
        const float* gp = g + ids[p] * kSymCols;
        for (int q = 0; q < d; ++q) o = fmaf(gp[packed16_index(q, j)], vs[p * d + q], o);
      } else {
        const float* gp = g + ids[p] * dd;
        for (int q = 0; q < d; ++q) o = fmaf(gp[q * d + j], vs[p * d + q], o);
      }
      out[ids[p] * d + j] = o / T2;
    }
  }
}

// short batches: the centroids are split over several CTAs per point; this sums their partial v (fixed order) and
// applies G^T / T^2.  One thread per (point, output dim).
__global__ void pythae_apply_g_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ g,
                                      int g_is_packed, int64_t n, int d, float T2, float* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * d) return;
  const int64_t p = gid / d;
  const int j = (int)(gid - p * d);
  float o = 0.f;
  for (int q = 0; q < d; ++q) {
    float v = 0.f;
    for (int sidx = 0; sidx < splits; ++sidx) v += partial[(p * splits + sidx) * d + q];
    const float gq = g_is_packed ? g[p * kSymCols + packed16_index(q, j)] : g[p * d * d + q * d + j];
    o = fmaf(gq, v, o);
  }
  out[gid] = o / T2;
}

template <int P, int SLOTS>
static int launch_pythae_exact_t(const rlvae_tables* t, const float* z, const float* g, int g_is_packed, const int* list,
                                 const int* count, int64_t n, float* out, float* partial, int splits, cudaStream_t s) {
  const int d = t->d;
  const size_t smem = sizeof(float) * ((size_t)P * d * (2 + PX_KT + d));
  auto kern = pythae_exact_kernel<P, SLOTS>;
  constexpr int DMAX = SLOTS == 1 ? 16 : (SLOTS == 4 ? 32 : 64);      // largest latent_dim this instantiation serves
  RLVAE_OPT_IN_SMEM(kern, (int)(sizeof(float) * ((size_t)P * DMAX * (2 + PX_KT + DMAX))));
  int64_t tiles = (n + P - 1) / P;
  // a device-side list is usually short: a grid-stride launch of a few waves covers any length
  const int64_t cap = (list != nullptr) ? 148 * 8 : ((int64_t)1 << 30);
  dim3 grid((unsigned)(tiles < cap ? tiles : cap), (unsigned)splits);
  // tile skipping needs the nearest-centroid pass over ALL centroids: not worth it when the centroids are split
  const float cut = (splits > 1) ? 0.f : 34.657359f * t->T2;           // 50 ln 2 T^2
  kern<<<grid, PX_THREADS, smem, s>>>(z, t->c, t->M, t->K, t->Kpad, d, t->T2, g, g_is_packed, list, count, n, partial,
                                      cut, out);
  RLVAE_LAUNCH_OK();
  if (splits > 1) {
    const int64_t total = n * d;
    pythae_apply_g_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(partial, splits, g, g_is_packed, n, d, t->T2,
                                                                          out);
    RLVAE_LAUNCH_OK();
  }
  return 0;
}

// g: [n,d,d] (g_is_packed == 0) or, d == 16, packed [n,144].  list / count (device): the rows to (re)compute, or
// NULL for all n rows.  A long batch gets 8 (d <= 16) or 2 (d <= 32) points per CTA (each row of M is then read once
// per CTA, not once per point); a short one (n <= kPythaeSplitBatch, no list) one CTA per (point, slice of the
// centroids) so that the whole GPU works on it -- `partial` must then hold n * kPythaeMaxSplits * d floats.
int launch_pythae_exact(const rlvae_tables* t, const float* z, const float* g, int g_is_packed, const int* list,
                        const int* count, int64_t n, float* out, float* partial, cudaStream_t s) {
  if (n == 0) return 0;
  RLVAE_REQUIRE(!g_is_packed || t->d == 16, "packed G is a latent_dim == 16 layout");
  const int dd = t->d * t->d;
  const bool small = n <= kPythaeSplitBatch;
  int splits = 1;
  if (small && list == nullptr && partial != nullptr) {
    const int tiles_k = t->Kpad / PX_KT;
    splits = (int)((148 * 4 + n - 1) / n);                       // ~4 CTAs per SM in total
    if (splits > tiles_k / 4) splits = tiles_k / 4;              // at least 4 centroid tiles per CTA
    if (splits > kPythaeMaxSplits) splits = kPythaeMaxSplits;
    if (splits < 1) splits = 1;
  }
  if (dd <= PX_THREADS)
    return small ? launch_pythae_exact_t<1, 1>(t, z, g, g_is_packed, list, count, n, out, partial, splits, s)
                 : launch_pythae_exact_t<8, 1>(t, z, g, g_is_packed, list, count, n, out, partial, 1, s);
  if (dd <= 4 * PX_THREADS)
    return small ? launch_pythae_exact_t<1, 4>(t, z, g, g_is_packed, list, count, n, out, partial, splits, s)
                 : launch_pythae_exact_t<2, 4>(t, z, g, g_is_packed, list, count, n, out, partial, 1, s);
  return launch_pythae_exact_t<1, 16>(t, z, g, g_is_packed, list, count, n, out, partial, small ? splits : 1, s);
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
