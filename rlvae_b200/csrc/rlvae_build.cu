// Metric construction: the local weighted-covariance matrices M_i that the
// reference extracts from a trained VAE before the hot path ever runs
//   ref scripts/train_and_extract_vanilla_vae.py:199-221
//     w_n   = exp(-||mu_n - c_i||^2 / T^2),  weights = w / (sum_n w_n + 1e-8)
//     mean  = sum_n weights_n mu_n,  cov_i = sum_n weights_n (mu_n - mean)(mu_n - mean)^T
// (M_i = cov_i + reg I, then the minimum-eigenvalue lift, are composed on the host side from this
// kernel and rlvae_sym_eigvalsh.)  The reference loops over centroids in Python, O(K) passes over
// all N latents; here one CTA owns one centroid (and, for d > 16, one slice of columns) and streams
// the latents once, accumulating in fp32 registers the moments CENTRED AT THE CENTROID
//   W = sum w,  s = sum w (mu - c),  C = sum w (mu - c)(mu - c)^T
// which avoids the cancellation of raw second moments; the exact covariance follows from
//   a = s / W', B = C / W', rho = W / W', mean = a + rho c, e = c - mean,
//   cov = B + a e^T + e a^T + rho e e^T           (W' = W + 1e-8)
#include "rlvae_internal.h"

namespace rlvae {

constexpr int LC_THREADS = 256;

// DP = padded latent dim (compile-time, <= 16): the CTA of centroid blockIdx.x accumulates
// 1 + DP + DP (DP + 1) / 2 moments per thread (upper triangle only).
template <int DP>
__global__ void __launch_bounds__(LC_THREADS)
local_covariance_kernel(const float* __restrict__ mus, int64_t n, const float* __restrict__ centroids, int d,
                        float inv_T2, float* __restrict__ cov) {
  constexpr int NT = DP * (DP + 1) / 2;
  constexpr int NE = 1 + DP + NT;
  const int i = blockIdx.x;
  __shared__ float cs[DP];
  __shared__ float red[LC_THREADS / 32][NE];
  __shared__ float tot[NE];
  if (threadIdx.x < DP) cs[threadIdx.x] = threadIdx.x < d ? centroids[(int64_t)i * d + threadIdx.x] : 0.f;
  __syncthreads();
  float c[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) c[j] = cs[j];
  float W = 0.f, s[DP], C[NT];
#pragma unroll
  for (int j = 0; j < DP; ++j) s[j] = 0.f;
#pragma unroll
  for (int e = 0; e < NT; ++e) C[e] = 0.f;
  for (int64_t p = threadIdx.x; p < n; p += LC_THREADS) {
    float x[DP];
    float d2 = 0.f;
    const float* row = mus + p * d;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      x[j] = (j < d) ? __ldg(row + j) - c[j] : 0.f;
      d2 = fmaf(x[j], x[j], d2);
    }
    // torch: norm -> square -> divide; exp in fp32 (expf, not the fast intrinsic: these are table values)
    const float nrm = sqrtf(d2);
    const float w = expf(-(nrm * nrm) * inv_T2);
    if (w == 0.f) continue;
    W += w;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      const float wx = w * x[j];
      s[j] += wx;
#pragma unroll
      for (int q = j; q < DP; ++q) C[j * DP - (j * (j - 1)) / 2 + (q - j)] = fmaf(wx, x[q], C[j * DP - (j * (j - 1)) / 2 + (q - j)]);
    }
  }
  // block reduction (fixed order -> deterministic): warp shuffles, then the 8 warp partials
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto wsum = [&](float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  {
    float v = wsum(W);
    if (lane == 0) red[warp][0] = v;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      v = wsum(s[j]);
      if (lane == 0) red[warp][1 + j] = v;
    }
#pragma unroll
    for (int e = 0; e < NT; ++e) {
      v = wsum(C[e]);
      if (lane == 0) red[warp][1 + DP + e] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NE; e += LC_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < LC_THREADS / 32; ++w8) v += red[w8][e];
    tot[e] = v;
  }
  __syncthreads();
  const float Wp = tot[0] + 1e-8f;
  const float rho = tot[0] / Wp;
  for (int e = threadIdx.x; e < d * d; e += LC_THREADS) {
    const int r = e / d, col = e % d;
    const int lo = r < col ? r : col, hi = r < col ? col : r;
    const float ar = tot[1 + r] / Wp, ac = tot[1 + col] / Wp;
    const float er = (1.f - rho) * cs[r] - ar, ec = (1.f - rho) * cs[col] - ac;
    const float B = tot[1 + DP + lo * DP - (lo * (lo - 1)) / 2 + (hi - lo)] / Wp;
    cov[((int64_t)i * d + r) * d + col] = B + ar * ec + er * ac + rho * er * ec;
  }
}

template <int DP>
__global__ void __launch_bounds__(LC_THREADS)
local_covariance_col_kernel(const float* __restrict__ mus, int64_t n, const float* __restrict__ centroids, int d,
                            float inv_T2, float* __restrict__ cov) {
  // one CTA = centroid blockIdx.x, covariance column blockIdx.y (d > 16: the full matrix does not fit
  // in registers; the weights are recomputed per column)
  const int i = blockIdx.x;
  const int col = blockIdx.y;
  __shared__ float cs[DP];
  __shared__ float red[LC_THREADS / 32][2 + 2 * DP];
  for (int j = threadIdx.x; j < DP; j += LC_THREADS) cs[j] = j < d ? centroids[(int64_t)i * d + j] : 0.f;
  __syncthreads();
  float W = 0.f, scol = 0.f, s[DP], C[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) { s[j] = 0.f; C[j] = 0.f; }
  for (int64_t p = threadIdx.x; p < n; p += LC_THREADS) {
    float x[DP];
    float d2 = 0.f;
    const float* row = mus + p * d;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      x[j] = (j < d) ? __ldg(row + j) - cs[j] : 0.f;
      d2 = fmaf(x[j], x[j], d2);
    }
    const float nrm = sqrtf(d2);
    const float w = expf(-(nrm * nrm) * inv_T2);
    if (w == 0.f) continue;
    const float xc = __ldg(row + col) - cs[col];
    W += w;
    const float wxc = w * xc;
    scol += wxc;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      s[j] = fmaf(w, x[j], s[j]);
      C[j] = fmaf(wxc, x[j], C[j]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto wsum = [&](float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  {
    float v = wsum(W);
    if (lane == 0) red[warp][0] = v;
    v = wsum(scol);
    if (lane == 0) red[warp][1] = v;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      v = wsum(s[j]);
      if (lane == 0) red[warp][2 + j] = v;
      v = wsum(C[j]);
      if (lane == 0) red[warp][2 + DP + j] = v;
    }
  }
  __syncthreads();
  __shared__ float tot[2 + 2 * DP];
  for (int e = threadIdx.x; e < 2 + 2 * DP; e += LC_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < LC_THREADS / 32; ++w8) v += red[w8][e];
    tot[e] = v;
  }
  __syncthreads();
  const float Wp = tot[0] + 1e-8f;
  const float rho = tot[0] / Wp;
  const float ac = tot[1] / Wp;
  const float ec = (1.f - rho) * cs[col] - ac;
  for (int r = threadIdx.x; r < d; r += LC_THREADS) {
    const float ar = tot[2 + r] / Wp;
    const float er = (1.f - rho) * cs[r] - ar;
    const float B = tot[2 + DP + r] / Wp;
    cov[((int64_t)i * d + r) * d + col] = B + ar * ec + er * ac + rho * er * ec;
  }
}

int launch_local_covariance(const float* mus, int64_t n, const float* centroids, int k, int d, float temperature,
                            float* cov, cudaStream_t s) {
  if (k == 0) return 0;
  const float inv_T2 = 1.f / (temperature * temperature);
  if (d <= 8) {
    local_covariance_kernel<8><<<dim3(k, 1), LC_THREADS, 0, s>>>(mus, n, centroids, d, inv_T2, cov);
  } else if (d <= 16) {
    local_covariance_kernel<16><<<dim3(k, 1), LC_THREADS, 0, s>>>(mus, n, centroids, d, inv_T2, cov);
  } else if (d <= 32) {
    local_covariance_col_kernel<32><<<dim3(k, d), LC_THREADS, 0, s>>>(mus, n, centroids, d, inv_T2, cov);
  } else {
    RLVAE_REQUIRE(d <= 64, "local_covariance: latent_dim must be <= 64");
    local_covariance_col_kernel<64><<<dim3(k, d), LC_THREADS, 0, s>>>(mus, n, centroids, d, inv_T2, cov);
  }
  RLVAE_LAUNCH_OK();
  return 0;
}

}  // namespace rlvae
