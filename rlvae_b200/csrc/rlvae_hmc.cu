// Elementwise stages of the fused HMC leapfrog ("K5").  The metric evaluation itself is the
// tensor/direct kernel + batched_inverse; these kernels turn (diag G, log|det G^{-1}|) into the
// momentum / position updates of ref src/models/samplers/hmc_sampler.py:120-163, so that one
// leapfrog step is exactly ONE metric evaluation (the reference re-evaluates the same gradient
// at the end of step k and the start of step k+1: SURVEY.md §3.2).
//
// One thread per chain; each thread touches a contiguous 4*d-byte row of every [N,d] array, with 128-bit
// loads / stores when d is a multiple of 4 (VEC).  This is the per-step path; certified d = 16 tables run the
// whole trajectory inside the fused kernel instead (rlvae_tc16.cu, HMC mode).
#include <cmath>

#include "rlvae_internal.h"

namespace rlvae {

// 0.5*log(clamp(det G^{-1}, 1e-10))  -- ref hmc_sampler.py:26-30 (det, clamp, log).
__device__ __forceinline__ float log_pi_from_slogdet(float lad, float sgn) {
  const float floor_lp = 0.5f * logf(1e-10f);
  if (!(sgn > 0.f)) return floor_lp;          // det <= 0 (or NaN sign) clamps to 1e-10
  if (lad > 88.72283f) return INFINITY;       // fp32 det overflows -> log(inf)
  return fmaxf(0.5f * lad, floor_lp);
}

// -grad_func(z)[j]: variant A == -(1 - lambda*G_jj)/T^2 (ref hmc_sampler.py:33-68)
__device__ __forceinline__ float neg_grad(int mode, float diag_g, float gexact, float lambda,
                                          float T2) {
  return mode == RLVAE_GRAD_MODULAR ? -((1.f - lambda * diag_g) / T2) : -gexact;
}

__global__ void hmc_begin_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                 const float* __restrict__ diag_g, const float* __restrict__ lad,
                                 const float* __restrict__ sgn, const float* __restrict__ gex,
                                 int64_t n, int d, float b0, float eps, float lambda, float T2,
                                 int mode, float* __restrict__ rho_half, float* __restrict__ z_new,
                                 float* __restrict__ h0) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float half_eps = eps / 2.f;
  float ss = 0.f;
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    const float rho = gamma[e] / b0;                                  // line 123
    ss = fmaf(rho, rho, ss);
    const float g = neg_grad(mode, diag_g[e], gex ? gex[e] : 0.f, lambda, T2);   // line 132
    const float rh = rho - half_eps * g;                              // line 135
    rho_half[e] = rh;
    z_new[e] = z[e] + eps * rh;                                       // line 138
  }
  const float nrm = sqrtf(ss);
  h0[p] = -log_pi_from_slogdet(lad[p], sgn[p]) + 0.5f * (nrm * nrm);  // line 127
}

__global__ void hmc_step_kernel(const float* __restrict__ diag_g, const float* __restrict__ lad,
                                const float* __restrict__ sgn, const float* __restrict__ gex,
                                int64_t n, int d, float eps, float lambda, float T2, int mode,
                                float scale, int last, float* __restrict__ rho_half,
                                float* z_cur /* may alias z_out */, const float* __restrict__ z_prev,
                                const float* __restrict__ acc, const float* __restrict__ h0,
                                float* __restrict__ h1, float* __restrict__ alpha_out,
                                float* __restrict__ moves, float* z_out /* may alias z_cur */) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float half_eps = eps / 2.f;
  if (!last) {
    for (int j = 0; j < d; ++j) {
      const int64_t e = p * d + j;
      const float g = neg_grad(mode, diag_g[e], gex ? gex[e] : 0.f, lambda, T2);  // line 141
      const float rho = scale * (rho_half[e] - half_eps * g);        // lines 144-148
      const float rh = rho - half_eps * g;                           // next step, lines 132-135
      rho_half[e] = rh;
      z_cur[e] = z_cur[e] + eps * rh;                                // line 138
    }
    return;
  }
  float ss = 0.f;
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    const float g = neg_grad(mode, diag_g[e], gex ? gex[e] : 0.f, lambda, T2);
    const float rho = scale * (rho_half[e] - half_eps * g);
    ss = fmaf(rho, rho, ss);
  }
  const float nrm = sqrtf(ss);
  const float H = -log_pi_from_slogdet(lad[p], sgn[p]) + 0.5f * (nrm * nrm);   // line 153
  float a = expf(-H) / (expf(-h0[p]) + 1e-10f);                                // line 156
  a = fminf(fmaxf(a, 0.f), 1.f);                                               // line 157
  const bool mv = acc[p] < a;                                                  // line 159
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    z_out[e] = mv ? z_cur[e] : z_prev[e];                                      // line 162
  }
  if (h1) h1[p] = H;
  if (alpha_out) alpha_out[p] = a;
  if (moves) moves[p] = mv ? 1.f : 0.f;
}

// The stages between the first and the last leapfrog step are purely element-wise (lines 141-148, then
// 132-138 of the next step): one float4 of the flat [N*d] arrays per thread, perfectly coalesced 128-bit
// accesses (the one-thread-per-chain kernel above walked a 64-byte row with scalar loads: 18 % of HBM peak).
__global__ void hmc_step_mid_vec4_kernel(const float4* __restrict__ diag_g, const float4* __restrict__ gex,
                                         int64_t total4, float eps, float lambda, float T2, int mode, float scale,
                                         float4* __restrict__ rho_half, float4* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const float half_eps = eps / 2.f;
  const float4 dg = __ldg(diag_g + i);
  const float4 ge = gex ? __ldg(gex + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 rh = rho_half[i], zc = z[i];
  const float dgv[4] = {dg.x, dg.y, dg.z, dg.w}, gev[4] = {ge.x, ge.y, ge.z, ge.w};
  float rhv[4] = {rh.x, rh.y, rh.z, rh.w}, zv[4] = {zc.x, zc.y, zc.z, zc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float g = neg_grad(mode, dgv[e], gev[e], lambda, T2);
    const float rho = scale * (rhv[e] - half_eps * g);
    rhv[e] = rho - half_eps * g;
    zv[e] = zv[e] + eps * rhv[e];
  }
  rho_half[i] = make_float4(rhv[0], rhv[1], rhv[2], rhv[3]);
  z[i] = make_float4(zv[0], zv[1], zv[2], zv[3]);
}

__global__ void axpy_grad_modular_kernel(float* __restrict__ z, const float* __restrict__ diag_g,
                                         int64_t total, float step, float lambda, float T2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  // z + step * (-(grad_func(z)))   ref hmc_sampler.py:247-251
  z[e] = z[e] + step * (-((1.f - lambda * diag_g[e]) / T2));
}

int launch_hmc_begin(const float* z, const float* gamma, const float* diag_g, const float* logabsdet,
                     const float* sign, const float* grad_exact, int64_t n, int d, float b0,
                     float eps, float lambda, float T2, int grad_mode, float* rho_half, float* z_new,
                     float* h0, cudaStream_t s) {
  if (n == 0) return 0;
  hmc_begin_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(z, gamma, diag_g, logabsdet, sign,
                                                               grad_exact, n, d, b0, eps, lambda, T2,
                                                               grad_mode, rho_half, z_new, h0);
  RLVAE_LAUNCH_OK();
  return 0;
}

int launch_hmc_step(const float* diag_g, const float* logabsdet, const float* sign,
                    const float* grad_exact, int64_t n, int d, float eps, float lambda, float T2,
                    int grad_mode, float scale, int last, float* rho_half, float* z_cur,
                    const float* z_prev, const float* acc, const float* h0, float* h1, float* alpha,
                    float* moves, float* z_out, cudaStream_t s) {
  if (n == 0) return 0;
  const int64_t total = n * d;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!last && (total & 3) == 0 && al16(diag_g) && al16(rho_half) && al16(z_cur) && (grad_exact == nullptr || al16(grad_exact))) {
    const int64_t total4 = total / 4;
    hmc_step_mid_vec4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const float4*>(diag_g), reinterpret_cast<const float4*>(grad_exact), total4, eps, lambda, T2,
        grad_mode, scale, reinterpret_cast<float4*>(rho_half), reinterpret_cast<float4*>(z_cur));
    RLVAE_LAUNCH_OK();
    return 0;
  }
  hmc_step_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(
      diag_g, logabsdet, sign, grad_exact, n, d, eps, lambda, T2, grad_mode, scale, last, rho_half,
      z_cur, z_prev, acc, h0, h1, alpha, moves, z_out);
  RLVAE_LAUNCH_OK();
  return 0;
}

int launch_axpy_grad_modular(float* z, const float* diag_g, int64_t n, int d, float step,
                             float lambda, float T2, cudaStream_t s) {
  const int64_t total = n * d;
  if (total == 0) return 0;
  axpy_grad_modular_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, diag_g, total, step,
                                                                           lambda, T2);
  RLVAE_LAUNCH_OK();
  return 0;
}

}  // namespace rlvae
