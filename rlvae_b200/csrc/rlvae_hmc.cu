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

// ---- pythae RHVAESampler.hmc_sampling (ref src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:98-148) ----
// log_pi = log(sqrt(det G^{-1}) + 1e-10) (:157-158), alpha = exp(-H)/exp(-H0) un-clamped (:139), the proposal mixed as
// z*m + (1-m)*z0 (:143).  Every product / sum below is rounded separately (the reference is eager tensor arithmetic).
__device__ __forceinline__ float pythae_log_pi(float lad, float sgn) {
  return logf(sqrtf(sgn * expf(lad)) + 1e-10f);
}

// start of one MCMC iteration for every chain: rho = gamma / b0 (:107), H0 (:108), first half-step + position update
// of leapfrog step 0 (:113-116).  from_eval: (lp0, g0) are first taken from a fresh evaluation (lad, sgn, grad) at z.
__global__ void pythae_hmc_begin_kernel(int64_t n, int d, float eps, float b0, int from_eval, const float* __restrict__ lad,
                                        const float* __restrict__ sgn, const float* __restrict__ grad,
                                        const float* __restrict__ gamma, float* __restrict__ z, float* __restrict__ z0,
                                        float* __restrict__ rho_half, float* __restrict__ g0, float* __restrict__ lp0,
                                        float* __restrict__ h0, float* __restrict__ rec_h0) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float half_eps = eps / 2.f;
  if (from_eval) lp0[p] = pythae_log_pi(lad[p], sgn[p]);
  float ss = 0.f;
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    if (from_eval) g0[e] = grad[e];
    const float rho = gamma[e] / b0;
    ss = __fadd_rn(ss, __fmul_rn(rho, rho));
    const float rh = __fadd_rn(rho, __fmul_rn(half_eps, g0[e]));      // rho - (eps/2) * (-grad log_pi)
    rho_half[e] = rh;
    const float zc = z[e];
    z0[e] = zc;
    z[e] = __fadd_rn(zc, __fmul_rn(eps, rh));
  }
  const float nrm = sqrtf(ss);
  const float H0 = __fadd_rn(-lp0[p], __fmul_rn(0.5f, __fmul_rn(nrm, nrm)));
  h0[p] = H0;
  if (rec_h0) rec_h0[p] = H0;
}

// after the evaluation at the position of leapfrog step k: second half-step, tempering (:121-131); then either the
// first half of step k + 1 (same gradient) or, on the last step, H, alpha, the accept decision and the bookkeeping of
// (z0, log_pi, gradient) at the position that is kept (:134-146).
__global__ void pythae_hmc_step_kernel(int64_t n, int d, float eps, float scale, int last, const float* __restrict__ lad,
                                       const float* __restrict__ sgn, const float* __restrict__ grad,
                                       const float* __restrict__ acc, float* __restrict__ z,
                                       float* __restrict__ z0, float* __restrict__ rho_half, float* __restrict__ g0,
                                       float* __restrict__ lp0, const float* __restrict__ h0, float* __restrict__ rec_h,
                                       float* __restrict__ rec_alpha, float* __restrict__ rec_moves,
                                       float* __restrict__ z_trace) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float half_eps = eps / 2.f;
  if (!last) {
    for (int j = 0; j < d; ++j) {
      const int64_t e = p * d + j;
      const float g = grad[e];
      const float rho = __fmul_rn(scale, __fadd_rn(rho_half[e], __fmul_rn(half_eps, g)));
      const float rh = __fadd_rn(rho, __fmul_rn(half_eps, g));
      rho_half[e] = rh;
      z[e] = __fadd_rn(z[e], __fmul_rn(eps, rh));
    }
    return;
  }
  float ss = 0.f;
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    const float rho = __fmul_rn(scale, __fadd_rn(rho_half[e], __fmul_rn(half_eps, grad[e])));
    ss = __fadd_rn(ss, __fmul_rn(rho, rho));
  }
  const float nrm = sqrtf(ss);
  const float lp = pythae_log_pi(lad[p], sgn[p]);
  const float H = __fadd_rn(-lp, __fmul_rn(0.5f, __fmul_rn(nrm, nrm)));
  const float a = expf(-H) / expf(-h0[p]);
  const bool mv = acc[p] < a;                      // NaN alpha: stays
  const float m = mv ? 1.f : 0.f;
  bool lost = false;
  for (int j = 0; j < d; ++j) {
    const int64_t e = p * d + j;
    const float zn = __fadd_rn(__fmul_rn(z[e], m), __fmul_rn(1.f - m, z0[e]));
    lost |= !isfinite(zn);
    z[e] = zn;
    z0[e] = zn;
    if (z_trace) z_trace[e] = zn;
    if (mv) g0[e] = grad[e];
  }
  if (mv) lp0[p] = lp;
  if (lost) {          // 0 * inf left NaN in z (as in the reference, which then evaluates log_pi at NaN)
    lp0[p] = NAN;
    for (int j = 0; j < d; ++j) g0[p * d + j] = NAN;
  }
  if (rec_h) rec_h[p] = H;
  if (rec_alpha) rec_alpha[p] = a;
  if (rec_moves) rec_moves[p] = m;
}

int launch_pythae_hmc_begin(int64_t n, int d, float eps, float b0, int from_eval, const float* lad, const float* sgn,
                            const float* grad, const float* gamma, float* z, float* z0, float* rho_half, float* g0,
                            float* lp0, float* h0, float* rec_h0, cudaStream_t s) {
  if (n == 0) return 0;
  pythae_hmc_begin_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, d, eps, b0, from_eval, lad, sgn, grad, gamma, z,
                                                                      z0, rho_half, g0, lp0, h0, rec_h0);
  RLVAE_LAUNCH_OK();
  return 0;
}

int launch_pythae_hmc_step(int64_t n, int d, float eps, float scale, int last, const float* lad, const float* sgn,
                           const float* grad, const float* acc, float* z, float* z0, float* rho_half, float* g0,
                           float* lp0, const float* h0, float* rec_h, float* rec_alpha, float* rec_moves,
                           float* z_trace, cudaStream_t s) {
  if (n == 0) return 0;
  pythae_hmc_step_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, d, eps, scale, last, lad, sgn, grad, acc, z, z0,
                                                                     rho_half, g0, lp0, h0, rec_h, rec_alpha, rec_moves,
                                                                     z_trace);
  RLVAE_LAUNCH_OK();
  return 0;
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
