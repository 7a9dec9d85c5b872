// Internal declarations shared by the rlvae_b200 CUDA translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>

#include "../../include/rlvae_b200.h"

namespace rlvae {

void set_error(const std::string& msg);

#define RLVAE_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::rlvae::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define RLVAE_REQUIRE(cond, msg)                                                         \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::rlvae::set_error(std::string("rlvae_b200: ") + (msg));                           \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

// every kernel launch of the hot path goes through one of these two (rlvae_launch_count reports the total)
void count_launch();
void prof_mark(int slot, cudaStream_t s);   // rlvae_profile_*: event `slot` of the current record (no-op unless profiling)
#define RLVAE_LAUNCH_OK()                    \
  do {                                       \
    ::rlvae::count_launch();                 \
    RLVAE_CUDA_OK(cudaGetLastError());       \
  } while (0)
#define RLVAE_LAUNCH_EX(expr)                \
  do {                                       \
    ::rlvae::count_launch();                 \
    RLVAE_CUDA_OK(expr);                     \
  } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (context),
// and one process may drive several GPUs (the Python wrappers switch with torch.cuda.device(z.device)),
// so the "already done" flag is one bit per device ordinal, not one per process.
#define RLVAE_OPT_IN_SMEM(kern, bytes)                                                          \
  do {                                                                                          \
    static std::atomic<unsigned long long> _done{0};                                            \
    int _dev = 0;                                                                               \
    RLVAE_CUDA_OK(cudaGetDevice(&_dev));                                                        \
    const unsigned long long _bit = 1ull << (_dev & 63);                                        \
    if (!(_done.load(std::memory_order_acquire) & _bit)) {                                      \
      RLVAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                         (int)(bytes)));                                        \
      _done.fetch_or(_bit, std::memory_order_release);                                          \
    }                                                                                           \
  } while (0)

constexpr int kMaxLatentDim = 64;
constexpr int kKPad = 128;  // centroid tables are zero-padded to a multiple of this

}  // namespace rlvae

// Device-resident packed tables (derived caches of the MetricTensor buffers).
struct rlvae_tables {
  int K = 0, d = 0, Kpad = 0;
  float T = 0.f, T2 = 0.f, lambda = 0.f;
  int symmetric = 0;       // every M_k symmetric to 2^-22 of its largest entry (packed tables hold (M + M^T)/2)
  int tensor_capable = 0;  // d == 16 and TMA descriptors built
  int tensor_auto = 0;     // AUTO picks a tensor path (expanded form accurate enough, or exact-distance mode available)
  int expanded_ok = 0;     // accuracy criterion for the expanded-distance form ||z||^2+||c||^2-2z.c holds
  float r2max = 0.f;       // max_k ||c_k||^2

  // natural layouts, zero padded to Kpad rows
  float* c = nullptr;   // [Kpad, d]
  float* cn = nullptr;  // [Kpad]  ||c_k||^2
  float* M = nullptr;   // [Kpad, d*d]

  // tensor path (d == 16)
  float* cstack = nullptr;  // [Kpad, 32] = [tf32_hi(c) | c - hi]
  float* cbias = nullptr;   // [Kpad]  -||c||^2 * log2(e)/T^2  (padding rows: -1e30)
  float* cn_inf = nullptr;  // [Kpad]  ||c||^2 with 3e38 on the padding rows (tensor nearest2)
  float* cmask = nullptr;   // [Kpad]  0 on real rows, -1e30 on padding rows (exact-distance tensor mode)
  float* Mt_hi = nullptr;   // [256, Kpad]  tf32_hi(M) transposed (centroid index contiguous)
  float* Mt_lo = nullptr;   // [256, Kpad]  M - hi
  float* Mn_hi = nullptr;   // [Kpad, 256]  natural, for the gradient pass
  float* Mn_lo = nullptr;   // [Kpad, 256]
  float* ct_hi = nullptr;   // [32, Kpad]  rows 0..15 tf32_hi(c)^T, row 16 = 1, rest 0 (gradient contraction)
  float* ct_lo = nullptr;   // [32, Kpad]  rows 0..15 (c - hi)^T, rest 0
  // symmetric tables only: packed upper triangle (136 -> 144 rows), transposed, hi/lo
  float* Mts_hi = nullptr;  // [144, Kpad]
  float* Mts_lo = nullptr;  // [144, Kpad]
  float* Mns_hi = nullptr;  // [Kpad, 160] packed natural (gradient kernel B operand)
  float* Mns_lo = nullptr;  // [Kpad, 160]
  CUtensorMap tm_cstack, tm_mt_hi, tm_mt_lo, tm_mn_hi, tm_mn_lo, tm_mts_hi, tm_mts_lo;
  // CTA-pair variants: each CTA of a pair fetches half of the B-tile rows (smaller boxes)
  CUtensorMap tm_mt2_hi, tm_mt2_lo, tm_mts2_hi, tm_mts2_lo;
  CUtensorMap tm_mn2_hi, tm_mn2_lo, tm_mns_hi, tm_mns_lo, tm_mns2_hi, tm_mns2_lo;
  CUtensorMap tm_ct_hi, tm_ct_lo, tm_ct2_hi, tm_ct2_lo;
  CUtensorMap tm_ct16_hi, tm_ct16_lo, tm_ct8_hi, tm_ct8_lo;   // boxes of 32 centroids x 16 (pair: 8) rows
  // split-fp16 tables (symmetric, d == 16): fp16(2^e M) and fp16 residual, packed-transposed [144, Kpad]
  void* Mh_hi = nullptr;
  void* Mh_lo = nullptr;
  void* Mnh_hi = nullptr;      // natural [Kpad, 192] (136 packed columns + zeros), gradient kernel B operand
  void* Mnh_lo = nullptr;
  float h16_out_scale = 0.f;   // 2^-(14+e)
  float h16_m_unscale = 0.f;   // 2^-e
  float m_absmax = 0.f;
  CUtensorMap tm_mh_hi, tm_mh_lo, tm_mh2_hi, tm_mh2_lo;
  CUtensorMap tm_mnh_hi, tm_mnh_lo, tm_mnh2_hi, tm_mnh2_lo;
  // d == 64 split-fp16 forward path (rlvae_tc64.cu): centroid rows [Kpad,128] fp16 = [hi | lo] of 2^ec c;
  // Mh_hi / Mh_lo then hold the packed-transposed [2176, Kpad] tables
  void* c16h = nullptr;        // d == 16: [Kpad,64] fp16 = [hi (16) | lo (16) | 0] of 2^ec c (GEMM1 of the fp16 kernels)
  float c16_unscale = 0.f;     // 2^-ec
  int psd_certified = 0;       // d == 16 symmetric: every M_k positive semi-definite (to rounding) and lambda > 0, so
                               //   G^{-1}(z) is positive definite and the Cholesky fallback list stays empty
  int hybrid_ok = 0;           // d == 16 symmetric: the hybrid weight mode has a valid threshold
  float hybrid_bits = 0.f;     //   weights below 2^-hybrid_bits cannot move G^{-1} by more than 1e-6 lambda
  float* cshift = nullptr;     // [d] mean centroid the fp16 GEMM1 operands are centred on (+ scratch; d == 16 or 64)
  float* cbias_h = nullptr;    // [Kpad] -||c - shift||^2 * log2(e)/T^2 (padding rows: -1e30)
  float* ctc_hi = nullptr;     // [16, Kpad] TF32 hi / lo of (c - shift)^T: B operand of the fp16 gradient
  float* ctc_lo = nullptr;     //            kernel's final contraction (tm_ct16_*, tm_ct8_*)
  float c_absmax = 0.f;
  CUtensorMap tm_c16h;
  void* c64h = nullptr;
  float c64_unscale = 0.f;     // 2^-ec
  float r2mean_centred = 0.f;  // mean ||c - cshift||^2 (d == 64: set by tc_build_h64_tables)
  CUtensorMap tm_c64, tm_c64_2;   // boxes of 32 centroids x 32 (pair: 16) rows
  // d == 64 gradient kernel: (c - shift)^T scaled by 2^ec, split fp16 [64, Kpad]; Mnh_hi / Mnh_lo then hold the
  // natural [Kpad, 2176] tables
  void* ct64_hi = nullptr;
  void* ct64_lo = nullptr;
  CUtensorMap tm_ct64_hi, tm_ct64_lo, tm_ct64_2_hi, tm_ct64_2_lo;
  // pythae-variant gradient (A8) on the tensor path (d == 16, symmetric): b_k = sym(M_k) (c_k - shift), TF32 hi / lo,
  // transposed [16, Kpad] -- the B operand of the gradient kernel's final contraction in its unit-weight mode
  // (sum_k w_k b_k); m_fro_rms = (sum_k ||M_k||_F) / sqrt(K) enters the error bound that sends points to the
  // per-centroid kernel
  float* bt_hi = nullptr;
  float* bt_lo = nullptr;
  float m_fro_rms = 0.f;
  CUtensorMap tm_bt16_hi, tm_bt16_lo, tm_bt8_hi, tm_bt8_lo;
};

namespace rlvae {

// ---- launchers (each returns 0 / error code, asynchronous on `s`) -----------------------------
int launch_inverse_metric_direct(const rlvae_tables* t, const float* z, int64_t n, float* ginv,
                                 cudaStream_t s);
int launch_metric_grad_direct(const rlvae_tables* t, const float* z, const float* u, int64_t n,
                              float scale, float* out, cudaStream_t s);
// variant C (pythae) with the difference c_k - z formed per centroid, like the reference (rlvae_direct.cu)
// partial (optional): n * kPythaeMaxSplits * d floats -- short batches (n <= kPythaeSplitBatch) then split the
// centroids over several CTAs per point
constexpr int64_t kPythaeSplitBatch = 2048;
constexpr int kPythaeMaxSplits = 16;
int launch_pythae_exact(const rlvae_tables* t, const float* z, const float* g, int g_is_packed, const int* list,
                        const int* count, int64_t n, float* out, float* partial, cudaStream_t s);
int launch_batched_inverse(const float* a, int64_t n, int d, float* inv, float* logabsdet,
                           float* sign, float* diag_inv, int transpose_inv, cudaStream_t s);
// d == 16 only: `a` is the packed symmetric layout [N,144] written by the symmetric tensor kernel
// 64 x 64 symmetric positive definite batch (no pivoting; failures re-run through the pivoting kernel):
// inv optional, logabsdet = lad_scale * log det, fail_ws = 1 + n ints
int launch_spd64(const float* a, int64_t n, float* inv, float* logabsdet, float lad_scale, int* fail_ws,
                 cudaStream_t s, float* sign = nullptr, float* diag_inv = nullptr);
int launch_batched_inverse_packed16(const float* a_packed, int64_t n, float* inv, float* logabsdet,
                                    float* sign, float* diag_inv, int transpose_inv, cudaStream_t s);
int launch_unpack_sym16(const float* a_packed, int64_t n, float* full, cudaStream_t s);
// d == 16, symmetric packed input: per-thread Cholesky with a pivoting fallback for the matrices that
// are not positive definite.  fail_ws: 1 + n ints.  logabsdet receives lad_scale * log|det A|.
int launch_sym16_inverse(const float* a_packed, int64_t n, float* g_packed, float* logabsdet,
                         float lad_scale, float* sign, float* diag_g, int* fail_ws, cudaStream_t s);
// eigenvalues (ascending) of symmetric 16x16 matrices: a = packed [N,144] or full [N,16,16] (upper triangle read)
int launch_sym16_eigvalsh(const float* a, int64_t n, int packed, float* eig, cudaStream_t s);
int launch_sym16_fallback(const float* a_packed, int64_t n, float* g_packed, float* logabsdet,
                          float lad_scale, float* sign, float* diag_g, int* fail_ws, cudaStream_t s,
                          float* g_full = nullptr, const rlvae_tables* recompute_from = nullptr,
                          const float* z = nullptr);
// split-fp16 tensor kernel (rlvae_tc16.cu): forward + fused per-point Cholesky outputs
int tc_build_h16_descriptors(rlvae_tables* t);
int h16_mode(const rlvae_tables* t);   // 0 expanded, 1 exact differences, 2 hybrid
// d == 64 (symmetric tables): tables + forward through a packed [N,2176] scratch
int tc_build_h64_tables(rlvae_tables* t, cudaStream_t s);
int launch_inverse_metric_h64(const rlvae_tables* t, const float* z, int64_t n, float* ginv, float* packed_scratch,
                              cudaStream_t s);
constexpr int kSym64Cols = 2176;
// d == 64 gradient on the tensor cores (column-tiled like the forward kernel; partial tiles + reduction)
bool metric_grad_h64_available(const rlvae_tables* t);
int64_t metric_grad_h64_scratch_floats(int64_t n);
int launch_metric_grad_h64(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale, float* out,
                           float* scratch, cudaStream_t s);
int launch_nearest2_tc(const rlvae_tables* t, const float* mu, int64_t n, int64_t* idx, float* dist, cudaStream_t s);
// u_packed: 0 full [N,256] U, 1 packed symmetric [N,144] U, 2 unit mode (u unused): out = scale * sum_k w_k b_k
// with the pythae table b (tm_bt*) in place of the centroids
int launch_metric_grad_h16(const rlvae_tables* t, const float* z, const float* u, int64_t n, float scale,
                           float* out, cudaStream_t s, int u_packed);
// a_full (optional): the expanded [N,16,16] G^{-1}, written by the same kernel
// a_packed_wanted == 0: a_packed is only the fallback's scratch -- for certified tables the kernel then
// skips the 576 B/point store and the (normally empty) fallback list is recomputed instead
int launch_inverse_metric_h16(const rlvae_tables* t, const float* z, int64_t n, float* a_packed,
                              float* g_packed, float* logabsdet, float lad_scale, float* sign, float* diag_g,
                              int* fail_ws, cudaStream_t s, float* a_full = nullptr, float* g_full = nullptr,
                              int a_packed_wanted = 1, float* s_diag = nullptr);
// single-launch HMC trajectory (variant-A drift): n_iters MCMC iterations, chain state on chip.
// scales_dev [n_iters * n_lf] (device) or, for n_iters == 1 and n_lf <= 64, h_scales (host).
int h16_hmc_available(const rlvae_tables* t);
int launch_hmc_trajectory_h16(const rlvae_tables* t, float* z, const float* gamma, const float* acc, int64_t n,
                              int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* scales_dev,
                              const float* h_scales, float* h0, float* h1, float* alpha, float* moves,
                              float* z_trace, int* fail_count, cudaStream_t s);
constexpr int kSymCols = 144;
constexpr int kSymNatCols = 160;
// pythae RHVAESampler.hmc_sampling stages (rlvae_hmc.cu)
int launch_pythae_hmc_begin(int64_t n, int d, float eps, float b0, int from_eval, const float* lad, const float* sgn,
                            const float* grad, const float* gamma, float* z, float* z0, float* rho_half, float* g0,
                            float* lp0, float* h0, float* rec_h0, cudaStream_t s);
int launch_pythae_hmc_step(int64_t n, int d, float eps, float scale, int last, const float* lad, const float* sgn,
                           const float* grad, const float* acc, float* z, float* z0, float* rho_half, float* g0,
                           float* lp0, const float* h0, float* rec_h, float* rec_alpha, float* rec_moves,
                           float* z_trace, cudaStream_t s);
int launch_chol_apply(const float* a, const float* eps, int64_t n, int d, float jitter, float* out,
                      int32_t* status, cudaStream_t s);
int launch_nearest2(const rlvae_tables* t, const float* mu, int64_t n, int64_t* idx, float* dist,
                    cudaStream_t s);

// tensor path (rlvae_tc.cu)
int tc_build_descriptors(rlvae_tables* t);
int tc_build_ct_centred_descriptors(rlvae_tables* t);
int launch_inverse_metric_tc(const rlvae_tables* t, const float* z, int64_t n, float* ginv,
                             cudaStream_t s);
// symmetric tables: packed [N,144] result (lambda already on the packed diagonal)
int launch_inverse_metric_tc_sym(const rlvae_tables* t, const float* z, int64_t n, float* packed,
                                 cudaStream_t s);
int tc_build_sym_descriptors(rlvae_tables* t);
// u_packed != 0 (symmetric tables only): u is a SYMMETRIC matrix in the packed [N,144] layout
int launch_metric_grad_tc(const rlvae_tables* t, const float* z, const float* u, int64_t n,
                          float scale, float* out, cudaStream_t s, int u_packed = 0);

// metric construction (rlvae_build.cu)
int launch_local_covariance(const float* mus, int64_t n, const float* centroids, int k, int d, float temperature,
                            float* cov, cudaStream_t s);

// HMC elementwise stages (rlvae_hmc.cu)
struct HmcBeginArgs;
int launch_hmc_begin(const float* z, const float* gamma, const float* diag_g, const float* logabsdet,
                     const float* sign, const float* grad_exact, int64_t n, int d, float inv_b0,
                     float eps, float lambda, float T2, int grad_mode, float* rho_half, float* z_new,
                     float* h0, cudaStream_t s);
int launch_hmc_step(const float* diag_g, const float* logabsdet, const float* sign,
                    const float* grad_exact, int64_t n, int d, float eps, float lambda, float T2,
                    int grad_mode, float scale, int last, float* rho_half, float* z_cur,
                    const float* z_prev, const float* acc, const float* h0, float* h1, float* alpha,
                    float* moves, float* z_out, cudaStream_t s);
int launch_axpy_grad_modular(float* z, const float* diag_g, int64_t n, int d, float step,
                             float lambda, float T2, cudaStream_t s);

}  // namespace rlvae
