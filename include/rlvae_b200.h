/*
 * rlvae_b200 C ABI  --  the drop-in boundary for RlVAE's metric-evaluation and
 * sampling hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (antoinelfg/RlVAE) has no FFI: its boundary is the Python
 * interface of MetricTensor / the samplers.  This header is
 * what a binding for that interface calls; rlvae_b200/_capi.py is the ctypes
 * binding we ship, INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name starts with h_;
 *     tensors are fp32, row-major, contiguous;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     every call is asynchronous on that stream and allocates nothing
 *     (callers pass workspaces; only rlvae_tables_create allocates);
 *   - return value 0 = ok, non-zero = error, text in rlvae_last_error()
 *     (thread-local);
 *   - a handle is thread-compatible, not thread-safe.
 *
 * Citations "ref:" are file:line in the reference checkout.
 */
#ifndef RLVAE_B200_H_
#define RLVAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rlvae_tables rlvae_tables_t;

/* which implementation evaluates the weighted sum / gradient contraction */
enum {
  RLVAE_PATH_AUTO   = 0, /* tensor path when eligible (d==16, accuracy criterion), else direct */
  RLVAE_PATH_DIRECT = 1, /* fp32 CUDA-core kernel, direct ||z-c||^2 differences (any d <= 64)   */
  RLVAE_PATH_TENSOR = 2  /* tcgen05/TMEM kernels: split-fp16 (symmetric tables, d == 16 or 64) or 3xTF32
                            (any d == 16 table); error if no tensor kernel exists for these tables */
};

/* log_pi / gradient flavours of the HMC sampler */
enum {
  RLVAE_GRAD_MODULAR = 0, /* ref: src/models/samplers/hmc_sampler.py:33-68  == (1 - lambda*G_ii)/T^2 */
  RLVAE_GRAD_EXACT   = 1, /* grad_z 1/2 log det G^{-1}  (autograd of hmc_sampler.py:26-30)           */
  RLVAE_HMC_NO_FUSION = 256 /* flag, OR-ed into grad_mode: per-step launches instead of the fused trajectory */
};

const char* rlvae_last_error(void);
int         rlvae_abi_version(void);
/* number of CUDA kernels this library has launched since load (or since the last reset != 0);
 * bench.py's gpu_launches */
long long   rlvae_launch_count(int reset);

/* Measurement aid (bench.py's roofline): while profiling is on, every rlvae_metric_eval call on the
 * packed tensor path (d == 16, symmetric tables) records CUDA events on its stream around the
 * fused forward launch, the fallback pass and the gradient launch.  rlvae_profile_read gives
 * ms[0] = forward kernel, ms[1] = fallback pass, ms[2] = gradient kernel of call `record`
 * (synchronises on that call's last event).  Not thread-safe.                                  */
int rlvae_profile_begin(int max_records);
int rlvae_profile_count(void);
int rlvae_profile_read(int record, float ms[3]);
int rlvae_profile_end(void);

/* ---- tables: replaces MetricTensor.load_pretrained's buffers --------------------------------
 * ref: src/models/components/metric_tensor.py:59-96 (centroids [K,d], metric_matrices [K,d,d],
 * temperature, regularization).  Packs device-side derived copies: zero-padded tables,
 * ||c_k||^2, TF32 hi/lo splits, the transposed M table and its TMA descriptors.            */
int rlvae_tables_create(rlvae_tables_t** out, const float* centroids, const float* matrices,
                        int n_centroids, int latent_dim, float temperature, float regularization,
                        void* stream);
int rlvae_tables_destroy(rlvae_tables_t* t);
/* info[0]=K, [1]=d, [2]=K padded, [3]=1 if every M_k is symmetric (to 2^-22 of its largest entry),
 * [4]=1 if the tensor path exists for this d, [5]=1 if AUTO would pick the tensor path,
 * [6]=1 if the expanded-distance form is accurate enough for these tables (other tensor kernels
 * are not used by AUTO when it is not), [7]=how the d = 16 symmetric kernels form the weights:
 * 0 expanded form, 1 exact differences, 2 hybrid (expanded form + exact refinement of the weights
 * that can matter), [8]=1 if every M_k was certified positive semi-definite and
 * lambda > 0 (G^{-1}(z) is then positive definite everywhere: the fused kernels skip the packed
 * G^{-1} store that only feeds the pivoting fallback, and the single-launch HMC trajectory kernel is
 * eligible), [9..11] reserved (0) */
int rlvae_tables_info(const rlvae_tables_t* t, int64_t info[12]);

/* ---- A2: G^{-1}(z) = sum_k M_k exp(-||z-c_k||^2/T^2) + lambda I ------------------------------
 * ref: src/models/components/metric_tensor.py:98-137.   z [N,d] -> ginv [N,d,d]
 * `work` (optional, rlvae_inverse_metric_workspace(n,d) bytes) lets symmetric tables use the
 * packed 136-column tensor kernel; with work == NULL the dense 256-column kernel runs.        */
int64_t rlvae_inverse_metric_workspace(int64_t n, int d);   /* bytes */
int rlvae_inverse_metric(const rlvae_tables_t* t, const float* z, int64_t n, float* ginv,
                         void* work, int path, void* stream);
/* Same quantity in the packed symmetric layout the tensor kernel produces for symmetric tables
 * (d == 16): packed [N,144], entry 16 i - i (i-1)/2 + (j - i) holds G^{-1}_ij for i <= j (lambda
 * included), entries 136..143 are zero.  Error unless the tables are symmetric and d == 16.    */
int rlvae_inverse_metric_packed(const rlvae_tables_t* t, const float* z, int64_t n, float* packed,
                                void* stream);

/* ---- A3/A4/A5: batched d x d inverse, log|det|, sign, diagonal of the inverse ----------------
 * ref: torch.linalg.inv / slogdet / det at metric_tensor.py:152,175 and hmc_sampler.py:28.
 * a [N,d,d] -> inv [N,d,d], logabsdet [N], sign [N], diag_inv [N,d]; any output may be NULL.
 * flags bit 0: store the inverse transposed (A^{-T}).  A zero pivot gives sign 0, logabsdet
 * -inf and a non-finite inverse (the caller decides whether to retry, cf. metric_tensor.py:153-158).
 * Partial-pivoting Gauss-Jordan, one matrix per d lanes.                                      */
int rlvae_batched_inverse(const float* a, int64_t n, int d, float* inv, float* logabsdet,
                          float* sign, float* diag_inv, int flags, void* stream);

/* ---- K4: (scale) * sum_k w_k <U, M_k> (c_k - z) ----------------------------------------------
 * u [N,d,d].  With U = G and scale = -2/T^2 this is grad_z log det G (north_star);
 * with U = dL/dG^{-1} and scale = 2/T^2 it is the autograd backward of A2.
 * out [N,d].                                                                                  */
int rlvae_metric_grad(const rlvae_tables_t* t, const float* z, const float* u, int64_t n,
                      float scale, float* out, int path, void* stream);
/* The same with a workspace of rlvae_metric_grad_workspace(n, d) bytes (0 for d != 64): d = 64 symmetric
 * tables then run the column-tiled tcgen05 gradient kernel (partial results per 128-column tile + a
 * fixed-order reduction); without a workspace d = 64 uses the CUDA-core kernel.                 */
int64_t rlvae_metric_grad_workspace(int64_t n, int d);   /* bytes */
int rlvae_metric_grad_ws(const rlvae_tables_t* t, const float* z, const float* u, int64_t n,
                         float scale, float* out, void* work, int path, void* stream);

/* ---- variant C (pythae): (1/T^2) G^T sum_k w_k M_k^T (c_k - z) -------------------------------
 * ref: src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:160-187.
 * g [N,d,d] (the metric at z) -> out [N,d].  `path` as everywhere: long batches (> 2048 points) on symmetric
 * d == 16 tables run as two table contractions on the tensor cores (packed G^{-1} from the forward kernel,
 * sum_k w_k M_k c_k from the gradient kernel's unit-weight mode) with an error bound per row; flagged rows, short
 * batches and every other table use the CUDA-core kernel that forms c_k - z per centroid like the reference.
 * `work` must hold rlvae_metric_grad_pythae_workspace(n, d) bytes. */
int64_t rlvae_metric_grad_pythae_workspace(int64_t n, int d);   /* bytes */
int rlvae_metric_grad_pythae(const rlvae_tables_t* t, const float* z, const float* g, int64_t n,
                             float* out, void* work, int path, void* stream);

/* The same gradient together with log|det G^{-1}(z)| and its sign (the log_pi of ref :150-158 is
 * log(sqrt(sign * exp(logabsdet)) + 1e-10)) in one call -- what one leapfrog step of RHVAESampler.hmc_sampling
 * (ref :98-148) consumes.  grad [N,d]; logabsdet / sign [N] or NULL.  `work`: rlvae_pythae_eval_workspace(n, d). */
int64_t rlvae_pythae_eval_workspace(int64_t n, int d);          /* bytes */
int rlvae_pythae_eval(const rlvae_tables_t* t, const float* z, int64_t n, float* grad, float* logabsdet,
                      float* sign, void* work, int path, void* stream);

/* The whole loop of RHVAESampler.hmc_sampling (ref pythae rhvae_sampler.py:98-148; what
 * OfficialRHVAESampler.sample_prior runs, ref src/models/samplers/rhvae_sampler.py:169-191): n_iters MCMC iterations of
 * n_lf leapfrog steps, one rlvae_pythae_eval per step, momentum / position / accept arithmetic in two small kernels.
 * z [N,d] in: start, out: final positions; gammas [n_iters,N,d] (:107), accs [n_iters,N] (:141);
 * h_scales (HOST) [n_iters*n_lf] = beta_old/beta_new per leapfrog step (the tempering state is never reset, :104);
 * h0 / h1 / alpha / moves [n_iters,N] and z_trace [n_iters,N,d]: optional records.  alpha is not clamped (:139). */
int64_t rlvae_pythae_hmc_workspace(int64_t n, int d);           /* bytes */
int rlvae_pythae_hmc_run(const rlvae_tables_t* t, float* z, const float* gammas, const float* accs, int64_t n,
                         int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* h_scales, float* h0,
                         float* h1, float* alpha, float* moves, float* z_trace, void* work, int path, void* stream);

/* ---- fused metric evaluation -----------------------------------------------------------------
 * z -> any subset of { ginv [N,d,d], g [N,d,d], logdet_g [N] (= log|det G|, ref
 * metric_tensor.py:162-182), grad_logdet_g [N,d] (= grad_z log det G) }.  NULL outputs are
 * skipped.  `work` must hold rlvae_metric_eval_workspace(n, d) bytes.                         */
int64_t rlvae_metric_eval_workspace(int64_t n, int d);   /* bytes */
int rlvae_metric_eval(const rlvae_tables_t* t, const float* z, int64_t n, float* ginv, float* g,
                      float* logdet_g, float* grad_logdet_g, void* work, int path, void* stream);

/* ---- A19: spectrum of the metric for the per-flow-step analysis consumers ---------------------
 * ref: src/visualizations/flow_analysis.py:104-126, src/visualizations/manifold.py:79-101,
 * src/models/modular_rlvae.py:434-457 (torch.linalg.eigvals / det of G^{-1}(z_t), G(z_t) per step).
 * rlvae_sym_eigvalsh: eigenvalues (ascending, like torch.linalg.eigvalsh) of n symmetric d x d
 * matrices, d == 16: a = [N,16,16] (upper triangle read) or, packed != 0, the packed [N,144]
 * layout of rlvae_inverse_metric_packed.  eig [N,16].  Per-thread cyclic Jacobi in registers.
 * rlvae_metric_spectrum: z -> eig(G^{-1}(z)) [N,16] ascending (eig(G) = 1/eig(G^{-1})) and,
 * optionally, log|det G| [N], in one pass over the tables (symmetric tables, d == 16);
 * work: rlvae_metric_eval_workspace(n, d) bytes.                                              */
int rlvae_sym_eigvalsh(const float* a, int64_t n, int d, int packed, float* eig, void* stream);
int rlvae_metric_spectrum(const rlvae_tables_t* t, const float* z, int64_t n, float* eig_ginv,
                          float* logdet_g, void* work, int path, void* stream);

/* ---- metric construction (the step before the path) ----------------------
 * ref: scripts/train_and_extract_vanilla_vae.py:199-221.  For every centroid c_i the covariance of
 * the encoded latents under the normalised Gaussian weights exp(-||mu_n - c_i||^2 / T^2):
 * latents [N,d], centroids [K,d] -> cov [K,d,d].  (M_i = cov_i + reg I and the minimum-eigenvalue
 * lift of lines 213-218 are composed by the host mirror, rlvae_b200.metric_builder.)            */
int rlvae_local_covariance(const float* latents, int64_t n, const float* centroids, int n_centroids,
                           int latent_dim, float temperature, float* cov, void* stream);

/* ---- A11: one MCMC iteration of RiemannianHMCSampler.sample ----------------------------------
 * ref: src/models/samplers/hmc_sampler.py:120-163.  In/out: z [N,d] (chain state, replaced by
 * the accepted state).  gamma [N,d] and acc [N] are the random draws of lines 122 and 158.
 * h_scales [n_lf] (HOST) = beta_sqrt_old/beta_sqrt per leapfrog step (lines 147-149).
 * Optional outputs: h0 [N], h1 [N], alpha [N], moves [N] (1.0 = accepted).
 * One metric evaluation per leapfrog step (+1 for H0).  work: rlvae_hmc_workspace(n,d) bytes. */
int64_t rlvae_hmc_workspace(int64_t n, int d);
int rlvae_hmc_iteration(const rlvae_tables_t* t, float* z, const float* gamma, const float* acc,
                        int64_t n, int n_lf, float eps_lf, float beta_zero_sqrt,
                        const float* h_scales, int grad_mode, float* h0, float* h1, float* alpha,
                        float* moves, void* work, int path, void* stream);
/* Fused trajectory (north_star "one leapfrog-step kernel advances many HMC chains per launch"): when
 * rlvae_hmc_fused_available(t, grad_mode, path) != 0 (d == 16, symmetric tables certified positive
 * semi-definite with lambda > 0, RLVAE_GRAD_MODULAR, n_lf <= 64 for rlvae_hmc_iteration) the whole
 * iteration -- all n_lf + 1 metric evaluations, the momentum / position updates of lines 127-148 and
 * the accept step of lines 153-162 -- is ONE kernel launch: a CTA pair keeps its 256 chains' z and
 * rho_half on chip for the entire trajectory (rlvae_hmc_run: n_iters * n_lf + 1 metric evaluations in
 * all -- the metric at the state an iteration starts from is the one its predecessor ended with).  The int32 at byte offset
 * rlvae_hmc_workspace(n,d) - 256 of `work` then counts chain-steps whose G^{-1} lost positive
 * definiteness to rounding (cond(G^{-1}) beyond ~1e5; log_pi takes its clamp value there); callers
 * that care re-run with RLVAE_HMC_FUSED=0 in the environment (per-step launches + pivoting fallback).
 *
 * rlvae_hmc_run: n_iters MCMC iterations (the loop of lines 120-163) in one call -- one launch on the
 * fused path.  gammas [n_iters,N,d], accs [n_iters,N], h_scales [n_iters*n_lf] (HOST; the tempering
 * state carries over between iterations like the reference's beta_sqrt_old, line 116).  Optional
 * outputs h0/h1/alpha/moves [n_iters,N] and z_trace [n_iters,N,d] (the state after every iteration).
 * work: rlvae_hmc_run_workspace(n, d, n_iters, n_lf) bytes.                                   */
int     rlvae_hmc_fused_available(const rlvae_tables_t* t, int grad_mode, int path);
int64_t rlvae_hmc_run_workspace(int64_t n, int d, int n_iters, int n_lf);
int rlvae_hmc_run(const rlvae_tables_t* t, float* z, const float* gammas, const float* accs, int64_t n,
                  int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* h_scales,
                  int grad_mode, float* h0, float* h1, float* alpha, float* moves, float* z_trace,
                  void* work, int path, void* stream);

/* ---- A13: z <- z + step * (-grad_func(z)), n_steps times (variant A) -------------------------
 * ref: src/models/samplers/hmc_sampler.py:242-257.  In/out z [N,d].                           */
int rlvae_hmc_refine(const rlvae_tables_t* t, float* z, int64_t n, int n_steps, float step_size,
                     void* work, int path, void* stream);

/* ---- A14/A15: two nearest centroids by Euclidean distance ------------------------------------
 * ref: src/models/samplers/riemannian_sampler.py:58-67,125-131.
 * mu [N,d] -> idx [N,2] (int64), dist [N,2]                                                   */
int rlvae_nearest2(const rlvae_tables_t* t, const float* mu, int64_t n, int64_t* idx, float* dist,
                   void* stream);

/* ---- A14-A17: out = cholesky(A + jitter I) @ eps ----------------------------------------------
 * ref: src/models/samplers/riemannian_sampler.py:83-84,156-157,200-201,271-272.
 * a [N,d,d], eps [N,d] -> out [N,d]; status [N] (int32, 0 ok / 1 not positive definite) or NULL */
int rlvae_chol_apply(const float* a, const float* eps, int64_t n, int d, float jitter, float* out,
                     int32_t* status, void* stream);

/* ---- A20: (z1-z2)^T G((z1+z2)/2) (z1-z2) is composed in the host mirror from the calls above. */

#ifdef __cplusplus
}
#endif
#endif /* RLVAE_B200_H_ */
