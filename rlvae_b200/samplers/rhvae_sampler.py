"""Pythae-variant manifold HMC.

Host-side mirror of the sampling loop ``OfficialRHVAESampler.sample_prior`` runs in the
reference: ``RHVAESampler.hmc_sampling`` (ref
``src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:98-148``) with
``log_pi = log(sqrt(det G^{-1}) + 1e-10)`` (:157-158) and the gradient
``(1/T^2) G^T sum_k w_k M_k^T (c_k - z)`` (:160-187, ``rlvae_metric_grad_pythae``).  Reference
behaviour that is kept on purpose: chains start at randomly chosen centroids, alpha =
exp(-H)/exp(-H0) is neither clamped nor regularised, the tempering state is not reset between MCMC
iterations, everything runs under ``no_grad``.  The wrapper class of the reference
(``src/models/samplers/rhvae_sampler.py``) builds a full pythae ``RHVAE`` model around the
encoder/decoder; that model object is out of scope, the sampling arithmetic is here.

Per leapfrog step: ONE call of ``rlvae_pythae_eval`` (forward kernel -> packed G^{-1}, packed G, log det;
the gradient kernel's unit-weight mode -> sum_k w_k M_k c_k; one finish kernel) -- the reference evaluates the
same gradient at the end of step k and at the start of step k + 1 (:110-131, z does not move in between) and
log_pi again at the accepted / rejected position (:108, :134); here the values are carried over (selected per
chain by the accept mask).  The momentum / position updates are element-wise torch ops on the device.

``OfficialRHVAESampler`` mirrors the reference's wrapper class (ref src/models/samplers/rhvae_sampler.py:
13-255) -- same name, methods and quirks (temperature hard-coded to 0.1 at :62/:80, prior batches of at
most 32 at :186, 100 x 15 leapfrog steps of 0.03 at :103-108) -- around that loop.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .. import _capi
from .base_sampler import BaseRiemannianSampler, kernel_path_for, tables_for


class RHVAEStyleHMCSampler(BaseRiemannianSampler):
    def __init__(self, model, mcmc_steps_nbr: int = 100, n_lf: int = 15, eps_lf: float = 0.03,
                 beta_zero: float = 1.0):
        super().__init__(model)
        self.mcmc_steps_nbr = int(mcmc_steps_nbr)
        self.n_lf = int(n_lf)
        self.eps_lf = float(eps_lf)
        self.beta_zero_sqrt = float(beta_zero) ** 0.5

    # ------------------------------------------------------------------ the two callables (ref :150-187)
    def _logp_grad(self, z):
        """(log_pi [N], grad log_pi [N,d]) at z from one fused evaluation."""
        tab = tables_for(self.model)
        grad, lad, sgn = _capi.pythae_eval(tab, z, path=kernel_path_for(self.model))
        det = sgn * torch.exp(lad)
        return torch.log(torch.sqrt(det) + 1e-10), grad

    def log_sqrt_det_G_inv(self, z: torch.Tensor) -> torch.Tensor:
        """log(sqrt(det G^{-1}(z)) + 1e-10)  (ref :157-158; a negative determinant gives NaN there too)."""
        return self._logp_grad(z.contiguous().float())[0]

    def grad_log_sqrt_det_G_inv(self, z: torch.Tensor) -> torch.Tensor:
        """[N,d]  (the reference returns [N,d,1] and reshapes, ref :118-120)."""
        return self._logp_grad(z.contiguous().float())[1]

    @staticmethod
    def tempering(k: int, K: int, beta_zero_sqrt: float) -> float:
        beta_k = ((1 - 1 / beta_zero_sqrt) * (k / K) ** 2) + 1 / beta_zero_sqrt
        return 1 / beta_k

    # ------------------------------------------------------------------ the sampling loop (ref :98-148)
    def hmc_sampling_with_streams(self, idx0: Optional[torch.Tensor], gammas: torch.Tensor, accs: torch.Tensor,
                                  record: Optional[dict] = None, z_start: Optional[torch.Tensor] = None,
                                  state: Optional[dict] = None) -> torch.Tensor:
        """The loop with its random draws injected: ``idx0 [n]`` (line 100), ``gammas [steps,n,d]``
        (line 107), ``accs [steps,n]`` (line 141).  ``z_start`` / ``state`` continue a chain whose draws
        arrive in several slabs (the tempering state is never reset, line 104).

        One library call (``rlvae_pythae_hmc_run``): per leapfrog step one fused evaluation and one update kernel."""
        with torch.no_grad():
            z = (self.model.centroids_tens[idx0] if z_start is None else z_start).float().contiguous().clone()
            iters = int(gammas.shape[0])
            b0 = self.beta_zero_sqrt
            beta_old = b0 if state is None else state['beta_old']
            scales = []
            for _ in range(iters):
                for k in range(self.n_lf):
                    beta_new = self.tempering(k + 1, self.n_lf, b0)
                    scales.append(beta_old / beta_new)
                    beta_old = beta_new
            res = _capi.pythae_hmc_run(tables_for(self.model), z, gammas.float().contiguous(), accs.float().contiguous(),
                                       self.n_lf, self.eps_lf, b0, scales, path=kernel_path_for(self.model),
                                       want_stats=record is not None, want_trace=record is not None)
            if record is not None:
                h0, h, alpha, moves = res['stats']
                for i in range(iters):
                    for name, val in (('H0', h0[i]), ('H', h[i]), ('alpha', alpha[i]), ('moves', moves[i].to(torch.int)),
                                      ('z', res['trace'][i])):
                        record.setdefault(name, []).append(val)
            if state is not None:
                state['beta_old'] = beta_old
            return z

    def hmc_sampling_with_streams_stepwise(self, idx0: Optional[torch.Tensor], gammas: torch.Tensor, accs: torch.Tensor,
                                           record: Optional[dict] = None, z_start: Optional[torch.Tensor] = None,
                                           state: Optional[dict] = None) -> torch.Tensor:
        """The same loop written out with device tensors, one ``rlvae_pythae_eval`` per leapfrog step (the
        cross-check of the library loop; ~10x the launch overhead at the reference's batch of 32)."""
        with torch.no_grad():
            z0 = (self.model.centroids_tens[idx0] if z_start is None else z_start).float().contiguous()
            n, d = z0.shape
            b0 = self.beta_zero_sqrt
            beta_old = b0 if state is None else state['beta_old']
            z = z0
            eps = self.eps_lf
            cached = None if state is None else state.get('logp_grad')
            lp0, g0 = cached if cached is not None else self._logp_grad(z)
            for i in range(gammas.shape[0]):
                rho = gammas[i] / b0
                h0 = -lp0 + 0.5 * torch.norm(rho, dim=1) ** 2
                lp, g = lp0, g0
                for k in range(self.n_lf):
                    rho_ = rho + (eps / 2) * g                    # rho - (eps/2) * (-grad log_pi)
                    z = (z + eps * rho_).contiguous()
                    lp, g = self._logp_grad(z)                    # also the first gradient of step k + 1
                    rho__ = rho_ + (eps / 2) * g
                    beta_new = self.tempering(k + 1, self.n_lf, b0)
                    rho = (beta_old / beta_new) * rho__
                    beta_old = beta_new
                h = -lp + 0.5 * torch.norm(rho, dim=1) ** 2
                alpha = torch.exp(-h) / torch.exp(-h0)
                mv = accs[i] < alpha
                moves = mv.to(torch.int).reshape(n, 1)
                z = (z * moves + (1 - moves) * z0).contiguous()
                z0 = z
                lp0 = torch.where(mv, lp, lp0)                    # log_pi / gradient at the position kept
                g0 = torch.where(mv.reshape(n, 1), g, g0)
                # (a rejected proposal that overflowed leaves 0 * inf = NaN in z, exactly as in the reference,
                # which then evaluates log_pi at NaN)
                lost = ~torch.isfinite(z).all(dim=1)
                lp0 = torch.where(lost, torch.full_like(lp0, float('nan')), lp0)
                g0 = torch.where(lost.reshape(n, 1), torch.full_like(g0, float('nan')), g0)
                if record is not None:
                    for name, val in (('H0', h0), ('H', h), ('alpha', alpha), ('moves', moves.reshape(-1)),
                                      ('z', z.clone())):
                        record.setdefault(name, []).append(val)
            if state is not None:
                state['logp_grad'] = (lp0, g0)
                state['beta_old'] = beta_old
            return z

    def hmc_sampling(self, n_samples: int) -> torch.Tensor:
        dev = self.device
        idx = torch.randint(len(self.model.centroids_tens), (n_samples,), device=dev)
        d = self.model.latent_dim
        # drawn per iteration, in the reference's order (gamma :107, then acc :141), in slabs of at most
        # ~256 MB so that a large batch does not hold mcmc_steps x n x d draws at once
        z = None
        step = max(1, min(self.mcmc_steps_nbr, (1 << 26) // max(n_samples * (d + 1), 1)))
        done = 0
        state = dict(beta_old=self.beta_zero_sqrt)
        while done < self.mcmc_steps_nbr:
            it = min(step, self.mcmc_steps_nbr - done)
            gammas = torch.empty(it, n_samples, d, device=dev)
            accs = torch.empty(it, n_samples, device=dev)
            for i in range(it):
                gammas[i].normal_()
                accs[i].uniform_()
            z = self.hmc_sampling_with_streams(idx if z is None else None, gammas, accs, z_start=z, state=state)
            done += it
        return z

    def hmc_sampling_batches(self, sizes) -> torch.Tensor:
        """``hmc_sampling(b)`` for every b in ``sizes`` -- what pythae's ``RHVAESampler.sample`` loops over (ref
        src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:61-67) -- as ONE library call per group of
        batches.  Chains are independent: the only thing that ties a chain to its batch is the ORDER of the random
        draws, so the draws are made here batch by batch exactly as the sequential loop makes them (start centroids
        :100, then per MCMC iteration gamma :107 and acc :141) into slices of the stacked streams; with the same
        generator state the samples are those of the sequential loop, at the cost of one batch."""
        sizes = [int(b) for b in sizes if int(b) > 0]
        dev, d, steps = self.device, self.model.latent_dim, self.mcmc_steps_nbr
        k_cent = len(self.model.centroids_tens)
        per_call = max(1, (1 << 26) // max(steps * (d + 1), 1))      # chains per call: <= ~256 MB of draws
        out, i = [], 0
        while i < len(sizes):
            grp, tot = [], 0
            while i < len(sizes) and (not grp or tot + sizes[i] <= per_call):
                grp.append(sizes[i]); tot += sizes[i]; i += 1
            if len(grp) == 1:
                out.append(self.hmc_sampling(grp[0]))
                continue
            idx = torch.empty(tot, dtype=torch.long, device=dev)
            gammas = torch.empty(steps, tot, d, device=dev)
            accs = torch.empty(steps, tot, device=dev)
            lo = 0
            for b in grp:
                idx[lo:lo + b] = torch.randint(k_cent, (b,), device=dev)
                for it in range(steps):
                    gammas[it, lo:lo + b].normal_()
                    accs[it, lo:lo + b].uniform_()
                lo += b
            out.append(self.hmc_sampling_with_streams(idx, gammas, accs))
        return torch.cat(out, dim=0) if out else torch.empty(0, d, device=dev)

    # ------------------------------------------------------------------ BaseRiemannianSampler surface
    def sample_prior(self, num_samples: int, method: str = 'official') -> torch.Tensor:
        return self.hmc_sampling(num_samples).detach()

    def sample_riemannian_latents(self, mu: torch.Tensor, log_var: torch.Tensor,
                                  method: str = 'standard') -> torch.Tensor:
        """Standard reparameterisation (the reference's 'official' training path wraps a pythae model
        object that is out of scope; its fallback, ref src/models/samplers/rhvae_sampler.py, is this)."""
        return mu + torch.randn_like(mu) * torch.exp(0.5 * log_var)

    def get_sampler_info(self) -> Dict[str, Any]:
        info = super().get_sampler_info()
        info.update({'mcmc_steps_nbr': self.mcmc_steps_nbr, 'n_lf': self.n_lf, 'eps_lf': self.eps_lf,
                     'beta_zero_sqrt': self.beta_zero_sqrt})
        return info


class _FixedTemperatureModel:
    """The sampler protocol of ``model`` with another temperature -- what setup_official_rhvae builds
    (ref src/models/samplers/rhvae_sampler.py:83-101: same centroids / M / lambda, temperature.data = 0.1)."""

    def __init__(self, model, temperature: float):
        self._model = model
        self.temperature = torch.as_tensor(float(temperature), device=model.centroids_tens.device)
        self.latent_dim = model.latent_dim

    centroids_tens = property(lambda self: self._model.centroids_tens)
    M_tens = property(lambda self: self._model.M_tens)
    lbd = property(lambda self: self._model.lbd)
    device = property(lambda self: self._model.centroids_tens.device)
    _rlvae_kernel_path = property(lambda self: kernel_path_for(self._model))

    def parameters(self):
        return self._model.parameters()

    def G_inv(self, z):
        from ..metric_tensor import _InverseMetricFn
        from .hmc_sampler import _TabOwner
        return _InverseMetricFn.apply(z.float(), _TabOwner(tables_for(self)), kernel_path_for(self._model))

    def G(self, z):
        return torch.linalg.inv(self.G_inv(z))


class OfficialRHVAESampler(BaseRiemannianSampler):
    """Drop-in for the reference's ``OfficialRHVAESampler`` (ref src/models/samplers/rhvae_sampler.py).

    The reference builds a pythae ``RHVAE`` model object around the encoder / decoder (out of scope) only to
    hand its sampler the metric with the temperature overwritten by 0.1 (:62, :80); what the two public
    methods then compute is mirrored here on the CUDA kernels:

    * ``sample_riemannian_latents(mu, log_var, 'official')`` (:108-167): G_inv(mu) at T = 0.1,
      ``cholesky(G_inv + 1e-6 I) @ eps``, ``z = mu + 0.1 * (L eps) * sigma`` -- differentiable w.r.t. mu and
      log_var like the reference; a failed factorisation or any other error falls back to standard
      reparameterisation;
    * ``sample_prior(n, 'official')`` (:169-191): pythae's manifold HMC (100 MCMC steps x 15 leapfrog steps of
      0.03, beta_zero 1) in batches of ``min(32, n)`` (:186) -- the latent samples are returned (the
      reference's pythae sampler goes on to decode them with the model's decoder, which is not part of
      the metric path).
    """

    HARD_CODED_TEMPERATURE = 0.1      # ref :62 and :80 ("Same hardcoded value as test")
    PRIOR_BATCH = 32                  # ref :186

    def __init__(self, model):
        super().__init__(model)
        self._rhvae_model = None
        self._rhvae_sampler = None

    def setup_official_rhvae(self):
        if not self.validate_metric_availability():
            raise RuntimeError('Model must have loaded metric tensors first')
        self._rhvae_model = _FixedTemperatureModel(self.model, self.HARD_CODED_TEMPERATURE)
        self._rhvae_sampler = RHVAEStyleHMCSampler(self._rhvae_model, mcmc_steps_nbr=100, n_lf=15, eps_lf=0.03,
                                                   beta_zero=1.0)

    def official_with_noise(self, mu, log_var, eps):
        """The main path of :121-152 with ``eps`` supplied."""
        from .riemannian_sampler import _CholApplyFn
        if self._rhvae_model is None:
            self.setup_official_rhvae()
        g_inv = self._rhvae_model.G_inv(mu)
        eps_t = _CholApplyFn.apply(g_inv, eps, 1e-6)      # raises if not positive definite -> caller's fallback
        return mu + eps_t * torch.exp(0.5 * log_var) * 0.1

    def sample_riemannian_latents(self, mu, log_var, method: str = 'official'):
        if method != 'official':
            return mu + torch.randn_like(mu) * torch.exp(0.5 * log_var)
        try:
            if self._rhvae_model is None:
                self.setup_official_rhvae()
            eps = torch.randn_like(mu)
            try:
                return self.official_with_noise(mu, log_var, eps)
            except Exception:           # ref :149-151: Cholesky failed -> standard sampling with the same eps
                return mu + eps * torch.exp(0.5 * log_var)
        except Exception as e:
            print(f'⚠️ Official RHVAE sampling failed: {e}, using standard reparam')
            return mu + torch.randn_like(mu) * torch.exp(0.5 * log_var)

    def sample_prior(self, num_samples: int, method: str = 'official'):
        if method != 'official':
            return torch.randn(num_samples, self.model.latent_dim, device=self.device)
        if self._rhvae_sampler is None:
            self.setup_official_rhvae()
        bs = min(self.PRIOR_BATCH, num_samples)
        if bs <= 0:
            return torch.empty(0, self.model.latent_dim, device=self.device)
        sizes = [bs] * (num_samples // bs) + ([num_samples % bs] if num_samples % bs else [])
        with torch.no_grad():
            # the batches of pythae RHVAESampler.sample :61-67, drawn in its order, run as one library call
            return self._rhvae_sampler.hmc_sampling_batches(sizes)

    def get_sampling_methods(self) -> Dict[str, str]:
        return {'official': 'Official RHVAE sampling with HMC',
                'standard': 'Standard reparameterization (fallback)'}

    def get_rhvae_info(self) -> Dict[str, Any]:
        info = {'rhvae_available': True, 'rhvae_model_created': self._rhvae_model is not None,
                'rhvae_sampler_created': self._rhvae_sampler is not None}
        if self._rhvae_model is not None:
            info.update({'rhvae_temperature': float(self._rhvae_model.temperature.item()),
                         'rhvae_regularization': float(self._rhvae_model.lbd),
                         'rhvae_latent_dim': self._rhvae_model.latent_dim})
        return info

    def validate_metric_availability(self) -> bool:
        return all(hasattr(self.model, a) for a in ('centroids_tens', 'M_tens', 'G', 'G_inv', 'temperature', 'lbd'))
