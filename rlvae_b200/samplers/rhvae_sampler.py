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

Per leapfrog step: two metric evaluations (fused forward kernel -> G) and two gradient
contractions; the momentum / position updates are element-wise torch ops on the device.
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
    def _eval(self, z):
        mt = getattr(self.model, 'metric_tensor', None)
        tab = tables_for(self.model)
        ev = _capi.metric_eval(tab, z, want_ginv=False, want_g=True, want_logdet=True, want_grad=False,
                               path=kernel_path_for(self.model))
        return tab, ev

    def log_sqrt_det_G_inv(self, z: torch.Tensor) -> torch.Tensor:
        """log(sqrt(det G^{-1}(z)) + 1e-10)  (ref :157-158; a negative determinant gives NaN there too)."""
        tab = tables_for(self.model)
        ginv = _capi.inverse_metric(tab, z.contiguous().float(), kernel_path_for(self.model))
        _, lad, sgn, _ = _capi.batched_inverse(ginv, want_inv=False, want_logabsdet=True, want_sign=True)
        det = sgn * torch.exp(lad)
        return torch.log(torch.sqrt(det) + 1e-10)

    def grad_log_sqrt_det_G_inv(self, z: torch.Tensor) -> torch.Tensor:
        """[N,d]  (the reference returns [N,d,1] and reshapes, ref :118-120)."""
        tab, ev = self._eval(z.contiguous().float())
        return _capi.metric_grad_pythae(tab, z.contiguous().float(), ev['g'])

    @staticmethod
    def tempering(k: int, K: int, beta_zero_sqrt: float) -> float:
        beta_k = ((1 - 1 / beta_zero_sqrt) * (k / K) ** 2) + 1 / beta_zero_sqrt
        return 1 / beta_k

    # ------------------------------------------------------------------ the sampling loop (ref :98-148)
    def hmc_sampling_with_streams(self, idx0: torch.Tensor, gammas: torch.Tensor, accs: torch.Tensor,
                                  record: Optional[dict] = None) -> torch.Tensor:
        """The loop with its random draws injected: ``idx0 [n]`` (line 100), ``gammas [steps,n,d]``
        (line 107), ``accs [steps,n]`` (line 141)."""
        with torch.no_grad():
            z0 = self.model.centroids_tens[idx0].float().contiguous()
            n, d = z0.shape
            b0 = self.beta_zero_sqrt
            beta_old = b0
            z = z0
            eps = self.eps_lf
            for i in range(gammas.shape[0]):
                rho = gammas[i] / b0
                h0 = -self.log_sqrt_det_G_inv(z) + 0.5 * torch.norm(rho, dim=1) ** 2
                for k in range(self.n_lf):
                    g = -self.grad_log_sqrt_det_G_inv(z)
                    rho_ = rho - (eps / 2) * g
                    z = (z + eps * rho_).contiguous()
                    g = -self.grad_log_sqrt_det_G_inv(z)
                    rho__ = rho_ - (eps / 2) * g
                    beta_new = self.tempering(k + 1, self.n_lf, b0)
                    rho = (beta_old / beta_new) * rho__
                    beta_old = beta_new
                h = -self.log_sqrt_det_G_inv(z) + 0.5 * torch.norm(rho, dim=1) ** 2
                alpha = torch.exp(-h) / torch.exp(-h0)
                moves = (accs[i] < alpha).to(torch.int).reshape(n, 1)
                z = (z * moves + (1 - moves) * z0).contiguous()
                z0 = z
                if record is not None:
                    for name, val in (('H0', h0), ('H', h), ('alpha', alpha), ('moves', moves.reshape(-1)),
                                      ('z', z.clone())):
                        record.setdefault(name, []).append(val)
            return z

    def hmc_sampling(self, n_samples: int) -> torch.Tensor:
        dev = self.device
        idx = torch.randint(len(self.model.centroids_tens), (n_samples,), device=dev)
        d = self.model.latent_dim
        gammas = torch.randn(self.mcmc_steps_nbr, n_samples, d, device=dev)
        accs = torch.rand(self.mcmc_steps_nbr, n_samples, device=dev)
        return self.hmc_sampling_with_streams(idx, gammas, accs)

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
