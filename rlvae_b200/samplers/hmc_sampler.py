"""``RiemannianHMCSampler`` on the fused CUDA leapfrog (ref src/models/samplers/hmc_sampler.py).

Same constructor, public attributes (``log_pi``, ``grad_func``, ``n_lf``, ``eps_lf``,
``beta_zero_sqrt``) and methods as the reference class (:13-296).  ``sample`` draws its
random numbers in the reference's order (:114 z0, :122 gamma, :158 acc) and hands each
chain to ``rlvae_hmc_run`` -- one metric evaluation per leapfrog step instead of the reference's four
[n,K,d,d] materialisations, and ONE kernel launch for the whole trajectory (all MCMC
iterations whose draws are in memory) when the fused trajectory kernel serves the tables.

Reference quirks reproduced on purpose: the 'gradient' integrated is
variant A == (1 - lambda*G_ii)/T^2; ``log_pi`` uses det + clamp(1e-10); the tempering
state ``beta_sqrt_old`` is not reset between MCMC iterations; alpha has ``+1e-10`` in the
denominator.  ``grad_mode='exact'`` switches the drift to the true gradient (additive).
"""
from __future__ import annotations

import math
from typing import Any, Dict

import torch

from .. import _capi
from ..metric_tensor import _InverseMetricFn, _LogAbsDetFn
from .base_sampler import BaseRiemannianSampler, kernel_path_for, tables_for

_LOG_CLAMP = 0.5 * math.log(1e-10)


def clamped_log_pi(lad: torch.Tensor, sgn: torch.Tensor) -> torch.Tensor:
    """0.5*log(clamp(det, 1e-10)) from (log|det|, sign); inf where fp32 det overflows."""
    lp = torch.clamp(0.5 * lad, min=_LOG_CLAMP)
    lp = torch.where(lad > 88.72283, torch.full_like(lp, float('inf')), lp)
    return torch.where(sgn > 0, lp, torch.full_like(lp, _LOG_CLAMP))


class RiemannianHMCSampler(BaseRiemannianSampler):
    def __init__(self, model, mcmc_steps_nbr=100, n_lf=15, eps_lf=0.03, beta_zero=1.0,
                 grad_mode: str = 'modular'):
        super().__init__(model)
        self.mcmc_steps_nbr = mcmc_steps_nbr
        self.n_lf = torch.tensor([n_lf], device=model.device)
        self.eps_lf = torch.tensor([eps_lf], device=model.device)
        self.beta_zero_sqrt = torch.tensor([beta_zero], device=model.device).sqrt()
        self.grad_mode = grad_mode
        if not self.validate_metric_availability():
            raise RuntimeError('RiemannianHMCSampler needs a model exposing G, G_inv, centroids_tens, M_tens')
        self.log_pi = self._log_sqrt_det_ginv
        self.grad_func = self._grad_modular

    # ---- public callables of the reference (:26-30, :33-68)
    def _log_sqrt_det_ginv(self, z):
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        if z.requires_grad and torch.is_grad_enabled():
            owner = _TabOwner(tab)
            ginv = _InverseMetricFn.apply(z, owner, path)
            lad = _LogAbsDetFn.apply(ginv)
            _, _, sgn, _ = _capi.batched_inverse(ginv.detach(), want_inv=False, want_sign=True)
            return clamped_log_pi(lad, sgn)
        ginv = _capi.inverse_metric(tab, z.detach(), path)
        _, lad, sgn, _ = _capi.batched_inverse(ginv, want_inv=False, want_logabsdet=True, want_sign=True)
        return clamped_log_pi(lad, sgn)

    def _grad_modular(self, z):
        """variant A: diag(-0.5 G^T ((-2/T^2) sum_k w_k M_k)^T) == (1 - lambda*G_ii)/T^2."""
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        ginv = _capi.inverse_metric(tab, z.detach(), path)
        _, _, _, diag = _capi.batched_inverse(ginv, want_inv=False, want_diag=True)
        return (1.0 - tab.regularization * diag) / (tab.temperature ** 2)

    def grad_exact(self, z):
        """variant D: grad_z 0.5*log det G^{-1} = (1/T^2) sum_k w_k tr(G M_k)(c_k - z)."""
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        zc = z.detach().contiguous()
        ginv = _capi.inverse_metric(tab, zc, path)
        gt, _, _, _ = _capi.batched_inverse(ginv, want_inv=True, transpose=True)   # tr(G M_k) = <G^T, M_k>
        return _capi.metric_grad(tab, zc, gt, 1.0 / tab.temperature ** 2, path)

    @staticmethod
    def _tempering(k, K, beta_zero_sqrt):
        beta_k = ((1 - 1 / beta_zero_sqrt) * (k / K) ** 2) + 1 / beta_zero_sqrt
        return 1 / beta_k

    def _scales(self, n_lf: int, beta_sqrt_old, beta_zero_sqrt):
        """per-step momentum rescale beta_sqrt_old/beta_sqrt in the reference's fp32 tensor
        arithmetic (:147-149); returns (list of floats, carried beta_sqrt_old).  Both tensors live on
        the HOST (IEEE fp32 reciprocal / multiply / add / divide give the same bits on either device):
        on the GPU every .item() here was a device synchronisation, n_lf of them per MCMC iteration,
        which dominated small-batch sampling (measured, K = 200, n_lf = 10: 1.1 ms -> 0.5 ms per MCMC iteration
        at 64-4096 chains; scripts/time_small_batch.py)."""
        out = []
        for k in range(n_lf):
            beta_sqrt = self._tempering(k + 1, n_lf, beta_zero_sqrt)
            out.append(float((beta_sqrt_old / beta_sqrt).item()))
            beta_sqrt_old = beta_sqrt
        return out, beta_sqrt_old

    # ---- A11
    _MAX_DRAW_BYTES = 1 << 30     # random draws generated ahead per launch of the fused trajectory kernel

    def _all_scales(self, iters: int, n_lf: int, beta_old, b0_host):
        """Tempering scales of ``iters`` consecutive MCMC iterations.  The schedule of one iteration only
        depends on the state it starts from -- beta_zero_sqrt (first iteration) or tempering(n_lf) (carried
        over, the reference never resets it) -- so each distinct schedule is computed once with the
        reference's tensor arithmetic and cached (1,500 small tensor ops per ``sample`` call otherwise:
        ~30 ms of host time, more than the fused kernel needs for the whole chain at small batches)."""
        cache = self.__dict__.setdefault('_scale_cache', {})
        out = []
        for _ in range(iters):
            key = (n_lf, float(b0_host.item()), float(beta_old.item()))
            if key not in cache:
                cache[key] = self._scales(n_lf, beta_old, b0_host)
            sc, beta_old = cache[key]
            out.extend(sc)
        return out, beta_old

    def _run_chain(self, tab, path, z, gammas, accs, n_lf, eps, b0, mode, beta_old, b0_host, record=None):
        """MCMC iterations gammas.shape[0] of the chain, in place on z: ONE kernel launch when the fused
        trajectory kernel serves these tables (``rlvae_hmc_run``), else one C-ABI call per iteration inside
        the library.  If the fused kernel reports a metric that lost positive definiteness to rounding, the
        same iterations are redone from the same state with the per-step kernels + pivoting fallback."""
        iters = gammas.shape[0]
        scales, beta_new = self._all_scales(iters, n_lf, beta_old, b0_host)
        fused = _capi.hmc_fused_available(tab, mode, path)
        z_start = z.clone() if fused else None
        want = record is not None
        res = _capi.hmc_run(tab, z, gammas, accs, n_lf, eps, b0, scales, mode, path, want_stats=want, want_trace=want)
        if fused and int(_capi.hmc_fail_count(res['work'], z.shape[0], z.shape[1]).item()) != 0:
            z.copy_(z_start)
            res = _capi.hmc_run(tab, z, gammas, accs, n_lf, eps, b0, scales, mode | _capi.HMC_NO_FUSION, path,
                                want_stats=want, want_trace=want)
        if want:
            for name, val in zip(('H0', 'H', 'alpha', 'moves'), res['stats']):
                record.setdefault(name, []).extend(list(val))
            record.setdefault('z', []).extend(list(res['trace']))
        return beta_new

    def sample_with_streams(self, z0, gammas, accs, z_forced=None, record=None):
        """``sample`` with the random draws supplied (z0 [n,d], gammas [I,n,d], accs [I,n]).
        ``z_forced[i]`` restarts iteration i from the given state (teacher forcing)."""
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        n_lf = int(self.n_lf.item())
        eps, b0 = float(self.eps_lf.item()), float(self.beta_zero_sqrt.item())
        mode = _capi.GRAD_EXACT if self.grad_mode == 'exact' else _capi.GRAD_MODULAR
        z = z0.detach().to(self.model.device, torch.float32).contiguous().clone()
        gammas = gammas.to(z.device, torch.float32).contiguous()
        accs = accs.to(z.device, torch.float32).contiguous()
        b0_host = self.beta_zero_sqrt.detach().float().cpu()
        beta_old = b0_host
        if z_forced is None:
            self._run_chain(tab, path, z, gammas, accs, n_lf, eps, b0, mode, beta_old, b0_host, record)
            return z
        for i in range(gammas.shape[0]):
            z.copy_(z_forced[i])
            beta_old = self._run_chain(tab, path, z, gammas[i:i + 1], accs[i:i + 1], n_lf, eps, b0, mode,
                                       beta_old, b0_host, record)
        return z

    def sample(self, n_samples, t=0):
        dev = self.model.device
        self.n_lf, self.eps_lf = self.n_lf.to(dev), self.eps_lf.to(dev)
        self.beta_zero_sqrt = self.beta_zero_sqrt.to(dev)
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        n_lf = int(self.n_lf.item())
        eps, b0 = float(self.eps_lf.item()), float(self.beta_zero_sqrt.item())
        mode = _capi.GRAD_EXACT if self.grad_mode == 'exact' else _capi.GRAD_MODULAR
        d = self.model.latent_dim
        z = torch.randn(n_samples, d, device=dev)
        b0_host = self.beta_zero_sqrt.detach().float().cpu()
        beta_old = b0_host
        # The draws of several MCMC iterations are generated ahead, in the reference's call order (gamma
        # :122 then acc :158 for every iteration -- nothing else draws in between), so that one launch of
        # the fused trajectory kernel runs them all; bounded so a large batch does not hold all of them.
        per_iter = 4 * n_samples * (d + 1)
        ahead = max(1, min(self.mcmc_steps_nbr, self._MAX_DRAW_BYTES // max(per_iter, 1)))
        done = 0
        while done < self.mcmc_steps_nbr:
            it = min(ahead, self.mcmc_steps_nbr - done)
            gammas = torch.empty(it, n_samples, d, device=dev)
            accs = torch.empty(it, n_samples, device=dev)
            for i in range(it):          # same generator calls as randn_like(z) / rand(n): normal_ / uniform_ on n*d / n elements
                gammas[i].normal_()
                accs[i].uniform_()
            beta_old = self._run_chain(tab, path, z, gammas, accs, n_lf, eps, b0, mode, beta_old, b0_host)
            done += it
        return z.detach()

    # ---- A12
    def sample_posterior_with_streams(self, mu, log_var, eps0, gammas, n_lf=5, step=0.01):
        tab, path = tables_for(self.model), kernel_path_for(self.model)
        T2 = tab.temperature ** 2
        mu, log_var = mu.detach().float(), log_var.detach().float()
        inv_var = torch.exp(-log_var)

        def grad_energy(zz):
            zc = zz.contiguous()
            ginv = _capi.inverse_metric(tab, zc, path)
            gt, lad, sgn, _ = _capi.batched_inverse(ginv, want_inv=True, want_logabsdet=True, want_sign=True,
                                                    transpose=True)
            glp = _capi.metric_grad(tab, zc, gt, 1.0 / T2, path)
            live = ((sgn > 0) & (0.5 * lad > _LOG_CLAMP)).to(glp.dtype)[:, None]   # clamp kills the gradient
            return -glp * live + (zz - mu) * inv_var

        z = (mu + eps0 * torch.exp(0.5 * log_var)).contiguous()
        g = grad_energy(z)      # re-used wherever the reference re-evaluates at an unchanged z
        for i in range(gammas.shape[0]):
            rho = gammas[i] * 0.1
            for _ in range(n_lf):
                rho = rho - (step / 2) * g
                z = z - step * rho            # sign as written in the reference (:210)
                g = grad_energy(z)
                rho = rho - (step / 2) * g
        return z.detach()

    def sample_posterior(self, mu, log_var, t=0):
        eps0 = torch.randn_like(mu)
        gammas = torch.stack([torch.randn_like(mu) for _ in range(20)])
        return self.sample_posterior_with_streams(mu, log_var, eps0, gammas)

    # ---- A13
    def sample_riemannian_latents(self, mu, log_var, method: str = 'hmc'):
        if method == 'posterior_hmc':
            return self.sample_posterior(mu, log_var)
        eps = torch.randn_like(mu)
        return self.refine_with_eps(mu, log_var, eps)

    def refine_with_eps(self, mu, log_var, eps, n_steps=3, step_size=0.01):
        z = (mu + eps * torch.exp(0.5 * log_var)).detach().float().contiguous().clone()
        try:
            _capi.hmc_refine(tables_for(self.model), z, n_steps, step_size, kernel_path_for(self.model))
        except Exception as e:   # the reference catches everything and keeps the standard sample
            print(f'⚠️ HMC refinement failed: {e}, using standard sampling')
        return z.detach()

    def sample_prior(self, num_samples: int, method: str = 'hmc'):
        if method == 'hmc':
            return self.sample(num_samples)
        return torch.randn(num_samples, self.model.latent_dim, device=self.device)

    def get_sampling_methods(self) -> Dict[str, str]:
        return {'hmc': 'Hamiltonian Monte Carlo sampling on manifold',
                'posterior_hmc': 'HMC sampling from posterior',
                'basic': 'Standard Gaussian sampling (fallback)'}

    def get_hmc_parameters(self) -> Dict[str, Any]:
        return {'mcmc_steps_nbr': self.mcmc_steps_nbr, 'n_lf': int(self.n_lf.item()),
                'eps_lf': float(self.eps_lf.item()), 'beta_zero': float(self.beta_zero_sqrt.item() ** 2)}


class _TabOwner:
    """lets the autograd Function of MetricTensor run against a bare Tables handle."""

    def __init__(self, tab):
        self._tab = tab

    def _tables(self, device):
        return self._tab
