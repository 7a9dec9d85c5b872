"""Synthetic pretrained-metric generator for the BASELINE.json workloads.

There is no network, so every benchmark and most parity tests run on synthetic
centroid tables of the named shapes.  The recipe is the one SURVEY.md §8(d)
fixes (it is part of the measurement contract, so bench.py, the tests and the
oracle-side golden generator must all draw the *same* tables from a seed):

* centroids  c = randn(K, d)                     (CPU generator, seed ``seed``)
* factors    L = tril(randn(K, d, d)) * d**-0.5  ->  M_k = L_k L_k^T  (SPD-ish)
* temperature T = 0.75 * sqrt(d)   (3.0 at d=16, the value the reference's
  conf/model/riemannian_flow_vae.yaml:49 uses), regularisation lambda = 0.01
* calibration: M is rescaled by one scalar so that det G^{-1}(z) has geometric
  mean ~1 over probe points.  Without it fp32 ``det`` overflows at K=10k and the
  HMC ``log_pi`` of the reference (hmc_sampler.py:26-30) is meaningless.

Everything here is plain CPU torch; nothing touches the CUDA extension.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class SyntheticMetric:
    centroids: torch.Tensor        # [K, d] fp32
    metric_matrices: torch.Tensor  # [K, d, d] fp32, symmetric PSD
    temperature: float
    regularization: float

    def as_load_kwargs(self) -> dict:
        """kwargs for ``MetricTensor.load_pretrained`` (metric_tensor.py:59-65)."""
        return dict(centroids=self.centroids, metric_matrices=self.metric_matrices,
                    temperature=self.temperature, regularization=self.regularization)


def _mean_logdet_ginv_fp64(z, c, M, T, lam, chunk=256):
    """mean over probe points of log det G^{-1}(z), all in fp64 (GEMM form)."""
    K, d = c.shape
    Mf = M.reshape(K, d * d)
    c2 = (c * c).sum(-1)
    eye = torch.eye(d, dtype=torch.float64)
    acc = 0.0
    for s in range(0, z.shape[0], chunk):
        zz = z[s:s + chunk]
        d2 = (zz * zz).sum(-1, keepdim=True) + c2[None, :] - 2.0 * zz @ c.T
        w = torch.exp(-d2.clamp_min(0.0) / (T * T))
        Ginv = (w @ Mf).reshape(-1, d, d) + lam * eye
        acc += torch.linalg.slogdet(Ginv).logabsdet.sum().item()
    return acc / z.shape[0]


def make_synthetic_metric(n_centroids: int, latent_dim: int, seed: int = 0,
                          temperature: float | None = None,
                          regularization: float = 0.01,
                          n_probe: int = 1024, calibrate: bool = True) -> SyntheticMetric:
    d, K = latent_dim, n_centroids
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(K, d, generator=g)
    L = torch.tril(torch.randn(K, d, d, generator=g)) * (d ** -0.5)
    M = L @ L.transpose(-1, -2)
    M = 0.5 * (M + M.transpose(-1, -2))          # exactly symmetric in fp32
    T = float(temperature) if temperature is not None else 0.75 * math.sqrt(d)
    lam = float(regularization)
    if calibrate:
        gp = torch.Generator().manual_seed(seed + 7919)
        zp = torch.randn(n_probe, d, generator=gp, dtype=torch.float64)
        c64, M64 = c.double(), M.double()
        for _ in range(3):
            m = _mean_logdet_ginv_fp64(zp, c64, M64, T, lam)
            if abs(m) <= 0.05:
                break
            M64 = M64 / math.exp(m / d)
        M = M64.float()
        M = 0.5 * (M + M.transpose(-1, -2))
    return SyntheticMetric(c.contiguous(), M.contiguous(), T, lam)


def make_points(n: int, latent_dim: int, seed: int = 1) -> torch.Tensor:
    """Latent points z = randn(N, d) on the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, latent_dim, generator=g)


def make_hmc_streams(n_chains: int, latent_dim: int, mcmc_steps: int, seed: int = 2):
    """Random draws in the order RiemannianHMCSampler.sample consumes them
    (hmc_sampler.py:114 z0, :122 gamma per MCMC iteration, :158 acc per iteration).

    Returns (z0 [n,d], gamma [mcmc,n,d], acc [mcmc,n]).  Generated globally so
    that a rank consuming rows [r*n/W, (r+1)*n/W) gets world-size independent
    results.
    """
    g = torch.Generator().manual_seed(seed)
    z0 = torch.randn(n_chains, latent_dim, generator=g)
    gam, acc = [], []
    for _ in range(mcmc_steps):
        gam.append(torch.randn(n_chains, latent_dim, generator=g))
        acc.append(torch.rand(n_chains, generator=g))
    return z0, torch.stack(gam), torch.stack(acc)
