"""Metric construction on the GPU: the step that produces the tables the hot path consumes.

Host-side mirror of ref ``scripts/train_and_extract_vanilla_vae.py:187-252`` (local weighted
covariance around every centroid -> ``M_matrices``; the result dictionary is what
``MetricLoader.load_from_file`` / ``MetricTensor.load_pretrained`` accept) and of the RHVAE
accumulation ``M = L L^T`` (ref ``src/lib/src/pythae/models/rhvae/rhvae_model.py:172-178,381-385``).

The reference loops over centroids in Python with O(K) passes over all N latents; here one CUDA
launch (``rlvae_local_covariance``) streams the latents once per centroid CTA, and the
minimum-eigenvalue lift uses the batched eigenvalue kernel.  Centroid *selection* (k-medoids in the
reference, via sklearn_extra) is not part of this module: pass the centroids (or their indices).
There is no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import _capi


def build_local_metrics(all_mus: torch.Tensor, centroids: torch.Tensor, temperature: float = 0.1,
                        regularization: float = 0.01, min_eigenvalue: float = 1e-6) -> torch.Tensor:
    """M [K,d,d] = weighted covariance of ``all_mus`` [N,d] around every centroid + reg I, lifted so
    that its smallest eigenvalue is at least ``min_eigenvalue`` (ref lines 204-226)."""
    if not (all_mus.is_cuda and centroids.is_cuda):
        raise RuntimeError('rlvae_b200.metric_builder runs on CUDA only (no CPU fallback)')
    d = all_mus.shape[1]
    cov = _capi.local_covariance(all_mus.float(), centroids.float(), float(temperature))
    eye = torch.eye(d, device=cov.device, dtype=cov.dtype)
    m = cov + regularization * eye
    m = 0.5 * (m + m.transpose(1, 2)) if d > 16 else m      # the d <= 16 kernel writes exactly symmetric output
    min_eig = _capi.sym_eigvalsh(m)[:, 0]
    lift = torch.clamp(min_eigenvalue - min_eig, min=0.0)
    lift = torch.where(min_eig < min_eigenvalue, lift, torch.zeros_like(lift))
    return m + lift[:, None, None] * eye


def build_metric_data(all_mus: torch.Tensor, centroids: Optional[torch.Tensor] = None,
                      centroid_indices: Optional[torch.Tensor] = None, temperature: float = 0.1,
                      regularization: float = 0.01) -> Dict[str, Any]:
    """The dictionary the reference saves to ``metric.pt`` (ref lines 241-248)."""
    if centroids is None:
        if centroid_indices is None:
            raise ValueError('pass centroids or centroid_indices (medoid selection is not part of this module)')
        centroids = all_mus[centroid_indices]
    m = build_local_metrics(all_mus, centroids, temperature, regularization)
    return {'centroids': centroids, 'M_matrices': m, 'temperature': torch.tensor(temperature),
            'regularization': torch.tensor(regularization), 'latent_dim': int(all_mus.shape[1]),
            'n_centroids': int(centroids.shape[0])}


def metric_statistics(m: torch.Tensor) -> Dict[str, float]:
    """The statistics the reference prints after construction (ref lines 228-239)."""
    ev = _capi.sym_eigvalsh(m)
    cond = ev[:, -1] / (ev[:, 0] + 1e-10)
    logdet = torch.log(ev.clamp_min(1e-38)).sum(-1)
    return {'min_eigenvalue': ev[:, 0].min().item(), 'max_eigenvalue': ev[:, -1].max().item(),
            'mean_condition_number': cond.mean().item(),
            'determinant_range': (torch.exp(logdet.min()).item(), torch.exp(logdet.max()).item())}


def matrices_from_cholesky_factors(l_factors: torch.Tensor) -> torch.Tensor:
    """M_k = L_k L_k^T (RHVAE parametrisation, ref rhvae_model.py:172-178): a plain batched GEMM."""
    return l_factors @ l_factors.transpose(-1, -2)
