"""Sampler base class + the model-side protocol (ref src/models/samplers/base_sampler.py:13-96).

Every sampler talks to the model through six attributes only (``G``, ``G_inv``,
``centroids_tens``, ``M_tens``, ``temperature``, ``lbd``) plus ``latent_dim``,
``device`` and ``parameters()``.  ``MetricModel`` below is the
smallest object with that surface, built on the CUDA ``MetricTensor``;
``tables_for(model)`` gives the fused kernels their packed-table handle for any
model that satisfies the protocol.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict

import torch
import torch.nn as nn

from .. import _capi
from ..metric_tensor import MetricTensor


class MetricModel(nn.Module):
    """Protocol adapter: what ``modular_rlvae.py:239-261`` installs on the real model."""

    def __init__(self, metric_tensor: MetricTensor):
        super().__init__()
        self.metric_tensor = metric_tensor
        self._anchor = nn.Parameter(torch.zeros(1, device=metric_tensor.centroids.device))
        self.latent_dim = metric_tensor.latent_dim

    @property
    def device(self):
        return self.metric_tensor.centroids.device

    @property
    def centroids_tens(self):
        return self.metric_tensor.centroids

    @property
    def M_tens(self):
        return self.metric_tensor.metric_matrices

    @property
    def temperature(self):
        return self.metric_tensor.temperature

    @property
    def lbd(self):
        return self.metric_tensor.regularization

    def G(self, z):
        return self.metric_tensor.compute_metric(z)

    def G_inv(self, z):
        return self.metric_tensor.compute_inverse_metric(z)


def tables_for(model) -> _capi.Tables:
    """Packed CUDA tables for a protocol model (cached on the model, keyed on the buffers)."""
    mt = getattr(model, 'metric_tensor', None)
    if isinstance(mt, MetricTensor):
        return mt._tables(mt.centroids.device)
    c, m = model.centroids_tens, model.M_tens
    T, lam = float(model.temperature), float(model.lbd)
    key = (c.data_ptr(), c._version, m.data_ptr(), m._version, tuple(c.shape), T, lam)
    cached = getattr(model, '_rlvae_b200_tables', None)
    if cached is None or cached[0] != key:
        cached = (key, _capi.Tables(c.float(), m.float(), T, lam))
        model._rlvae_b200_tables = cached
    return cached[1]


def kernel_path_for(model) -> int:
    mt = getattr(model, 'metric_tensor', None)
    if isinstance(mt, MetricTensor):
        return mt._path()
    return int(getattr(model, '_rlvae_kernel_path', _capi.PATH_AUTO))     # wrappers forward their model's choice


class BaseRiemannianSampler(ABC):
    def __init__(self, model):
        self.model = model
        self.device = next(model.parameters()).device

    @abstractmethod
    def sample_riemannian_latents(self, mu: torch.Tensor, log_var: torch.Tensor,
                                  method: str = 'enhanced') -> torch.Tensor:
        """mu, log_var [N,d] -> z [N,d]."""

    @abstractmethod
    def sample_prior(self, num_samples: int, method: str = 'geodesic') -> torch.Tensor:
        """-> [num_samples, d]."""

    def validate_metric_availability(self) -> bool:
        return all(hasattr(self.model, a) for a in ('centroids_tens', 'M_tens', 'G', 'G_inv'))

    def get_sampling_methods(self) -> Dict[str, str]:
        return {
            'enhanced': 'Enhanced Riemannian sampling with centroid influence',
            'geodesic': 'Geodesic-aware sampling along manifold paths',
            'basic': 'Basic metric-aware sampling',
            'standard': 'Standard reparameterization (no Riemannian)',
        }

    def get_sampler_info(self) -> Dict[str, Any]:
        return {
            'sampler_type': self.__class__.__name__,
            'available_methods': list(self.get_sampling_methods().keys()),
            'metric_available': self.validate_metric_availability(),
            'device': str(self.device),
        }
