"""Drop-in ``MetricLoader`` (host side only; ref src/models/components/metric_loader.py).

Same public surface and error behaviour as the reference class (:21-334):
``load_from_file`` -> ``{centroids, metric_matrices, temperature, regularization}``
with the reference's key fallbacks, ``save_to_file``, ``convert_old_format`` and
``validate_metric_file``.  The only functional change is that the per-centroid
Python eigenvalue loops (:211-214, :306-310 -- minutes at K=10k) are batched.
"""
from __future__ import annotations

import warnings
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch

_CENTROID_KEYS = ('centroids', 'metric_centroids', 'centers', 'mu')
_TEMPERATURE_KEYS = ('temperature', 'temp', 'T', 'beta')
_REGULARIZATION_KEYS = ('regularization', 'reg', 'lambda', 'lbd')


def _as_tensor(x, device):
    return x.to(device) if isinstance(x, torch.Tensor) else torch.tensor(x, device=device)


def _scalar(data, keys, override, default, what):
    if override is not None:
        return float(override)
    for k in keys:
        if k in data:
            v = data[k]
            return float(v.item()) if isinstance(v, torch.Tensor) else float(v)
    warnings.warn(f'No {what} found, using default: {default}')
    return default


def _batched_eigvals_real(m: torch.Tensor) -> torch.Tensor:
    """real parts of the eigenvalues of every matrix, one batched CPU call."""
    return torch.linalg.eigvals(m.detach().to('cpu', torch.float32)).real


class MetricLoader:
    def __init__(self, device: Optional[torch.device] = None):
        self.device = device or torch.device('cuda' if torch.cuda.is_available() else 'cpu')

    # ---------------------------------------------------------------- load
    def load_from_file(self, path: Union[str, Path], temperature_override: Optional[float] = None,
                       regularization_override: Optional[float] = None) -> Dict[str, Any]:
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f'Metric file not found: {path}')
        print(f'🔧 Loading metric data from: {path}')
        try:
            raw = torch.load(path, map_location=self.device)
        except Exception as e:  # same wrapping as ref :60-63
            raise RuntimeError(f'Failed to load metric file: {e}')

        centroids = self._extract_centroids(raw)
        matrices = self._extract_metric_matrices(raw, centroids.shape)
        temperature = _scalar(raw, _TEMPERATURE_KEYS, temperature_override, 0.1, 'temperature')
        regularization = _scalar(raw, _REGULARIZATION_KEYS, regularization_override, 0.01, 'regularization')
        self._validate_data_consistency(centroids, matrices)
        print(f'✅ Loaded metric: {len(centroids)} centroids, {centroids.shape[1]}D, '
              f'T={temperature:.3f}, λ={regularization:.3f}')
        return {'centroids': centroids, 'metric_matrices': matrices,
                'temperature': temperature, 'regularization': regularization}

    def _extract_centroids(self, data: Dict[str, Any]) -> torch.Tensor:
        for k in _CENTROID_KEYS:
            if k in data:
                return _as_tensor(data[k], self.device)
        raise ValueError(f'No centroids found. Expected one of: {list(_CENTROID_KEYS)}')

    def _extract_metric_matrices(self, data: Dict[str, Any], centroid_shape) -> torch.Tensor:
        n_centroids, latent_dim = centroid_shape
        if 'M_matrices' in data:
            m = data['M_matrices']
        elif 'metric_vars' in data:
            m = data['metric_vars']
        elif 'M_i_flat' in data:                     # flattened diagonals
            m = torch.diag_embed(_as_tensor(data['M_i_flat'], self.device))
        elif 'M_tens' in data:
            m = data['M_tens']
        else:
            warnings.warn('No metric matrices found, using identity matrices')
            m = torch.eye(latent_dim, device=self.device).unsqueeze(0).repeat(n_centroids, 1, 1)
        m = _as_tensor(m, self.device)
        expected = (n_centroids, latent_dim, latent_dim)
        if tuple(m.shape) != expected:
            raise ValueError(f'Metric matrices shape {tuple(m.shape)} != expected {expected}')
        return m

    def _extract_temperature(self, data, override):
        return _scalar(data, _TEMPERATURE_KEYS, override, 0.1, 'temperature')

    def _extract_regularization(self, data, override):
        return _scalar(data, _REGULARIZATION_KEYS, override, 0.01, 'regularization')

    def _validate_data_consistency(self, centroids: torch.Tensor, matrices: torch.Tensor) -> None:
        n_centroids, latent_dim = centroids.shape
        if tuple(matrices.shape) != (n_centroids, latent_dim, latent_dim):
            raise ValueError(f'Inconsistent shapes: centroids {tuple(centroids.shape)}, '
                             f'matrices {tuple(matrices.shape)}')
        if torch.isnan(centroids).any() or torch.isinf(centroids).any():
            raise ValueError('Centroids contain NaN or inf values')
        if torch.isnan(matrices).any() or torch.isinf(matrices).any():
            raise ValueError('Metric matrices contain NaN or inf values')
        mins = _batched_eigvals_real(matrices).min(dim=1).values
        for i in torch.nonzero(mins < -1e-6).flatten().tolist():
            warnings.warn(f'Metric matrix {i} is not positive semidefinite (min eigenval: {mins[i]:.3e})')

    # ---------------------------------------------------------------- save / convert / validate
    def save_to_file(self, path: Union[str, Path], centroids: torch.Tensor, metric_matrices: torch.Tensor,
                     temperature: float, regularization: float,
                     metadata: Optional[Dict[str, Any]] = None) -> None:
        data = {'centroids': centroids.cpu(), 'metric_matrices': metric_matrices.cpu(),
                'temperature': temperature, 'regularization': regularization}
        if metadata:
            data['metadata'] = metadata
        torch.save(data, Path(path))
        print(f'✅ Saved metric data to: {path}')

    def convert_old_format(self, old_path, new_path, temperature_override: Optional[float] = None,
                           regularization_override: Optional[float] = None) -> None:
        d = self.load_from_file(old_path, temperature_override, regularization_override)
        self.save_to_file(new_path, d['centroids'], d['metric_matrices'], d['temperature'],
                          d['regularization'], metadata={'converted_from': str(old_path)})
        print(f'✅ Converted {old_path} → {new_path}')

    def validate_metric_file(self, path) -> Dict[str, Any]:
        try:
            d = self.load_from_file(path)
            c, m = d['centroids'], d['metric_matrices']
            ev = _batched_eigvals_real(m)
            cond = ev.max(dim=1).values / (ev.min(dim=1).values + 1e-8)
            det = torch.linalg.det(m.detach().to('cpu', torch.float32))
            return {
                'valid': True, 'n_centroids': c.shape[0], 'latent_dim': c.shape[1],
                'temperature': d['temperature'], 'regularization': d['regularization'],
                'eigenvalue_range': (ev.min().item(), ev.max().item()),
                'condition_number_range': (cond.min().item(), cond.max().item()),
                'determinant_range': (det.min().item(), det.max().item()),
                'has_negative_eigenvals': bool((ev < -1e-6).any().item()),
                'mean_condition_number': cond.mean().item(),
            }
        except Exception as e:
            return {'valid': False, 'error': str(e), 'error_type': type(e).__name__}
