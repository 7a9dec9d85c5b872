"""Load the *real* reference modules from /root/reference, file by file.

TEST INFRASTRUCTURE ONLY (see oracle/metric_oracle.py).  Works only where the
reference checkout exists (the build container); on the GPU box
``available()`` is False and callers skip.  Nothing is copied: the modules are
executed from where they lie.  Recipe from SURVEY.md §8(c): metric_tensor.py,
metric_loader.py and the samplers import only torch, so
``importlib.util.spec_from_file_location`` is enough (the samplers need a stub
parent package for ``from .base_sampler import ...``); the package-level
``import src.models`` would pull pythae -> sklearn_extra/imageio (absent).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get('RLVAE_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'src/models/components/metric_tensor.py'))


def _load(name, relpath, package=None):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def modules():
    """-> dict(metric_tensor, metric_loader, base_sampler, hmc_sampler, riemannian_sampler)."""
    if 'metric_tensor' in _cache:
        return _cache
    if not available():
        raise RuntimeError(f'reference checkout not found under {REF_ROOT}')
    _cache['metric_tensor'] = _load('_ref_metric_tensor', 'src/models/components/metric_tensor.py')
    _cache['metric_loader'] = _load('_ref_metric_loader', 'src/models/components/metric_loader.py')
    pkg = types.ModuleType('_ref_samplers')
    pkg.__path__ = [os.path.join(REF_ROOT, 'src/models/samplers')]
    sys.modules['_ref_samplers'] = pkg
    for m in ('base_sampler', 'hmc_sampler', 'riemannian_sampler'):
        _cache[m] = _load(f'_ref_samplers.{m}', f'src/models/samplers/{m}.py', package='_ref_samplers')
    return _cache


def flow_modules():
    """Reference FlowManager + vendored pythae IAF (needs two tiny stubs)."""
    if 'flow_manager' in _cache:
        return _cache
    lib = os.path.join(REF_ROOT, 'src/lib/src')
    if lib not in sys.path:
        sys.path.insert(0, lib)
    if 'sklearn_extra' not in sys.modules:
        se = types.ModuleType('sklearn_extra')
        sec = types.ModuleType('sklearn_extra.cluster')
        sec.KMedoids = object
        se.cluster = sec
        sys.modules['sklearn_extra'] = se
        sys.modules['sklearn_extra.cluster'] = sec
    if 'imageio' not in sys.modules:
        io = types.ModuleType('imageio')
        io.imwrite = lambda *a, **k: None
        sys.modules['imageio'] = io
    _cache['flow_manager'] = _load('_ref_flow_manager', 'src/models/components/flow_manager.py')
    return _cache


class RefModel(torch.nn.Module):
    """Minimal model exposing the protocol every reference sampler touches
    (base_sampler.py:27-28,67; hmc_sampler.py:19-21,39-42,108): G, G_inv,
    centroids_tens, M_tens, temperature, lbd, latent_dim, device, parameters()."""

    def __init__(self, metric_tensor_module):
        super().__init__()
        self._p = torch.nn.Parameter(torch.zeros(1))
        self.mt = metric_tensor_module
        self.latent_dim = metric_tensor_module.latent_dim
        self.device = torch.device('cpu')
        self.centroids_tens = metric_tensor_module.centroids
        self.M_tens = metric_tensor_module.metric_matrices
        self.temperature = metric_tensor_module.temperature
        self.lbd = metric_tensor_module.regularization

    def G(self, z):
        return self.mt.compute_metric(z)

    def G_inv(self, z):
        return self.mt.compute_inverse_metric(z)


def make_ref_metric(centroids, matrices, temperature, regularization):
    """Reference MetricTensor on CPU loaded with the given tables."""
    import contextlib
    import io
    MT = modules()['metric_tensor'].MetricTensor
    mt = MT(latent_dim=centroids.shape[1], device=torch.device('cpu'))
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(centroids.clone(), matrices.clone(),
                           temperature=float(temperature), regularization=float(regularization))
    return mt


class RecordingRNG:
    """Context manager that records every torch.randn / randn_like / rand /
    randint draw made by reference code, in call order, so the oracle and the
    CUDA path can be fed the identical stream."""

    def __init__(self):
        self.draws = []
        self._orig = {}

    def __enter__(self):
        for name in ('randn', 'randn_like', 'rand', 'randint'):
            self._orig[name] = getattr(torch, name)

            def wrap(*a, __f=self._orig[name], __n=name, **k):
                out = __f(*a, **k)
                self.draws.append((__n, out.detach().clone()))
                return out
            setattr(torch, name, wrap)
        return self

    def __exit__(self, *exc):
        for name, f in self._orig.items():
            setattr(torch, name, f)
        return False
