"""Golden vectors for the metric-dependent losses (A21), produced by the REAL reference code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_losses

Runs ``LossManager.compute_riemannian_kl_loss`` (ref src/models/components/loss_manager.py:75-146) with
the reference ``MetricTensor``, and the monolith's ``compute_riemannian_metric_kl_loss`` /
``compute_riemannian_kl_loss`` (ref src/models/riemannian_flow_vae.py:1004-1077, 1328-1394; the class is
imported with the stubs of SURVEY.md 8c plus a 3-line omegaconf stub and its methods are called on a
minimal object exposing ``G`` and ``training``), each with backward to mu and log_var.  Two table sets:
K = 300 (stored) and the benchmark's K = 10,000 (regenerated from the seed by the tests).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.make_golden import _quiet, _save, near_centroids  # noqa: E402
from rlvae_b200.synthetic import make_synthetic_metric  # noqa: E402


def _ref_classes():
    ref_loader.flow_modules()                       # pythae path + sklearn_extra / imageio stubs
    if 'omegaconf' not in sys.modules:
        oc = types.ModuleType('omegaconf')
        oc.DictConfig = dict
        oc.OmegaConf = type('OmegaConf', (), {})
        sys.modules['omegaconf'] = oc
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    import importlib
    with _quiet():
        mono = importlib.import_module('src.models.riemannian_flow_vae')
    lm = ref_loader._load('_ref_loss_manager', 'src/models/components/loss_manager.py')
    return lm.LossManager, mono.RiemannianFlowVAE


class _FakeMonolith:
    training = False

    def __init__(self, mt):
        self._mt = mt

    def G(self, z):
        return self._mt.compute_metric(z)


def loss_case(name, K, n, seed, store_tables):
    LossManager, Mono = _ref_classes()
    sm = make_synthetic_metric(K, 16, seed=0)
    c, M, T, lam = sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    g = torch.Generator().manual_seed(seed)
    mu = near_centroids(c, n, 0.5, seed + 1)
    log_var = -1.0 + 0.3 * torch.randn(n, 16, generator=g)
    eps = torch.randn(n, 16, generator=g)
    out = dict(mu=mu, log_var=log_var, eps=eps, table_seed=np.int64(0), n_centroids=np.int64(K))
    if store_tables:
        out.update(centroids=c, matrices=M, temperature=np.float64(T), regularization=np.float64(lam))
    lm = LossManager(beta=1.0, device=torch.device('cpu'))
    fake = _FakeMonolith(mt)

    def run(tag, fn):
        m = mu.clone().requires_grad_(True)
        lv = log_var.clone().requires_grad_(True)
        z = m + eps * torch.exp(0.5 * lv)                    # reparameterised sample: gradients reach mu, log_var through z too
        with _quiet():
            loss = fn(m, lv, z)
        loss.backward()
        out[tag + '_loss'] = loss.detach()
        out[tag + '_dmu'] = m.grad
        out[tag + '_dlogvar'] = lv.grad

    run('modular_kl', lambda m, lv, z: lm.compute_riemannian_kl_loss(m, lv, z, mt))
    run('mono_metric_kl', lambda m, lv, z: Mono.compute_riemannian_metric_kl_loss(fake, m, lv, z))
    run('mono_kl', lambda m, lv, z: Mono.compute_riemannian_kl_loss(fake, m, lv, z))
    _save(name, **out)


def main():
    warnings.simplefilter('ignore')
    assert ref_loader.available(), 'needs /root/reference'
    loss_case('losses_d16_k300', 300, 48, 51, True)
    loss_case('losses_d16_k10k', 10000, 64, 52, False)


if __name__ == '__main__':
    main()
