"""Golden vectors for the prior samplers of WorkingRiemannianSampler (A17) and for the training-time
path of OfficialRHVAESampler, produced by the REAL reference code with its random draws recorded.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_priors

* ``sample_prior(method='centroid_aware' | 'weighted_mixture' | 'basic' | 'geodesic')``
  (ref src/models/samplers/riemannian_sampler.py:222-355).  The weighted-mixture sampler draws one
  ``randn(count_i, d)`` per non-empty component in increasing component order (:337-343); the fixture
  stores those draws concatenated in call order.
* ``OfficialRHVAESampler.sample_riemannian_latents(method='official')`` (ref
  src/models/samplers/rhvae_sampler.py:108-167): G_inv evaluated with the HARD-CODED temperature 0.1
  (:62, :80), cholesky(+1e-6 I) @ eps, z = mu + 0.1 * (L eps) * sigma.  The pythae RHVAE model object the
  reference builds around an encoder / decoder is replaced by the ten lines of it that the method touches.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.make_golden import _quiet, _save, near_centroids  # noqa: E402
from rlvae_b200.synthetic import make_synthetic_metric  # noqa: E402


def main():
    warnings.simplefilter('ignore')
    assert ref_loader.available(), 'needs /root/reference'
    mods = ref_loader.modules()
    sm = make_synthetic_metric(300, 16, seed=0)
    c, M, T, lam = sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    model = ref_loader.RefModel(mt)
    ws = mods['riemannian_sampler'].WorkingRiemannianSampler(model)
    n = 64
    out = dict(table_seed=np.int64(0), n_centroids=np.int64(300))

    torch.manual_seed(61)
    with ref_loader.RecordingRNG() as rec, _quiet():
        z = ws.sample_prior(n, method='centroid_aware')
    assert [k for k, _ in rec.draws] == ['randint', 'randn_like'], [k for k, _ in rec.draws]
    out['ca_idx'], out['ca_noise'], out['ca_z'] = rec.draws[0][1], rec.draws[1][1], z

    torch.manual_seed(62)
    with ref_loader.RecordingRNG() as rec, _quiet():
        z = ws.sample_prior(n, method='weighted_mixture')
    kinds = [k for k, _ in rec.draws]
    assert kinds[0] == 'randint' and all(k == 'randn' for k in kinds[1:]), kinds
    out['wm_idx'] = rec.draws[0][1]
    out['wm_noise'] = torch.cat([d for _, d in rec.draws[1:]], dim=0)     # call order = increasing component
    assert out['wm_noise'].shape == (n, 16)
    out['wm_z'] = z

    torch.manual_seed(63)
    with ref_loader.RecordingRNG() as rec, _quiet():
        z = ws.sample_prior(n, method='basic')
    assert [k for k, _ in rec.draws] == ['randn']
    out['basic_noise'], out['basic_z'] = rec.draws[0][1], z

    # ---- OfficialRHVAESampler, training-time path, on the piece of the pythae model it touches
    ref_loader.flow_modules()
    off = ref_loader._load('_ref_samplers.rhvae_sampler', 'src/models/samplers/rhvae_sampler.py', package='_ref_samplers')
    s = off.OfficialRHVAESampler(model)

    class _Mini:                       # what setup_official_rhvae leaves on self._rhvae_model (:83-101)
        centroids_tens, M_tens, latent_dim = c, M, 16
        temperature = torch.as_tensor(0.1)
        lbd = torch.as_tensor(lam)

    mini = _Mini()

    def g_inv(zz):                     # verbatim structure of the closure at rhvae_sampler.py:88-93
        diff = mini.centroids_tens.unsqueeze(0) - zz.unsqueeze(1)
        weights = torch.exp(-torch.norm(diff, dim=-1) ** 2 / (mini.temperature ** 2))
        weighted = mini.M_tens.unsqueeze(0) * weights.unsqueeze(-1).unsqueeze(-1)
        return weighted.sum(dim=1) + mini.lbd * torch.eye(mini.latent_dim)

    mini.G_inv = g_inv
    s._rhvae_model = mini              # skips setup_official_rhvae (needs encoder / decoder: out of scope)
    mu = torch.cat([c[:16] + 0.02 * torch.randn(16, 16, generator=torch.Generator().manual_seed(7)),
                    near_centroids(c, 16, 0.5, 71)])
    log_var = -1.0 + 0.3 * torch.randn(32, 16, generator=torch.Generator().manual_seed(72))
    torch.manual_seed(64)
    with ref_loader.RecordingRNG() as rec, _quiet():
        z = s.sample_riemannian_latents(mu, log_var, method='official')
    assert [k for k, _ in rec.draws] == ['randn_like']
    out['off_mu'], out['off_log_var'], out['off_eps'], out['off_z'] = mu, log_var, rec.draws[0][1], z
    _save('priors_synth_d16_k300', **out)


if __name__ == '__main__':
    main()
