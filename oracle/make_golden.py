"""Generate tests/golden/*.npz by running the REAL reference code on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference is Python, so it cannot travel to the GPU box; its inputs and
outputs travel instead, as small .npz fixtures, together with this script.
Each fixture stores the tables, the points, every random draw the reference
made (recorded in call order) and the reference's outputs.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from rlvae_b200.synthetic import make_synthetic_metric  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **{k: _np(v) for k, v in arrs.items()})
    print('wrote', name, {k: tuple(_np(v).shape) for k, v in arrs.items()})


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def metric_case(name, c, M, T, lam, z, z2=None):
    """A2/A3/A4/A5/A6/A8/A9/A20 outputs of the reference on (tables, z)."""
    mods = ref_loader.modules()
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    model = ref_loader.RefModel(mt)
    with _quiet():
        hmc = mods['hmc_sampler'].RiemannianHMCSampler(model, mcmc_steps_nbr=1, n_lf=1)
    ginv = mt.compute_inverse_metric(z)
    g = mt.compute_metric(z)
    ld = mt.compute_log_det_metric(z)
    logpi = hmc.log_pi(z)
    grad_a = hmc.grad_func(z).detach()
    zz = z.clone().requires_grad_(True)
    grad_ld = torch.autograd.grad(mt.compute_log_det_metric(zz).sum(), zz)[0]
    zz = z.clone().requires_grad_(True)
    grad_lp = torch.autograd.grad(hmc.log_pi(zz).sum(), zz)[0]
    out = dict(centroids=c, matrices=M, temperature=np.float64(T), regularization=np.float64(lam),
               z=z, G_inv=ginv, G=g, logdet_G=ld, log_pi=logpi, grad_modular=grad_a,
               grad_logdet_G=grad_ld, grad_log_pi=grad_lp)
    if z2 is not None:
        out['z2'] = z2
        out['riem_dist2'] = mt.compute_riemannian_distance_squared(z, z2)
    # backward of G^{-1} for a random upstream gradient (autograd through A2)
    gU = torch.Generator().manual_seed(11)
    U = torch.randn(ginv.shape, generator=gU)
    zz = z.clone().requires_grad_(True)
    (mt.compute_inverse_metric(zz) * U).sum().backward()
    out['U'] = U
    out['grad_ginv_U'] = zz.grad
    _save(name, **out)


def near_centroids(c, n, scale, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, c.shape[0], (n,), generator=g)
    return c[idx] + scale * torch.randn(n, c.shape[1], generator=g)


def hmc_case(name, c, M, T, lam, n, mcmc, n_lf, eps, beta_zero, seed):
    mods = ref_loader.modules()
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    model = ref_loader.RefModel(mt)
    with _quiet():
        s = mods['hmc_sampler'].RiemannianHMCSampler(model, mcmc_steps_nbr=mcmc, n_lf=n_lf,
                                                     eps_lf=eps, beta_zero=beta_zero)
    torch.manual_seed(seed)
    with ref_loader.RecordingRNG() as rec:
        zf = s.sample(n)
    kinds = [k for k, _ in rec.draws]
    assert kinds == ['randn'] + ['randn_like', 'rand'] * mcmc, kinds
    z0 = rec.draws[0][1]
    gam = torch.stack([rec.draws[1 + 2 * i][1] for i in range(mcmc)])
    acc = torch.stack([rec.draws[2 + 2 * i][1] for i in range(mcmc)])
    _save(name, centroids=c, matrices=M, temperature=np.float64(T), regularization=np.float64(lam),
          z0=z0, gamma=gam, acc=acc, n_lf=np.int64(n_lf), eps_lf=np.float64(eps),
          beta_zero=np.float64(beta_zero), z_final=zf)


def sampler_case(name, c, M, T, lam, n, seed):
    mods = ref_loader.modules()
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    model = ref_loader.RefModel(mt)
    ws = mods['riemannian_sampler'].WorkingRiemannianSampler(model)
    with _quiet():
        hs = mods['hmc_sampler'].RiemannianHMCSampler(model)
    g = torch.Generator().manual_seed(seed)
    mu = near_centroids(c, n, 0.5, seed + 1)
    log_var = -1.0 + 0.3 * torch.randn(n, c.shape[1], generator=g)
    out = dict(centroids=c, matrices=M, temperature=np.float64(T), regularization=np.float64(lam),
               mu=mu, log_var=log_var)
    for method in ('enhanced', 'geodesic', 'basic'):
        torch.manual_seed(seed + 10)
        with ref_loader.RecordingRNG() as rec, _quiet():
            zs = ws.sample_riemannian_latents(mu, log_var, method=method)
        out[f'{method}_eps'] = rec.draws[0][1]
        if method == 'geodesic':
            assert [k for k, _ in rec.draws] == ['randn_like', 'rand']
            out['geodesic_t'] = rec.draws[1][1]
        out[f'{method}_z'] = zs
    torch.manual_seed(seed + 20)
    with ref_loader.RecordingRNG() as rec, _quiet():
        zp = ws.sample_prior(n, method='geodesic')
    assert [k for k, _ in rec.draws] == ['randint', 'randint', 'rand', 'randn_like'], [k for k, _ in rec.draws]
    out['prior_idx1'], out['prior_idx2'] = rec.draws[0][1], rec.draws[1][1]
    out['prior_t'], out['prior_eps'], out['prior_z'] = rec.draws[2][1], rec.draws[3][1], zp
    torch.manual_seed(seed + 30)
    with ref_loader.RecordingRNG() as rec, _quiet():
        zr = hs.sample_riemannian_latents(mu, log_var, method='hmc')
    out['refine_eps'], out['refine_z'] = rec.draws[0][1], zr
    torch.manual_seed(seed + 40)
    with ref_loader.RecordingRNG() as rec, _quiet():
        zq = hs.sample_riemannian_latents(mu[:8], log_var[:8], method='posterior_hmc')
    out['post_eps0'] = rec.draws[0][1]
    out['post_gamma'] = torch.stack([d for _, d in rec.draws[1:]])
    out['post_z'] = zq
    # nearest-2 as the samplers compute it (riemannian_sampler.py:58-67)
    dist = torch.norm(mu.unsqueeze(1) - c.unsqueeze(0), dim=-1)
    _, idx = torch.topk(dist, k=2, dim=-1, largest=False)
    out['near_idx'], out['near_dist'] = idx, torch.gather(dist, 1, idx)
    _save(name, **out)


def flow_case(name, seed):
    mods = ref_loader.flow_modules()
    torch.manual_seed(seed)
    fm = mods['flow_manager'].FlowManager(latent_dim=16, n_flows=3, flow_hidden_size=32,
                                          flow_n_blocks=2, flow_n_hidden=1)
    z0 = torch.randn(6, 16)
    with torch.no_grad():
        zs, lds = fm.apply_flows([z0], n_obs=6)   # reuses the last flow beyond n_flows
    out = {'z0': z0, 'z_seq': torch.stack(zs), 'log_dets': torch.stack(lds)}
    for k, v in fm.state_dict().items():
        out['sd::' + k] = v
    _save(name, **out)


def main():
    warnings.simplefilter('ignore')
    assert ref_loader.available(), 'needs /root/reference'
    mods = ref_loader.modules()
    torch.manual_seed(0)

    # (iii) the reference test's own synthetic fixture: K=10, d=16, M=I, T=0.1, lambda=0.01
    g = torch.Generator().manual_seed(100)
    c = torch.randn(10, 16, generator=g)
    M = torch.eye(16).repeat(10, 1, 1)
    z = near_centroids(c, 8, 0.05, 101)
    metric_case('ident_k10_T01', c, M, 0.1, 0.01, z, near_centroids(c, 8, 0.05, 102))

    # (iv) the real fixtures, through the reference MetricLoader, T overridden to 0.7
    loader = mods['metric_loader'].MetricLoader(device=torch.device('cpu'))
    with _quiet():
        d = loader.load_from_file(os.path.join(ref_loader.REF_ROOT, 'data/pretrained/metric.pt'),
                                  temperature_override=0.7)
    c, M = d['centroids'], d['metric_matrices']
    z = near_centroids(c, 16, 0.3, 103)
    metric_case('metricpt_T07', c, M, d['temperature'], d['regularization'], z,
                near_centroids(c, 16, 0.3, 104))
    with _quiet():
        d3 = loader.load_from_file(os.path.join(ref_loader.REF_ROOT, 'data/pretrained/metric.pt'),
                                   temperature_override=3.0)
    metric_case('metricpt_T30', c, M, d3['temperature'], d3['regularization'],
                torch.randn(16, 16, generator=torch.Generator().manual_seed(105)))
    with _quiet():
        d2 = loader.load_from_file(
            os.path.join(ref_loader.REF_ROOT, 'data/pretrained/metric_T0.7_scaled.pt'),
            temperature_override=0.7)
    c2, M2 = d2['centroids'], d2['metric_matrices']
    metric_case('scaled_T07', c2, M2, d2['temperature'], d2['regularization'],
                near_centroids(c2, 16, 0.4, 106), near_centroids(c2, 16, 0.4, 107))

    # calibrated synthetic tables at several latent dims
    for (K, dd, N, tag) in ((300, 16, 64, 'synth_d16_k300'), (100, 8, 32, 'synth_d8_k100'),
                            (20, 2, 16, 'synth_d2_k20'), (64, 32, 16, 'synth_d32_k64')):
        sm = make_synthetic_metric(K, dd, seed=0)
        gz = torch.Generator().manual_seed(1)
        z = torch.randn(N, dd, generator=gz)
        z2 = z + 0.2 * torch.randn(N, dd, generator=gz)
        metric_case(tag, sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization, z, z2)

    # a non-symmetric M table: the reference uses M as given
    sm = make_synthetic_metric(48, 16, seed=3)
    gns = torch.Generator().manual_seed(5)
    Mns = sm.metric_matrices + 0.02 * torch.randn(48, 16, 16, generator=gns) * sm.metric_matrices.abs().mean()
    metric_case('nonsym_d16_k48', sm.centroids, Mns, sm.temperature, sm.regularization,
                torch.randn(16, 16, generator=gns))

    # HMC (A11): default tempering (beta0=1) and an active schedule (beta0=0.3)
    sm = make_synthetic_metric(300, 16, seed=0)
    hmc_case('hmc_d16_k300', sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization,
             n=48, mcmc=3, n_lf=5, eps=0.03, beta_zero=1.0, seed=21)
    hmc_case('hmc_d16_k300_beta03', sm.centroids, sm.metric_matrices, sm.temperature,
             sm.regularization, n=48, mcmc=2, n_lf=4, eps=0.05, beta_zero=0.3, seed=22)

    # samplers (A12-A17) on the real fixture at T=0.7 and on synthetic tables
    sampler_case('samplers_metricpt_T07', c, M, 0.7, 0.01, n=24, seed=31)
    sampler_case('samplers_synth_d16_k300', sm.centroids, sm.metric_matrices, sm.temperature,
                 sm.regularization, n=24, seed=32)

    flow_case('flow_d16', seed=41)


if __name__ == '__main__':
    main()
