"""Pin the CPU oracle: oracle/metric_oracle.py vs the committed outputs of the
REAL reference (tests/golden/*.npz, made by oracle/make_golden.py), vs the closed
form of the reference test's identity fixture, and -- when /root/reference
exists -- vs the live reference on fresh inputs."""
import math

import pytest
import torch

from conftest import load_golden, rel_fro, tables_of
from oracle import metric_oracle as O
from oracle import ref_loader

METRIC_CASES = ['ident_k10_T01', 'metricpt_T07', 'metricpt_T30', 'scaled_T07', 'synth_d16_k300',
                'synth_d8_k100', 'synth_d2_k20', 'synth_d32_k64', 'nonsym_d16_k48']


@pytest.mark.parametrize('case', METRIC_CASES)
def test_metric_functions_match_reference_outputs(case):
    g = load_golden(case)
    t = tables_of(g)
    z = g['z']
    # same torch ops in the same order on the same CPU -> agreement to rounding
    assert rel_fro(O.inverse_metric(z, *t), g['G_inv']) < 1e-6
    assert rel_fro(O.metric(z, *t), g['G']) < 1e-6
    torch.testing.assert_close(O.log_det_metric(z, *t), g['logdet_G'], rtol=1e-6, atol=1e-5)
    torch.testing.assert_close(O.hmc_log_pi(z, *t), g['log_pi'], rtol=1e-6, atol=1e-5)
    assert rel_fro(O.hmc_grad_modular(z, *t), g['grad_modular']) < 1e-5
    assert rel_fro(O.hmc_grad_modular_closed_form(z, *t), g['grad_modular']) < 1e-4
    if 'riem_dist2' in g:
        torch.testing.assert_close(O.riemannian_distance_squared(z, g['z2'], *t), g['riem_dist2'],
                                   rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case', METRIC_CASES)
def test_gradients_match_reference_autograd(case):
    g = load_golden(case)
    c, M, T, lam = tables_of(g)
    z = g['z']
    # grad log det G by autograd on the oracle == reference autograd
    assert rel_fro(O.grad_log_det_metric_autograd(z, c, M, T, lam), g['grad_logdet_G']) < 1e-4
    # closed form (variant D) == -1/2 grad log det G, and == grad of log_pi where unclamped
    closed = O.grad_log_sqrt_det_ginv_exact(z, c, M, T, lam)
    assert rel_fro(-2.0 * closed, g['grad_logdet_G']) < 2e-4
    unclamped = g['log_pi'] > 0.5 * math.log(1e-10) + 1e-3
    if unclamped.any():
        assert rel_fro(closed[unclamped], g['grad_log_pi'][unclamped]) < 2e-4
    # backward of G^{-1} for arbitrary upstream U
    assert rel_fro(O.metric_backward(z, c, M, T, g['U']), g['grad_ginv_U']) < 1e-4


def test_identity_fixture_closed_form():
    """tests/test_modular_components.py:68-73: M_k = I  =>  G^{-1} = (sum_k w_k + lambda) I,
    log det G = -d log(sum_k w_k + lambda)."""
    g = load_golden('ident_k10_T01')
    c, M, T, lam = tables_of(g)
    z = g['z'].double()
    w = torch.exp(-((c.double()[None] - z[:, None]) ** 2).sum(-1) / T ** 2).sum(1)
    want = -16 * torch.log(w + lam)
    torch.testing.assert_close(O.log_det_metric(g['z'], c, M, T, lam).double(), want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g['logdet_G'].double(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('case', ['hmc_d16_k300', 'hmc_d16_k300_beta03'])
def test_hmc_chain_matches_reference(case):
    g = load_golden(case)
    rec = {}
    z = O.hmc_sample(tables_of(g), g['z0'], g['gamma'], g['acc'], int(g['n_lf']),
                     float(g['eps_lf']), float(g['beta_zero']), record=rec)
    torch.testing.assert_close(z, g['z_final'], rtol=1e-5, atol=1e-5)
    assert len(rec['moves']) == g['gamma'].shape[0]


@pytest.mark.parametrize('case', ['samplers_metricpt_T07', 'samplers_synth_d16_k300'])
def test_samplers_match_reference(case):
    g = load_golden(case)
    t = tables_of(g)
    mu, lv = g['mu'], g['log_var']
    idx, dist = O.nearest2(mu, t[0])
    assert torch.equal(idx, g['near_idx'])
    torch.testing.assert_close(dist, g['near_dist'])
    torch.testing.assert_close(O.sample_enhanced(mu, lv, g['enhanced_eps'], t), g['enhanced_z'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.sample_geodesic(mu, lv, g['geodesic_eps'], g['geodesic_t'], t), g['geodesic_z'], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(O.sample_basic(mu, lv, g['basic_eps'], t), g['basic_z'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.sample_geodesic_prior(g['prior_idx1'], g['prior_idx2'], g['prior_t'], g['prior_eps'], t),
                               g['prior_z'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.hmc_refine(mu, lv, g['refine_eps'], t), g['refine_z'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.hmc_sample_posterior(mu[:8], lv[:8], g['post_eps0'], g['post_gamma'], t),
                               g['post_z'], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')
def test_oracle_vs_live_reference_fresh_inputs():
    from rlvae_b200.synthetic import make_synthetic_metric
    sm = make_synthetic_metric(128, 16, seed=9)
    t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    mt = ref_loader.make_ref_metric(*t)
    z = torch.randn(40, 16, generator=torch.Generator().manual_seed(77))
    assert rel_fro(O.inverse_metric(z, *t), mt.compute_inverse_metric(z)) < 1e-6
    assert rel_fro(O.metric(z, *t), mt.compute_metric(z)) < 1e-6
    torch.testing.assert_close(O.log_det_metric(z, *t), mt.compute_log_det_metric(z), rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize('case', ['builder_d16_T01', 'builder_d16_T05', 'builder_d2_T03', 'builder_d32_T10'])
def test_metric_construction_oracle_matches_reference_lines(case):
    """oracle.build_local_metrics vs the output of the reference's own loop
    (scripts/train_and_extract_vanilla_vae.py:204-226, executed by oracle/make_golden_builder.py)."""
    g = load_golden(case)
    got = O.build_local_metrics(g['latents'], g['centroids'], float(g['temperature']), float(g['regularization']))
    torch.testing.assert_close(got, g['M'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('case', ['rhvae_hmc_d16_k120', 'rhvae_hmc_d16_k120_beta03'])
def test_pythae_hmc_oracle_matches_reference_chain(case):
    """oracle.rhvae_hmc_sample vs the final state of the real RHVAESampler.hmc_sampling
    (oracle/make_golden_rhvae.py) under the recorded RNG stream."""
    g = load_golden(case)
    t = tables_of(g)
    z = O.rhvae_hmc_sample(t, g['idx0'], g['gamma'], g['acc'], int(g['n_lf']), float(g['eps_lf']),
                           float(g['beta_zero']))
    torch.testing.assert_close(z, g['z_final'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('case', ['losses_d16_k300', 'losses_d16_k10k'])
def test_loss_oracle_matches_reference_losses(case):
    """A21: the restated loss arithmetic (oracle/losses_oracle.py) on the CPU oracle metric reproduces the
    real reference's LossManager.compute_riemannian_kl_loss and the monolith KLs -- value and gradients
    w.r.t. mu / log_var (goldens: oracle/make_golden_losses.py)."""
    from oracle import losses_oracle as LO
    from rlvae_b200.synthetic import make_synthetic_metric
    g = load_golden(case)
    if 'centroids' in g:
        t = tables_of(g)
    else:
        sm = make_synthetic_metric(int(g['n_centroids']), 16, seed=int(g['table_seed']))
        t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    metric = LO.OracleMetric(*t)
    fns = {'modular_kl': lambda m, lv, z: LO.modular_riemannian_kl(m, lv, z, metric),
           'mono_metric_kl': lambda m, lv, z: LO.monolith_metric_kl(m, lv, z, metric.compute_metric),
           'mono_kl': lambda m, lv, z: LO.monolith_riemannian_kl(m, lv, z, metric.compute_metric)}
    for tag, fn in fns.items():
        m = g['mu'].clone().requires_grad_(True)
        lv = g['log_var'].clone().requires_grad_(True)
        z = m + g['eps'] * torch.exp(0.5 * lv)
        loss = fn(m, lv, z)
        loss.backward()
        assert abs(float(loss) - float(g[tag + '_loss'])) <= 1e-6 * (1 + abs(float(g[tag + '_loss']))), tag
        torch.testing.assert_close(m.grad, g[tag + '_dmu'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lv.grad, g[tag + '_dlogvar'], rtol=1e-5, atol=1e-6)
