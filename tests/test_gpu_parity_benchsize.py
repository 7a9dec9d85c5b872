"""GPU parity at the BENCHMARK's own table sizes (VERDICT r1 "weak 1"): the tensor-core product path
against the CPU oracle -- not against the repo's own direct kernels -- on

* BASELINE.json configs[1]: K = 10,000, d = 16: the first 4,096 points of the bench input in
  <= 512-point chunks, every output of the step: G^{-1}, G, log det G, grad_z log det G;
* the same tables at the reference's small temperature T = 0.7 (hybrid weight mode);
* BASELINE.json configs[2]: HMC at K = 10,000, 256 chains x 2 MCMC iterations x 20 leapfrog steps against
  the chain the REAL reference produced (tests/golden/hmc_d16_k10k.npz, oracle/make_golden_k10k.py),
  free-running and teacher-forced, identical accept decisions;
* BASELINE.json configs[4]: d = 64, K = 50,000: 32 points in chunks of 4 (3.3 GB oracle intermediate).

Tolerances are north_star's: rel 1e-5 (per-matrix Frobenius) on G^{-1} and G, 1e-4 on log det and
gradients, identical HMC accept decisions for a fixed RNG stream.
"""
import pytest
import torch

from conftest import load_golden, rel_fro
from oracle import metric_oracle as O
from test_gpu_parity import TOL_LD, TOL_MAT, close_ld, dev, make_mt

pytestmark = pytest.mark.gpu


def _bench_tables(temperature=None):
    from rlvae_b200.synthetic import make_synthetic_metric
    sm = make_synthetic_metric(10000, 16, seed=0)
    T = sm.temperature if temperature is None else temperature
    return (sm.centroids, sm.metric_matrices, T, sm.regularization)


def _oracle_all(z, t, chunk):
    """G^{-1}, G, log det G, grad_z log det G the way the reference computes them, chunked to bound
    the [n,K,d,d] intermediate (metric_tensor.py:98-182; gradient = closed form of autograd, checked
    against autograd in tests/test_oracle_golden.py)."""
    ginv = O.chunked(O.inverse_metric, z, *t, chunk=chunk)
    g = torch.linalg.inv(ginv)
    ld = torch.linalg.slogdet(g).logabsdet
    grad = -2.0 * O.chunked(O.grad_log_sqrt_det_ginv_exact, z, *t, chunk=chunk)
    return ginv, g, ld, grad


def test_k10k_tensor_path_against_oracle_4096_points():
    from rlvae_b200.synthetic import make_points
    t = _bench_tables()
    z = make_points(1 << 20, 16, seed=1)[:4096].contiguous()      # the first 4,096 points of bench.py's input
    mt = make_mt(t, 'tensor')
    tab = mt._tables(dev())
    assert tab.tensor_auto and tab.symmetric and tab.weight_mode == 0
    ev = mt.evaluate(z.to(dev()), want_ginv=True, want_g=True, want_logdet=True, want_grad=True)
    ginv, g, ld, grad = _oracle_all(z, t, chunk=512)
    assert rel_fro(ev['ginv'].cpu(), ginv) < TOL_MAT
    assert rel_fro(ev['g'].cpu(), g) < TOL_MAT
    close_ld(ev['logdet_g'], ld)
    assert rel_fro(ev['grad_logdet_g'].cpu(), grad) < TOL_LD
    # the public single-output entry points take the same kernels
    assert rel_fro(mt.compute_inverse_metric(z[:512].to(dev())).cpu(), ginv[:512]) < TOL_MAT
    assert rel_fro(mt.compute_metric(z[:512].to(dev())).cpu(), g[:512]) < TOL_MAT
    close_ld(mt.compute_log_det_metric(z[:512].to(dev())), ld[:512])


def test_k10k_small_temperature_hybrid_against_oracle():
    """T = 0.7 (conf/model/hybrid_rlvae.yaml:41) at K = 10k: log det and gradient too, against the
    oracle.  Half of the points sit next to centroids (otherwise every weight underflows and
    G^{-1} = lambda I, which tests nothing)."""
    from rlvae_b200.synthetic import make_points
    t = _bench_tables(temperature=0.7)
    c = t[0]
    z = torch.cat([make_points(512, 16, seed=1), c[:512] + 0.05 * make_points(512, 16, seed=2)]).contiguous()
    mt = make_mt(t, 'auto')
    assert mt._tables(dev()).weight_mode == 2
    ev = mt.evaluate(z.to(dev()), want_ginv=True, want_g=True, want_logdet=True, want_grad=True)
    ginv, g, ld, grad = _oracle_all(z, t, chunk=256)
    assert rel_fro(ev['ginv'].cpu(), ginv) < TOL_MAT
    assert rel_fro(ev['g'].cpu(), g) < TOL_MAT
    close_ld(ev['logdet_g'], ld)
    live = grad.norm(dim=1) > 1e-6 * grad.norm(dim=1).max()
    assert live.sum() > 300
    assert rel_fro(ev['grad_logdet_g'].cpu()[live], grad[live]) < TOL_LD


@pytest.mark.parametrize('temperature', [None, 0.7, 0.1])
def test_k10k_pythae_gradient_tensor_path_against_oracle(temperature):
    """A8 at the bench table size: variant C, (1/T^2) G^T sum_k w_k M_k^T (c_k - z) (pythae rhvae_sampler.py:160-187),
    against the oracle at the default temperature, at T = 0.7 (hybrid weights) and at the T = 0.1
    OfficialRHVAESampler hard-codes, most points a fraction of T away from a centroid (where the chains live).

    Long batch (> 2048 points): forward kernel + unit-weight gradient kernel + finish; the finish kernel's error bound
    sends the rows where the table-contraction form sum_k w_k M_k c_k - (sum_k w_k M_k) z would lose digits (next to
    a centroid at small T: measured 1e-2 without it) to the per-centroid kernel.  Short batch: forward kernel +
    per-centroid kernel with the centroids split over CTAs.  CUDA-core path: per-centroid kernel throughout."""
    from rlvae_b200 import _capi
    from rlvae_b200.synthetic import make_points
    t = _bench_tables(temperature=temperature)
    c = t[0]
    T = t[2]
    near = c[:2048] + 0.3 * T * make_points(2048, 16, seed=2) / 4.0      # a fraction of T away from a centroid
    z = torch.cat([make_points(64, 16, seed=1), near]).contiguous()
    mt = make_mt(t, 'auto')
    tab = mt._tables(dev())
    zd = z.to(dev())
    ref = O.chunked(O.grad_pythae, z, *t, chunk=128).reshape(z.shape)
    live = ref.norm(dim=1) > 1e-3 * ref.norm(dim=1).max()
    assert live.sum() >= 2000
    got, lad, sgn = _capi.pythae_eval(tab, zd, path=_capi.PATH_TENSOR)               # long batch: split + bound
    assert rel_fro(got.cpu()[live], ref[live]) < TOL_LD
    assert (got.cpu() - ref).norm(dim=1).max() < 2e-5 * ref.norm(dim=1).max()        # every row, absolute
    sl = torch.linalg.slogdet(O.chunked(O.inverse_metric, z, *t, chunk=128).double())
    close_ld(lad, sl.logabsdet)
    assert torch.equal(sgn.cpu().double(), sl.sign)
    g = mt.compute_metric(zd)
    got2 = _capi.metric_grad_pythae(tab, zd, g, path=_capi.PATH_TENSOR)
    assert rel_fro(got2.cpu()[live], ref[live]) < TOL_LD
    sub = slice(32, 544)                                                             # short batch: 32 far + 480 near
    got3, _, _ = _capi.pythae_eval(tab, zd[sub].contiguous(), path=_capi.PATH_TENSOR)
    assert rel_fro(got3.cpu()[live[sub]], ref[sub][live[sub]]) < TOL_LD
    got4 = _capi.metric_grad_pythae(tab, zd[sub].contiguous(), g[sub].contiguous(), path=_capi.PATH_DIRECT)
    assert rel_fro(got4.cpu()[live[sub]], ref[sub][live[sub]]) < TOL_LD
    # the per-centroid kernel keeps the relative accuracy of small rows too
    small = ref[sub].norm(dim=1) > 1e-9 * ref.norm(dim=1).max()
    assert rel_fro(got4.cpu()[small], ref[sub][small]) < 5 * TOL_LD
    assert rel_fro(got3.cpu()[small], ref[sub][small]) < 5 * TOL_LD


def test_hmc_k10k_matches_the_reference_chain():
    """256 chains x 2 MCMC iterations x n_lf = 20 at K = 10,000 against RiemannianHMCSampler.sample of
    the real reference (golden): free-running final state, then per-iteration teacher forcing with
    identical accept decisions (a flip only where acc sits within rounding of alpha; count must be 0)."""
    from rlvae_b200 import MetricModel, RiemannianHMCSampler
    g = load_golden('hmc_d16_k10k')
    t = _bench_tables()
    assert int(g['n_centroids']) == t[0].shape[0]
    iters, n_lf = g['gamma'].shape[0], int(g['n_lf'])
    s = RiemannianHMCSampler(MetricModel(make_mt(t, 'auto')), mcmc_steps_nbr=iters, n_lf=n_lf,
                             eps_lf=float(g['eps_lf']), beta_zero=float(g['beta_zero']))
    D = lambda k: g[k].to(dev())
    zf = s.sample_with_streams(D('z0'), D('gamma'), D('acc'))
    # chains whose decisions all sat clear of alpha must end where the reference's ended
    clear = ((g['acc'] - g['rec_alpha']).abs() > 1e-4).all(dim=0)
    assert clear.float().mean() > 0.95
    torch.testing.assert_close(zf.cpu()[clear], g['z_final'][clear], rtol=2e-4, atol=2e-4)
    forced = [g['z0']] + [g['rec_z'][i] for i in range(iters - 1)]
    got = {}
    s.sample_with_streams(D('z0'), D('gamma'), D('acc'), z_forced=[f.to(dev()) for f in forced], record=got)
    mism = 0
    for i in range(iters):
        close_ld(got['H0'][i], g['rec_H0'][i], 1e-4)
        close_ld(got['H'][i], g['rec_H'][i], 1e-4)
        flip = got['moves'][i].cpu() != g['rec_moves'][i]
        assert torch.all((g['acc'][i][flip] - g['rec_alpha'][i][flip]).abs() < 1e-5)
        mism += int(flip.sum())
        torch.testing.assert_close(got['z'][i].cpu()[~flip], g['rec_z'][i][~flip], rtol=1e-4, atol=1e-4)
    assert mism == 0, f'{mism} accept decisions differ'


def test_d64_k50k_against_oracle():
    """BASELINE.json configs[4] tables (d = 64, K = 50,000): G^{-1}, log det G and grad_z log det G on
    32 points against the oracle in chunks of 4 (the reference's [n,K,d,d] intermediate is 3.3 GB per
    chunk), plus a ragged 129-point batch for tile tails against the first rows."""
    from rlvae_b200.synthetic import make_points, make_synthetic_metric
    sm = make_synthetic_metric(50000, 64, seed=5, n_probe=128)
    t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    z = make_points(129, 64, seed=6)
    mt = make_mt(t, 'tensor')
    assert mt._tables(dev()).tensor_auto
    ev = mt.evaluate(z.to(dev()), want_ginv=True, want_g=False, want_logdet=True, want_grad=True)
    zr = z[:32]
    ginv = O.chunked(O.inverse_metric, zr, *t, chunk=4)
    ld = -torch.linalg.slogdet(ginv).logabsdet
    grad = -2.0 * O.chunked(O.grad_log_sqrt_det_ginv_exact, zr, *t, chunk=4)
    assert rel_fro(ev['ginv'][:32].cpu(), ginv) < TOL_MAT
    close_ld(ev['logdet_g'][:32], ld)
    assert rel_fro(ev['grad_logdet_g'][:32].cpu(), grad) < TOL_LD
    one = mt.evaluate(z[:32].to(dev()), want_ginv=True, want_logdet=True, want_grad=True)
    assert rel_fro(one['ginv'].cpu(), ev['ginv'][:32].cpu()) < 1e-6


def test_fused_hmc_trajectory_is_one_launch_and_matches_the_per_step_path():
    """VERDICT r1 item 3 (north_star "one leapfrog-step kernel advances many HMC chains per launch"):
    on certified tables rlvae_hmc_iteration / rlvae_hmc_run are ONE kernel launch for the whole
    trajectory (counted inside the library), and the result equals the per-step path (forward kernel +
    element-wise stage per leapfrog step) -- same Hamiltonians, same decisions, same states -- for one
    iteration and for a free-running multi-iteration chain, at ragged sizes and at K = 10k."""
    from rlvae_b200 import MetricModel, RiemannianHMCSampler, _capi
    from rlvae_b200.synthetic import make_hmc_streams, make_synthetic_metric
    lib = _capi.lib()
    for K, n, iters, n_lf, beta0 in ((300, 777, 4, 7, 1.0), (300, 130, 3, 5, 0.3), (10000, 1000, 2, 20, 1.0),
                                     (37, 300, 3, 1, 1.0), (100, 64, 2, 2, 0.5), (200, 1, 5, 3, 1.0)):
        sm = make_synthetic_metric(K, 16, seed=0)
        t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
        mt = make_mt(t, 'auto')
        tab = mt._tables(dev())
        assert tab.psd_certified and _capi.hmc_fused_available(tab)
        z0, gam, acc = make_hmc_streams(n, 16, iters, seed=40 + n)
        z0, gam, acc = z0.to(dev()), gam.to(dev()), acc.to(dev())
        s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=iters, n_lf=n_lf, eps_lf=0.03, beta_zero=beta0)
        b0h = s.beta_zero_sqrt.detach().float().cpu()
        scales, _ = s._all_scales(iters, n_lf, b0h, b0h)
        b0 = float(s.beta_zero_sqrt.item())
        zf = z0.clone()
        torch.cuda.synchronize()
        lib.rlvae_launch_count(1)
        rf = _capi.hmc_run(tab, zf, gam, acc, n_lf, 0.03, b0, scales, want_stats=True, want_trace=True)
        assert lib.rlvae_launch_count(0) == 1, 'the fused path must be a single kernel launch'
        assert int(_capi.hmc_fail_count(rf['work'], n, 16).item()) == 0
        zu = z0.clone()
        lib.rlvae_launch_count(1)
        ru = _capi.hmc_run(tab, zu, gam, acc, n_lf, 0.03, b0, scales, _capi.GRAD_MODULAR | _capi.HMC_NO_FUSION,
                           want_stats=True, want_trace=True)
        assert lib.rlvae_launch_count(0) >= iters * (n_lf + 1)
        # one-iteration entry point: also a single launch
        z1 = z0.clone()
        lib.rlvae_launch_count(1)
        st1 = _capi.hmc_iteration(tab, z1, gam[0], acc[0], n_lf, 0.03, b0, scales[:n_lf], want_stats=True)
        assert lib.rlvae_launch_count(0) == 1
        torch.testing.assert_close(z1, rf['trace'][0], rtol=0, atol=0)
        torch.testing.assert_close(st1[0], rf['stats'][0][0], rtol=0, atol=0)
        # iteration 0 starts from identical states: decisions may differ only for acc within rounding of alpha
        flip0 = rf['stats'][3][0] != ru['stats'][3][0]
        assert torch.all((acc[0][flip0] - ru['stats'][2][0][flip0]).abs() < 1e-5)
        for name, a, b in zip(('H0', 'H', 'alpha'), rf['stats'][:3], ru['stats'][:3]):
            close_ld(a[0], b[0], 1e-4)
        same = ~flip0
        torch.testing.assert_close(rf['trace'][0][same], ru['trace'][0][same], rtol=1e-4, atol=1e-5)
        # free-running: chains whose decisions never sat within 1e-4 of alpha end in the same state
        clear = ((acc - ru['stats'][2]).abs() > 1e-4).all(dim=0)
        assert clear.float().mean() > 0.9 or n < 8
        torch.testing.assert_close(zf[clear], zu[clear], rtol=2e-4, atol=2e-4)
        assert torch.equal(rf['stats'][3][:, clear], ru['stats'][3][:, clear])


@pytest.mark.parametrize('case', ['losses_d16_k300', 'losses_d16_k10k'])
def test_losses_through_the_drop_in_match_the_reference(case):
    """A21 (SURVEY.md 8f rank 1's purpose): the reference's loss arithmetic -- LossManager.
    compute_riemannian_kl_loss (loss_manager.py:75-146) and the monolith KLs (riemannian_flow_vae.py:
    1004-1077, 1328-1394), restated in oracle/losses_oracle.py and pinned against the real reference --
    driven through the CUDA drop-in MetricTensor: loss value and the gradients w.r.t. mu and log_var
    (autograd through rlvae_metric_grad / the batched inverse) against the reference's own numbers."""
    from oracle import losses_oracle as LO
    from rlvae_b200.synthetic import make_synthetic_metric
    g = load_golden(case)
    if 'centroids' in g:
        t = (g['centroids'], g['matrices'], float(g['temperature']), float(g['regularization']))
    else:
        sm = make_synthetic_metric(int(g['n_centroids']), 16, seed=int(g['table_seed']))
        t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    mt = make_mt(t, 'auto').to(dev())        # LossManager calls .to(device) on it (loss_manager.py:106-107)
    fns = {'modular_kl': lambda m, lv, z: LO.modular_riemannian_kl(m, lv, z, mt),
           'mono_metric_kl': lambda m, lv, z: LO.monolith_metric_kl(m, lv, z, mt.compute_metric),
           'mono_kl': lambda m, lv, z: LO.monolith_riemannian_kl(m, lv, z, mt.compute_metric)}
    for tag, fn in fns.items():
        m = g['mu'].to(dev()).requires_grad_(True)
        lv = g['log_var'].to(dev()).requires_grad_(True)
        z = m + g['eps'].to(dev()) * torch.exp(0.5 * lv)
        loss = fn(m, lv, z)
        loss.backward()
        ref = float(g[tag + '_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * (1 + abs(ref)), (tag, float(loss.detach()), ref)
        for got, want in ((m.grad, g[tag + '_dmu']), (lv.grad, g[tag + '_dlogvar'])):
            err = (got.cpu().double() - want.double()).norm() / want.double().norm().clamp_min(1e-30)
            assert err < 1e-4, (tag, float(err))


def test_prior_samplers_and_official_sampler_match_the_reference():
    """A17 (riemannian_sampler.py:222-355) with the reference's recorded draws injected -- centroid_aware,
    weighted_mixture (draws in the reference's per-component call order) and basic, next to the geodesic
    prior of test_samplers_match_reference -- and the training-time path of OfficialRHVAESampler
    (src/models/samplers/rhvae_sampler.py:108-167: temperature hard-coded to 0.1) against the outputs of the
    real reference classes (oracle/make_golden_priors.py)."""
    from rlvae_b200 import MetricModel, OfficialRHVAESampler, WorkingRiemannianSampler
    from rlvae_b200.synthetic import make_synthetic_metric
    g = load_golden('priors_synth_d16_k300')
    sm = make_synthetic_metric(int(g['n_centroids']), 16, seed=int(g['table_seed']))
    t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    model = MetricModel(make_mt(t, 'auto'))
    ws = WorkingRiemannianSampler(model)
    D = lambda k: g[k].to(dev())
    torch.testing.assert_close(ws.centroid_aware_prior_with_noise(D('ca_idx'), D('ca_noise')).cpu(), g['ca_z'],
                               rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ws.weighted_mixture_prior_with_noise(D('wm_idx'), D('wm_noise')).cpu(), g['wm_z'],
                               rtol=1e-6, atol=1e-6)
    assert torch.equal(g['basic_noise'], g['basic_z'])        # the reference's basic prior IS its randn draw
    for m in ('geodesic', 'centroid_aware', 'weighted_mixture', 'basic'):
        z = ws.sample_prior(33, method=m)
        assert z.shape == (33, 16) and torch.isfinite(z).all()
    # weighted mixture: every sample sits within a few noise standard deviations of ITS component's centroid
    comp = torch.randint(0, 300, (500,), device=dev())
    zz = ws.weighted_mixture_prior_with_noise(comp, torch.randn(500, 16, device=dev()))
    assert ((zz - model.centroids_tens[comp]).norm(dim=1) < 0.1 * 16 ** 0.5 * 3).all()
    # OfficialRHVAESampler: name, surface, hard-coded temperature, training-time path
    off = OfficialRHVAESampler(model)
    assert set(off.get_sampling_methods()) == {'official', 'standard'}
    z = off.official_with_noise(D('off_mu'), D('off_log_var'), D('off_eps'))
    assert abs(off.get_rhvae_info()['rhvae_temperature'] - 0.1) < 1e-7
    torch.testing.assert_close(z.cpu(), g['off_z'], rtol=2e-5, atol=2e-5)
    # differentiable w.r.t. mu and log_var like the reference's (:143-148)
    mu = D('off_mu').clone().requires_grad_(True)
    lv = D('off_log_var').clone().requires_grad_(True)
    off.official_with_noise(mu, lv, D('off_eps')).sum().backward()
    assert torch.isfinite(mu.grad).all() and torch.isfinite(lv.grad).all() and lv.grad.abs().sum() > 0
    assert off.sample_riemannian_latents(D('off_mu'), D('off_log_var')).shape == (32, 16)
    assert off.sample_riemannian_latents(D('off_mu'), D('off_log_var'), method='standard').shape == (32, 16)
    # prior: batches of at most 32 (:186), 100 MCMC steps x 15 leapfrog; latents returned
    off._rhvae_sampler.mcmc_steps_nbr = 3                      # keep the test short: the loop itself is pinned by
    zp = off.sample_prior(40)                                  # test_pythae_variant_hmc_matches_reference_chain
    assert zp.shape == (40, 16) and torch.isfinite(zp).all()
    assert off.sample_prior(5, method='basic').shape == (5, 16)
    # the 32-chain batches run as ONE library call with the draws made in the sequential loop's order: same generator
    # state -> the samples of the batch-by-batch loop (pythae RHVAESampler.sample :61-67)
    off._rhvae_sampler.mcmc_steps_nbr = 4
    torch.manual_seed(1234)
    merged = off.sample_prior(100)
    torch.manual_seed(1234)
    seq = torch.cat([off._rhvae_sampler.hmc_sampling(b) for b in (32, 32, 32, 4)])
    assert merged.shape == (100, 16)
    same = ((merged - seq).abs().max(dim=1).values < 1e-4)
    assert same.float().mean() >= 0.97, same.float().mean()       # (a borderline accept may flip a chain)


def test_d64_tensor_gradient_kernel_general_u_and_tile_tails():
    """The column-tiled tcgen05 gradient kernel at d = 64 (rlvae_metric_grad_ws) with an ARBITRARY
    (non-symmetric, badly scaled) U -- the autograd backward of compute_inverse_metric -- against the oracle
    formula (2/T^2) sum_k w_k <U, M_k> (c_k - z), at ragged batch sizes, and against the CUDA-core kernel."""
    from rlvae_b200 import _capi
    from rlvae_b200.synthetic import make_points, make_synthetic_metric
    sm = make_synthetic_metric(700, 64, seed=9)
    t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    mt = make_mt(t, 'tensor')
    tab = mt._tables(dev())
    assert 'gradient: partial tiles' in mt.kernel_info()['implementation']
    gen = torch.Generator().manual_seed(10)
    for n in (1, 127, 130, 300):
        z = make_points(n, 64, seed=11 + n)
        U = torch.randn(n, 64, 64, generator=gen) * torch.logspace(-6, 6, n)[:, None, None]
        ref = O.chunked(lambda zz, uu: O.metric_backward(zz, *t[:3], uu), torch.arange(n), chunk=n) if False else \
            torch.cat([O.metric_backward(z[i:i + 16], *t[:3], U[i:i + 16]) for i in range(0, n, 16)])
        lib = _capi.lib()
        lib.rlvae_launch_count(1)
        got = _capi.metric_grad(tab, z.to(dev()), U.to(dev()), 2.0 / sm.temperature ** 2, _capi.PATH_TENSOR)
        assert lib.rlvae_launch_count(0) == 3          # per-point scale, tensor kernel, tile reduction
        assert rel_fro(got.cpu(), ref) < TOL_LD, n
        direct = _capi.metric_grad(tab, z.to(dev()), U.to(dev()), 2.0 / sm.temperature ** 2, _capi.PATH_DIRECT)
        assert rel_fro(got.cpu(), direct.cpu()) < TOL_LD, n
    # through autograd
    zg = make_points(64, 64, seed=3).to(dev()).requires_grad_(True)
    w = torch.randn(64, 64, 64, generator=gen).to(dev())
    (mt.compute_inverse_metric(zg) * w).sum().backward()
    ref = O.metric_backward(zg.detach().cpu(), *t[:3], w.cpu())
    assert rel_fro(zg.grad.cpu(), ref) < TOL_LD


@pytest.mark.parametrize('d,K', [(16, 300), (64, 300)])
def test_gradient_rows_of_far_points_keep_relative_accuracy(d, K):
    """The gradient kernels form u = w t and contract it on the tensor core (d = 16: 3xTF32, fp32 exponent range;
    d = 64: kind::f16).  A point far from every centroid has uniformly tiny weights (down to e^-36 here): with a
    fixed fp16 scale its u would sit in the subnormals and the RELATIVE accuracy of its row would be lost, so
    the d = 64 kernel scales u per (point, block) by a power of two.  Rows at increasing distance, arbitrary U,
    against the oracle formula."""
    from rlvae_b200 import _capi
    from rlvae_b200.synthetic import make_points, make_synthetic_metric
    sm = make_synthetic_metric(K, d, seed=13)
    T = 0.9 if d == 16 else sm.temperature     # d = 64: the tensor path is only selected where the expanded distance form is accurate
    t = (sm.centroids, sm.metric_matrices, T, sm.regularization)
    mt = make_mt(t, 'auto')
    tab = mt._tables(dev())
    assert tab.tensor_auto
    gen = torch.Generator().manual_seed(14)
    base = sm.centroids[torch.arange(160) % K]
    # (from 0.3 T outwards: exactly ON a centroid its own term vanishes and the row is a cancellation residue)
    shift = torch.linspace(0.3, 6.0, 160)[:, None] * torch.nn.functional.normalize(torch.randn(160, d, generator=gen), dim=1)
    z = (base + shift * T).contiguous()
    U = torch.randn(160, d, d, generator=gen)
    ref = torch.cat([O.metric_backward(z[i:i + 16].double(), t[0].double(), t[1].double(), T, U[i:i + 16].double())
                     for i in range(0, 160, 16)]).float()
    got = _capi.metric_grad(tab, z.to(dev()), U.to(dev()), 2.0 / T ** 2, _capi.PATH_AUTO).cpu()
    rows = ref.norm(dim=1)
    assert rows.min() < 1e-3 * rows.max()                      # the far rows really are tiny in absolute terms
    live = rows > 0
    err = ((got.double() - ref.double()).norm(dim=1) / rows.double().clamp_min(1e-300))[live]
    assert err.max() < TOL_LD, float(err.max())
