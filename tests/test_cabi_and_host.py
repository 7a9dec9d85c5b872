"""CPU-side tests: the C-ABI library loads and exports every declared symbol (no compute
without a GPU), the host mirrors validate like the reference, the loader's key fallbacks."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from rlvae_b200 import _capi, build
    build.build()
    h = _capi.lib()
    hdr = open(os.path.join(ROOT, 'include', 'rlvae_b200.h')).read()
    declared = set(re.findall(r'\b(rlvae_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'rlvae_tables'}
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(h, name), name
    assert h.rlvae_abi_version() == 2
    # argument validation happens before any CUDA call: safe without a GPU
    assert h.rlvae_inverse_metric(None, None, 4, None, None, 0, None) != 0
    assert b'not loaded' in h.rlvae_last_error()
    assert h.rlvae_metric_eval_workspace(10, 16) == 4 * (3 * 10 * 256 + 10 + 4)


def test_no_product_module_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'rlvae_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f


def test_metric_tensor_validation_matches_reference_errors():
    from rlvae_b200 import MetricTensor
    mt = MetricTensor(latent_dim=16, device=torch.device('cpu'))
    assert not mt.is_loaded() and mt.get_config()['n_centroids'] == 0
    with pytest.raises(RuntimeError, match='Metric tensor not loaded'):
        mt.compute_inverse_metric(torch.zeros(2, 16))
    with pytest.raises(ValueError, match='latent_dim'):
        mt.load_pretrained(torch.zeros(3, 8), torch.zeros(3, 8, 8))
    with pytest.raises(ValueError, match='Number of metric matrices'):
        mt.load_pretrained(torch.zeros(3, 16), torch.zeros(4, 16, 16))
    with pytest.raises(ValueError, match='Metric matrix shape'):
        mt.load_pretrained(torch.zeros(3, 16), torch.zeros(3, 16, 8))
    mt.load_pretrained(torch.zeros(3, 16), torch.eye(16).repeat(3, 1, 1), temperature=0.7)
    cfg = mt.get_config()
    assert cfg['is_loaded'] and cfg['n_centroids'] == 3 and abs(cfg['temperature'] - 0.7) < 1e-6
    assert set(mt.state_dict()) == {'centroids', 'metric_matrices', 'temperature', 'regularization'}
    # no CPU fallback: a CPU tensor is an error, not a silent eager path
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        mt.compute_inverse_metric(torch.zeros(2, 16))


def test_metric_loader_key_fallbacks(tmp_path):
    from rlvae_b200 import MetricLoader
    ld = MetricLoader(device=torch.device('cpu'))
    c = torch.randn(5, 4)
    m = torch.eye(4).repeat(5, 1, 1)
    torch.save({'metric_centroids': c, 'metric_vars': m, 'metric_temperature': 0.7}, tmp_path / 'a.pt')
    with pytest.warns(UserWarning):
        d = ld.load_from_file(tmp_path / 'a.pt')     # 'metric_temperature' is unknown -> default 0.1
    assert d['temperature'] == 0.1 and d['regularization'] == 0.01 and torch.equal(d['centroids'], c)
    torch.save({'mu': c.tolist(), 'M_i_flat': torch.ones(5, 4), 'T': torch.tensor(0.5), 'lbd': 0.02},
               tmp_path / 'b.pt')
    d = ld.load_from_file(tmp_path / 'b.pt', regularization_override=0.03)
    assert d['temperature'] == 0.5 and d['regularization'] == 0.03
    assert torch.equal(d['metric_matrices'], torch.diag_embed(torch.ones(5, 4)))
    torch.save({'centers': c}, tmp_path / 'c.pt')
    with pytest.warns(UserWarning):
        d = ld.load_from_file(tmp_path / 'c.pt')
    assert torch.equal(d['metric_matrices'][2], torch.eye(4))
    with pytest.raises(FileNotFoundError):
        ld.load_from_file(tmp_path / 'missing.pt')
    torch.save({'x': 1}, tmp_path / 'd.pt')
    with pytest.raises(ValueError, match='No centroids'):
        ld.load_from_file(tmp_path / 'd.pt')
    torch.save({'centroids': c, 'M_matrices': torch.zeros(5, 3, 3)}, tmp_path / 'e.pt')
    with pytest.raises(ValueError, match='shape'):
        ld.load_from_file(tmp_path / 'e.pt')
    ld.save_to_file(tmp_path / 'f.pt', c, m, 0.3, 0.05, metadata={'k': 1})
    d = ld.load_from_file(tmp_path / 'f.pt')
    assert d['temperature'] == 0.3 and d['regularization'] == 0.05
    rep = ld.validate_metric_file(tmp_path / 'f.pt')
    assert rep['valid'] and rep['n_centroids'] == 5 and not rep['has_negative_eigenvals']
    assert not ld.validate_metric_file(tmp_path / 'd.pt')['valid']
    ld.convert_old_format(tmp_path / 'a.pt', tmp_path / 'g.pt', temperature_override=0.7)
    assert ld.load_from_file(tmp_path / 'g.pt')['temperature'] == 0.7


def test_synthetic_metric_is_calibrated_and_deterministic():
    from rlvae_b200.synthetic import make_hmc_streams, make_synthetic_metric
    a = make_synthetic_metric(200, 16, seed=0)
    b = make_synthetic_metric(200, 16, seed=0)
    assert torch.equal(a.centroids, b.centroids) and torch.equal(a.metric_matrices, b.metric_matrices)
    assert torch.equal(a.metric_matrices, a.metric_matrices.transpose(1, 2))
    assert abs(a.temperature - 3.0) < 1e-12 and a.regularization == 0.01
    from oracle import metric_oracle as O
    z = torch.randn(4096, 16, generator=torch.Generator().manual_seed(5))
    ld = torch.linalg.slogdet(O.inverse_metric(z.double(), a.centroids.double(), a.metric_matrices.double(),
                                               a.temperature, a.regularization)).logabsdet
    assert abs(ld.mean().item()) < 1.0          # det G^{-1} has geometric mean ~1
    z0, g, acc = make_hmc_streams(10, 16, 3, seed=2)
    assert z0.shape == (10, 16) and g.shape == (3, 10, 16) and acc.shape == (3, 10)


def test_new_host_modules_refuse_cpu_tensors_and_keep_reference_surfaces():
    """metric_builder / RHVAE-style sampler: no CPU fallback, reference-compatible surfaces."""
    import inspect
    from rlvae_b200 import RHVAEStyleHMCSampler, metric_builder
    x = torch.randn(20, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        metric_builder.build_local_metrics(x, x[:3], 0.1, 0.01)
    with pytest.raises(ValueError):
        metric_builder.build_metric_data(x.cuda() if torch.cuda.is_available() else x)
    lf = torch.tril(torch.randn(3, 4, 4))
    assert torch.allclose(metric_builder.matrices_from_cholesky_factors(lf), lf @ lf.transpose(1, 2))
    sig = inspect.signature(RHVAEStyleHMCSampler.__init__)
    assert list(sig.parameters)[1:] == ['model', 'mcmc_steps_nbr', 'n_lf', 'eps_lf', 'beta_zero']
    assert sig.parameters['mcmc_steps_nbr'].default == 100 and sig.parameters['n_lf'].default == 15
    assert RHVAEStyleHMCSampler.tempering(5, 10, 0.5) == pytest.approx(1.0 / ((1 - 2.0) * 0.25 + 2.0))


def test_merged_prior_batches_draw_in_the_sequential_order():
    """OfficialRHVAESampler.sample_prior runs the reference's 32-chain batches as one library call; what ties a
    chain to its batch is only the order of the random draws (pythae RHVAESampler.sample :61-67 -> hmc_sampling
    :100, :107, :141).  Host logic, checked on the CPU generator: the stacked streams are exactly the streams the
    batch-by-batch loop draws."""
    import torch.nn as nn
    from rlvae_b200.samplers.rhvae_sampler import RHVAEStyleHMCSampler

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.p = nn.Parameter(torch.zeros(1))
            self.latent_dim = 5
            self.centroids_tens = torch.arange(70.0).reshape(14, 5)

    class Recording(RHVAEStyleHMCSampler):
        def __init__(self, model):
            super().__init__(model, mcmc_steps_nbr=3)
            self.calls = []

        def hmc_sampling_with_streams(self, idx0, gammas, accs, record=None, z_start=None, state=None):
            self.calls.append((None if idx0 is None else idx0.clone(), gammas.clone(), accs.clone()))
            return self.model.centroids_tens[idx0] if z_start is None else z_start

    sizes = [4, 4, 3]
    merged = Recording(Model())
    torch.manual_seed(7)
    out = merged.hmc_sampling_batches(sizes)
    assert out.shape == (11, 5) and len(merged.calls) == 1
    seq = Recording(Model())
    torch.manual_seed(7)
    for b in sizes:
        seq.hmc_sampling(b)
    assert len(seq.calls) == 3
    idx_m, gam_m, acc_m = merged.calls[0]
    assert torch.equal(idx_m, torch.cat([c[0] for c in seq.calls]))
    assert torch.equal(gam_m, torch.cat([c[1] for c in seq.calls], dim=1))
    assert torch.equal(acc_m, torch.cat([c[2] for c in seq.calls], dim=1))
    # a single batch goes through hmc_sampling itself; empty input is an empty result
    one = Recording(Model())
    assert one.hmc_sampling_batches([6]).shape == (6, 5) and one.hmc_sampling_batches([]).shape == (0, 5)


def test_zero_padded_path_is_off_where_it_does_not_apply():
    """MetricTensor._embedded: only for loaded CUDA tables with latent_dim not in {16, 64}, K >= 512, lambda > 0 and
    kernel_path != 'direct' -- everything else keeps the native tables (no CUDA call is made for these answers)."""
    from rlvae_b200 import MetricTensor
    cpu = torch.device('cpu')
    assert MetricTensor(16, device=cpu)._embedded(cpu) is None
    assert MetricTensor(10, device=cpu)._embedded(cpu) is None                       # not loaded
    mt = MetricTensor(10, device=cpu, kernel_path='direct')
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(torch.zeros(600, 10), torch.eye(10).repeat(600, 1, 1), temperature=1.0, regularization=0.01)
    assert mt._embedded(cpu) is None                                                 # direct path requested
    few = MetricTensor(10, device=cpu)
    with contextlib.redirect_stdout(io.StringIO()):
        few.load_pretrained(torch.zeros(100, 10), torch.eye(10).repeat(100, 1, 1), temperature=1.0, regularization=0.01)
    assert few._embedded(cpu) is None                                                # small table
    assert MetricTensor.EMBED_MIN_CENTROIDS == 512
