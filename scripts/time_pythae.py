"""Variant C (pythae) on the tensor path vs the CUDA-core path: rlvae_pythae_eval at N = 2^20 / 2^16 (K = 10k, d = 16),
and OfficialRHVAESampler.sample_prior(32) (100 x 15 leapfrog steps at T = 0.1) wall time."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricModel, MetricTensor, OfficialRHVAESampler, RHVAEStyleHMCSampler, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric

dev = torch.device('cuda:0')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sm = make_synthetic_metric(K, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
tab = mt._tables(dev)

def timeit(name, fn, n, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{name:52s} {ms:9.3f} ms  {n / ms * 1e3:.3e} evals/s', flush=True)

with torch.no_grad():
    for n, paths in ((1 << 20, (_capi.PATH_TENSOR,)), (1 << 14, (_capi.PATH_TENSOR, _capi.PATH_DIRECT)),
                     (32, (_capi.PATH_TENSOR, _capi.PATH_DIRECT))):
        z = make_points(n, 16, seed=3).to(dev)
        for p in paths:
            nm = 'tensor' if p == _capi.PATH_TENSOR else 'direct'
            timeit(f'pythae_eval n={n} {nm}', lambda: _capi.pythae_eval(tab, z, path=p), n, reps=3 if n > 1000 else 50)
    # the hard corner: T = 0.1, every point a fraction of T away from a centroid -> the error bound flags every row
    # and the per-centroid kernel redoes it
    mt01 = MetricTensor(16, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        mt01.load_pretrained(sm.centroids, sm.metric_matrices, temperature=0.1, regularization=sm.regularization)
    n = 1 << 16
    zn = (sm.centroids[torch.randint(K, (n,))] + 0.0075 * make_points(n, 16, seed=5)).to(dev)
    timeit(f'pythae_eval n={n} T=0.1 next to centroids (all rows flagged)', lambda: _capi.pythae_eval(mt01._tables(dev), zn), n)
    model = MetricModel(mt)
    for path in ('auto', 'direct'):
        mt.kernel_path = path
        s = OfficialRHVAESampler(model)
        s.sample_prior(32); torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.sample_prior(32); torch.cuda.synchronize()
        print(f'OfficialRHVAESampler.sample_prior(32) [{path}]: {(time.perf_counter() - t0) * 1e3:.1f} ms wall', flush=True)
    mt.kernel_path = 'auto'
    s = RHVAEStyleHMCSampler(model, mcmc_steps_nbr=2, n_lf=15, eps_lf=0.03)
    n = 1 << 18
    s.hmc_sampling(n); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.hmc_sampling(n); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'RHVAEStyleHMCSampler 2 x 15 leapfrog, {n} chains: {dt * 1e3:.1f} ms = {n * 30 / dt:.3e} chain-steps/s', flush=True)
