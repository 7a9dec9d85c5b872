"""Where the time of one HMC iteration (20 leapfrog steps, 2^20 chains) goes."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler, _capi
from rlvae_b200.synthetic import make_hmc_streams, make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
tab = mt._tables(dev)
n = 1 << 20
z = make_points(n, 16, seed=1).to(dev)
def timeit(name, fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    print(f'{name:60s} {e0.elapsed_time(e1)/reps:8.2f} ms')
out = {}
def evals(k):
    global out
    for _ in range(k):
        out = _capi.metric_eval(tab, z, want_ginv=False, want_g=False, want_logdet=True, want_grad=False, out=out)
timeit('21 x metric_eval(log det only), back to back', lambda: evals(21))
z0, gam, acc = make_hmc_streams(n, 16, 1, seed=2)
z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=1, n_lf=20, eps_lf=0.03)
timeit('RiemannianHMCSampler: 1 MCMC iteration, 20 leapfrog', lambda: s.sample_with_streams(z0, gam, acc))
work = _capi.hmc_workspace(n, 16, dev)
import ctypes
zz = z0.clone()
timeit('rlvae_hmc_iteration alone (C ABI)', lambda: _capi.hmc_iteration(tab, zz, gam[0], acc[0], 20, 0.03, 1.0, [1.0] * 20, _capi.GRAD_MODULAR, work, _capi.PATH_AUTO))
