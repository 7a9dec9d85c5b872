"""RiemannianHMCSampler.sample at the reference's defaults (100 MCMC steps x 15 leapfrog, K = 200), kernel time only:
CUDA events around rlvae_hmc_run with the draws already in memory, repeated to average out the clock state.
usage: python scripts/time_sample_small.py [temperature]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_hmc_streams, make_synthetic_metric

T = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
dev = torch.device('cuda:0')
sm = make_synthetic_metric(200, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(sm.centroids, sm.metric_matrices, temperature=T, regularization=sm.regularization)
tab = mt._tables(dev)
iters, n_lf = 100, 15
for n in (64, 1024):
    z0, gam, acc = make_hmc_streams(n, 16, iters, seed=2)
    z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
    scales = [1.0] * (iters * n_lf)
    for mode, name in ((_capi.GRAD_MODULAR, 'fused'), (_capi.GRAD_MODULAR | _capi.HMC_NO_FUSION, 'per-step')):
        ts, work = [], None
        for rep in range(6):
            z = z0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            work = _capi.hmc_run(tab, z, gam, acc, n_lf, 0.03, 1.0, scales, mode, work=work)['work']
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[1:])
        print(f'T={T} n={n:5d} {name:9s}: median {ts[len(ts) // 2]:7.2f} ms  min {ts[0]:7.2f} ms')
