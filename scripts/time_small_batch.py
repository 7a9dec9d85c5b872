"""Latency at the reference's own working sizes (its metric.pt has K = 200 centroids, d = 16; batches of
64-4096 latents): one fused evaluation and one HMC iteration.  Usage: python scripts/time_small_batch.py"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler
from rlvae_b200.synthetic import make_hmc_streams, make_points, make_synthetic_metric

dev = torch.device('cuda:0')
sm = make_synthetic_metric(200, 16, seed=0)
for T in (0.7, 3.0):
    mt = MetricTensor(16, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=T,
                           regularization=sm.regularization)
    print('T', T, mt.kernel_info()['implementation'])
    for n in (64, 1024, 4096, 65536):
        z = make_points(n, 16, seed=1).to(dev)
        out = {}
        for _ in range(5):
            out = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            out = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True, out=out)
        torch.cuda.synchronize()
        ev_us = (time.perf_counter() - t0) / reps * 1e6
        z0, gam, acc = make_hmc_streams(n, 16, 3, seed=2)
        s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=3, n_lf=10, eps_lf=0.03)
        z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
        s.sample_with_streams(z0, gam, acc)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            s.sample_with_streams(z0, gam, acc)
        torch.cuda.synchronize()
        hmc_us = (time.perf_counter() - t0) / 5 / 3 * 1e6
        print(f'  n={n:6d}: evaluate (G^-1 + log det + grad) {ev_us:8.1f} us   HMC iteration (10 leapfrog) {hmc_us:8.1f} us')
