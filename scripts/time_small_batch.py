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

# the reference-facing entry points at one small batch (wall clock per call, Python overhead included)
from rlvae_b200 import WorkingRiemannianSampler, _capi
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=0.7,
                       regularization=sm.regularization)
model = MetricModel(mt)
ws, hs = WorkingRiemannianSampler(model), RiemannianHMCSampler(model)
mu = make_points(1024, 16, seed=3).to(dev)
lv = torch.full_like(mu, -2.0)


def wall(name, fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f'  {name:46s} {(time.perf_counter() - t0) / reps * 1e6:8.1f} us')


print('entry points at n = 1024, K = 200, T = 0.7:')
with torch.no_grad():
    wall('nearest2', lambda: _capi.nearest2(mt._tables(dev), mu))
    for m in ('enhanced', 'geodesic', 'basic'):
        wall(f'WorkingRiemannianSampler {m}', lambda m=m: ws.sample_riemannian_latents(mu, lv, method=m))
    wall("RiemannianHMCSampler 'hmc' refine", lambda: hs.sample_riemannian_latents(mu, lv, method='hmc'))
    wall('compute_inverse_metric', lambda: mt.compute_inverse_metric(mu))
    wall('compute_metric', lambda: mt.compute_metric(mu))
    wall('compute_log_det_metric', lambda: mt.compute_log_det_metric(mu))
    wall('compute_metric_spectrum', lambda: mt.compute_metric_spectrum(mu))
    wall('riemannian_distance_squared', lambda: mt.compute_riemannian_distance_squared(mu, mu.roll(1, 0)))
    wall('log_pi', lambda: hs.log_pi(mu))
    wall('grad_func', lambda: hs.grad_func(mu))
mug = mu.clone().requires_grad_(True)


def bwd():
    mug.grad = None
    mt.compute_log_det_metric(mug).sum().backward()


wall('log det + autograd backward', bwd)

# RiemannianHMCSampler.sample at the reference's default schedule (100 MCMC steps x 15 leapfrog).
# (Replaying each MCMC iteration as one CUDA graph was tried and changed nothing -- 49.3 vs 48.7 ms: at
# these sizes the loop is bound by the ~30 us of GPU-side kernel latency per metric evaluation, not by
# the host's launches.)
hs_full = RiemannianHMCSampler(model, mcmc_steps_nbr=100, n_lf=15, eps_lf=0.03)
for nn in (64, 1024):
    hs_full.sample(nn)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hs_full.sample(nn)
    torch.cuda.synchronize()
    print(f'  sample({nn}), 100 x 15 leapfrog: {(time.perf_counter() - t0) * 1e3:7.1f} ms')
