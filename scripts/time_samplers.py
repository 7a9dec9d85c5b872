"""Throughput of the WorkingRiemannianSampler / HMC public entry points at N = 2^20, K = 10k, d = 16."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler, WorkingRiemannianSampler, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
model = MetricModel(mt)
ws, hs = WorkingRiemannianSampler(model), RiemannianHMCSampler(model)
mu = make_points(n, 16, seed=3).to(dev)
lv = torch.full_like(mu, -2.0)

def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{name:42s} {ms:9.2f} ms  {n / ms * 1e3:.3e} points/s')

with torch.no_grad():
    timeit('nearest2 (top-2 centroids)', lambda: _capi.nearest2(mt._tables(dev), mu))
    for m in ('enhanced', 'geodesic', 'basic'):
        timeit(f'WorkingRiemannianSampler {m}', lambda m=m: ws.sample_riemannian_latents(mu, lv, method=m))
    timeit("RiemannianHMCSampler 'hmc' refine (3 steps)", lambda: hs.sample_riemannian_latents(mu, lv, method='hmc'))
    timeit('compute_inverse_metric', lambda: mt.compute_inverse_metric(mu))
    timeit('compute_metric', lambda: mt.compute_metric(mu))
    timeit('compute_log_det_metric', lambda: mt.compute_log_det_metric(mu))
    timeit('compute_metric_spectrum', lambda: mt.compute_metric_spectrum(mu))
    timeit('riemannian_distance_squared', lambda: mt.compute_riemannian_distance_squared(mu, mu.roll(1, 0)))
mug = mu[: 1 << 18].clone().requires_grad_(True)
def bwd():
    mug.grad = None
    mt.compute_log_det_metric(mug).sum().backward()
n = 1 << 18
timeit('log det + autograd backward (2^18 points)', bwd)
