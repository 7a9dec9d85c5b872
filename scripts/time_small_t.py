"""Time the d = 16 symmetric tensor kernels at a small temperature (default T = 0.7, the reference's
conf/model/hybrid_rlvae.yaml value) in the weight mode the library picks, and compare with the direct
kernels.  Usage: python scripts/time_small_t.py [T] [N] [near_fraction]; RLVAE_TC_EXACT=1|2 forces a mode."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric

T = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
near = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)


def make(path):
    mt = MetricTensor(16, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=T,
                           regularization=sm.regularization)
    return mt


mt = make('auto')
print('T', T, mt.kernel_info()['implementation'])
z = make_points(N, 16, seed=1)
nn = int(near * N)
if nn:   # points next to centroids, like encoder outputs
    idx = torch.randint(0, 10000, (nn,), generator=torch.Generator().manual_seed(3))
    z[:nn] = sm.centroids[idx] + 0.1 * z[:nn]
z = z.to(dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


fwd = timed(lambda: mt.evaluate(z, want_ginv=False, want_logdet=True))
full = timed(lambda: mt.evaluate(z, want_ginv=False, want_logdet=True, want_grad=True))
print(f'forward {fwd:.2f} ms   forward+gradient {full:.2f} ms   ({N / full / 1e3:.1f} M evals/s)')
step = timed(lambda: mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True))
print(f'the bench step (expanded G^-1 written as well): {step:.2f} ms')
m = 1 << 14
sel = torch.cat([torch.arange(m // 2), torch.arange(N - m // 2, N)]).to(dev)
zs = z[sel].contiguous()
a = make('direct').evaluate(zs, want_ginv=True, want_logdet=True, want_grad=True)
b = mt.evaluate(zs, want_ginv=True, want_logdet=True, want_grad=True)
rel = lambda x, y: ((x - y).flatten(1).norm(dim=1) / y.flatten(1).norm(dim=1).clamp_min(1e-30)).max().item()
g = a['grad_logdet_g']
live = g.norm(dim=1) > 1e-6 * g.norm(dim=1).max()
print('vs direct: ginv', rel(b['ginv'], a['ginv']), ' logdet', (b['logdet_g'] - a['logdet_g']).abs().max().item(),
      ' grad', rel(b['grad_logdet_g'][live], g[live]))
