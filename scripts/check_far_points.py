"""Accuracy of the tensor path against the direct kernels for points progressively farther from the
centroid cloud (|z| scaled 1-5x) at three temperatures: expanded form, gated expanded form, hybrid mode."""
import contextlib, io, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
def make(path, T):
    mt = MetricTensor(16, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=T, regularization=sm.regularization)
    return mt
rel = lambda x, y: ((x - y).flatten(1).norm(dim=1) / y.flatten(1).norm(dim=1).clamp_min(1e-30)).max().item()
print('T default', sm.temperature, 'lambda', sm.regularization)
for T in (sm.temperature, 1.5, 0.7):
    a, b = make('direct', T), make('auto', T)
    print('T', T, b.kernel_info()['implementation'][:60])
    for scale in (1.0, 1.5, 2.0, 3.0, 5.0):
        z = (scale * make_points(8192, 16, seed=5)).to(dev)
        ea = a.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True)
        eb = b.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True)
        g = ea['grad_logdet_g']; live = g.norm(dim=1) > 1e-6 * g.norm(dim=1).max()
        print(f'  |z| scale {scale}: ginv {rel(eb["ginv"], ea["ginv"]):.2e}  logdet {(eb["logdet_g"]-ea["logdet_g"]).abs().max().item():.2e}  grad {rel(eb["grad_logdet_g"][live], g[live]):.2e} (fro {((eb["grad_logdet_g"]-g).norm()/g.norm()).item():.2e})')
