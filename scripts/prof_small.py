"""Per-tile cycle breakdown of the fused forward kernel (prologue / main loop / epilogue) at K = 200 and
K = 10,000; needs a profile build: RLVAE_NVCC_EXTRA=-DRLVAE_TC_PROFILE python -m rlvae_b200.build"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric
dev = torch.device('cuda:0')
for K in (200, 10000):
    sm = make_synthetic_metric(K, 16, seed=0)
    mt = MetricTensor(16, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=3.0, regularization=sm.regularization)
    for n in (64, 148 * 128 * 4):
        z = make_points(n, 16, seed=1).to(dev)
        print(f'--- K={K} n={n}', flush=True)
        for _ in range(2):
            mt.evaluate(z, want_ginv=False, want_logdet=True)
            torch.cuda.synchronize()
