"""End-to-end (pinned host buffers) evaluation rate of HostEvaluator for several chunk sizes."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.host_pipeline import HostEvaluator
from rlvae_b200.synthetic import make_points, make_synthetic_metric

dev = torch.device('cuda:0')
n = 1 << 20
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
z = make_points(n, 16, seed=1).pin_memory()
ld = torch.empty(n).pin_memory()
gr = torch.empty(n, 16).pin_memory()
ginv = torch.empty(n, 16, 16, device=dev)
for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 148 * 128 * 6, 148 * 128 * 7, 148 * 128 * 14):
    he = HostEvaluator(mt, chunk=chunk, want_grad=True)
    for _ in range(3):
        he(z, ld, gr, ginv_dev=ginv)
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); he(z, ld, gr, ginv_dev=ginv); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f'chunk {chunk:8d}: {best:7.2f} ms  {n / best / 1e3:.1f} M evals/s')
