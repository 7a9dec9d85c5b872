"""Run one forward and one gradient launch (N = 64k points) with the in-kernel cycle profile
(build with RLVAE_NVCC_EXTRA=-DRLVAE_TC_PROFILE)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev, kernel_path='tensor')
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
z = make_points(1 << 18, 16, seed=1).to(dev)
for _ in range(2):
    ev = mt.evaluate(z, want_ginv=False, want_logdet=True, want_grad=True)
    torch.cuda.synchronize()
print('ok', ev['logdet_g'][0].item())
