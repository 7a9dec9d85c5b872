"""One fused metric-evaluation step at BASELINE.json configs[1] size, for ncu (see profiles/README.md).
usage: python scripts/profile_step.py [n_points] [steps]"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
z = make_points(n, 16, seed=1).to(dev)
out = {}
for _ in range(steps):
    out = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True, out=out)
torch.cuda.synchronize()
print('ok', float(out['logdet_g'][:8].sum()))
if len(sys.argv) > 3 and sys.argv[3] == 'hmc':     # plus one short HMC iteration (2 leapfrog steps)
    from rlvae_b200 import MetricModel, RiemannianHMCSampler
    from rlvae_b200.synthetic import make_hmc_streams
    z0, gam, acc = make_hmc_streams(n, 16, 1, seed=2)
    s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=1, n_lf=2, eps_lf=0.03)
    zf = s.sample_with_streams(z0.to(dev), gam.to(dev), acc.to(dev))
    torch.cuda.synchronize()
    print('hmc ok', float(zf[:4].sum()))
    from rlvae_b200 import _capi
    idx, dist = _capi.nearest2(mt._tables(dev), z)             # A14/A15: tensor-core nearest2
    sp = mt.compute_metric_spectrum(z)                          # A19: forward + per-thread Jacobi
    torch.cuda.synchronize()
    print('nearest2 / spectrum ok', int(idx[0, 0]), float(sp['condition_number'][:4].mean()))
    pg, pl, ps = _capi.pythae_eval(mt._tables(dev), z)          # A8: forward + unit-weight gradient pass + finish
    torch.cuda.synchronize()
    print('pythae ok', float(pg[:4].sum()))
