"""A few fused evaluations and one short HMC iteration at n = 64 and 4096 (K = 200), for an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv`): per-kernel latency at small sizes."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler
from rlvae_b200.synthetic import make_hmc_streams, make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(200, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=3.0, regularization=sm.regularization)
for n in (64, 4096):
    z = make_points(n, 16, seed=1).to(dev)
    for _ in range(3):
        mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True)
    z0, gam, acc = make_hmc_streams(n, 16, 1, seed=2)
    s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=1, n_lf=3, eps_lf=0.03)
    s.sample_with_streams(z0.to(dev), gam.to(dev), acc.to(dev))
torch.cuda.synchronize()
print('ok')
