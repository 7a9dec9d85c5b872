"""torchrun --nproc-per-node W scripts/check_distributed_nccl.py
NCCL check of the batch-sharding plumbing with the real CUDA kernels: table broadcast, sharded
evaluation + all_gather equals the single-GPU evaluation, HMC is independent of the world size."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler
from rlvae_b200 import distributed as D
from rlvae_b200.synthetic import make_hmc_streams, make_points, make_synthetic_metric

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
sm = make_synthetic_metric(1000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    if rank == 0:
        mt.load_pretrained(**sm.as_load_kwargs())
    else:   # other ranks start from garbage tables of the right shape
        mt.load_pretrained(torch.zeros_like(sm.centroids), torch.eye(16).repeat(1000, 1, 1), temperature=1.0,
                           regularization=0.5)
D.broadcast_tables(mt, src=0)
assert torch.equal(mt.centroids.cpu(), sm.centroids) and abs(float(mt.temperature) - sm.temperature) < 1e-6
n = 10007
z = make_points(n, 16, seed=1).to(dev)
full = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True)
got = D.sharded_apply(lambda zz: {k: v for k, v in mt.evaluate(zz, want_ginv=True, want_logdet=True, want_grad=True).items()
                                  if k in ('ginv', 'logdet_g', 'grad_logdet_g')}, z)
for k in ('ginv', 'logdet_g', 'grad_logdet_g'):
    assert got[k].shape == full[k].shape, k
    torch.testing.assert_close(got[k], full[k], rtol=1e-6, atol=1e-7)
z0, gam, acc = make_hmc_streams(4099, 16, 2, seed=2)
z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=2, n_lf=3, eps_lf=0.03)
ref = s.sample_with_streams(z0, gam, acc)
zl, gl, al = D.shard_hmc_streams(z0, gam, acc)
loc = s.sample_with_streams(zl, gl, al)
allz = D.all_gather_rows(loc, 4099)
torch.testing.assert_close(allz, ref, rtol=1e-6, atol=1e-7)
moved = D.all_reduce_scalar((loc != zl).any(dim=1).sum().float())
assert int(moved.item()) == int((ref != z0).any(dim=1).sum().item())
dist.barrier()
if rank == 0:
    print(f'NCCL sharding check ok on {world} GPUs: evaluate + all_gather == single GPU, HMC independent of world size')
dist.destroy_process_group()
