"""BASELINE.json configs[4] ("large-metric stress"): d = 64, K = 50,000 centroids, tables streamed via TMA.
One GPU's share of the work at reduced N (default 2^17 points; the config's 1M points per GPU is 8x this).
Checks the split-fp16 tensor kernel against the direct kernel and an fp64 evaluation, and times it.
usage: python scripts/stress_d64.py [n_points] [K]"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
d = 64
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
c = torch.randn(K, d, device=dev, generator=g)
L = torch.tril(torch.randn(K, d, d, device=dev, generator=g)) * d ** -0.5
T, lam = 0.75 * d ** 0.5, 0.01
M = L @ L.transpose(1, 2)
del L
z = torch.randn(n, d, device=dev, generator=g)
# scale M so that G^-1 is O(1): sum_k w_k ~ K * mean weight at a few probe points (fp64)
w = torch.exp(-torch.cdist(z[:64].double(), c.double()) ** 2 / T ** 2).sum(1).mean().item()
M = M / w
rel = lambda x, y: ((x - y).flatten(1).norm(dim=1) / y.flatten(1).norm(dim=1)).max().item()

def mk(path):
    mt = MetricTensor(d, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(c, M, temperature=T, regularization=lam)
    return mt

mt = mk('tensor')
ev = mt.evaluate(z[:2048].contiguous(), want_ginv=True, want_logdet=True)
ref = mk('direct').evaluate(z[:2048].contiguous(), want_ginv=True, want_logdet=True)
print('tensor vs direct: ginv rel', rel(ev['ginv'], ref['ginv']), 'logdet abs',
      (ev['logdet_g'] - ref['logdet_g']).abs().max().item())
z4 = z[:4].double()
wk = torch.exp(-((z4[:, None, :] - c.double()[None]) ** 2).sum(-1) / T ** 2)
g64 = torch.einsum('nk,kij->nij', wk, M.double()) + lam * torch.eye(d, device=dev, dtype=torch.float64)
print('tensor vs fp64: ginv rel', rel(ev['ginv'][:4].double(), g64), 'logdet abs',
      (ev['logdet_g'][:4].double() + torch.linalg.slogdet(g64).logabsdet).abs().max().item())
out = {}
for _ in range(2):
    out = mt.evaluate(z, want_ginv=True, want_logdet=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
out = mt.evaluate(z, want_ginv=True, want_logdet=True, out=out)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
flops = n * 2.0 * K * d * (d + 1)
print(f'd=64 K={K} N={n}: G^-1 + log det {ms:.1f} ms = {n / ms * 1e3:.3e} evals/s, '
      f'{flops / ms / 1e9:.1f} TFLOP/s fp32-equivalent (dense-M definition)')
nd = min(n, 4096)
md = mk('direct')
md.evaluate(z[:nd].contiguous(), want_ginv=True, want_logdet=False)
torch.cuda.synchronize()
e0.record()
md.evaluate(z[:nd].contiguous(), want_ginv=True, want_logdet=False)
e1.record(); e1.synchronize()
print(f'direct kernel on {nd} points: {e0.elapsed_time(e1):.1f} ms = {nd / e0.elapsed_time(e1) * 1e3:.3e} evals/s')
