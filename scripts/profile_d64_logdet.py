"""One d = 64 forward-only chunk (G^-1 + log det G: tensor kernel, unpack, spd64_logdet_kernel) for ncu.
usage: python scripts/profile_d64_logdet.py [n_points] [K]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
d = 64
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1234)
c = torch.randn(K, d, device=dev, generator=g)
L = torch.tril(torch.randn(K, d, d, device=dev, generator=g)) * d ** -0.5
M = L @ L.transpose(1, 2)
M = 0.5 * (M + M.transpose(1, 2))
T, lam = 0.75 * d ** 0.5, 0.01
z = torch.randn(n, d, device=dev, generator=g)
w = torch.exp(-torch.cdist(z[:64].double(), c.double()) ** 2 / T ** 2).sum(1).mean().item()
mt = MetricTensor(d, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(c, (M / w).contiguous(), temperature=T, regularization=lam)
out = {}
for _ in range(2):
    out = mt.evaluate(z, want_ginv=True, want_logdet=True, out=out)
torch.cuda.synchronize()
print('ok', float(out['logdet_g'][:4].sum()))
