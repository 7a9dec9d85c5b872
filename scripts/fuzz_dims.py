"""Randomised check of every latent_dim 1..64 (power of two or not) against an fp64 evaluation of the reference
expressions on the GPU: G^-1, G, log det G, grad_z log det G, the pythae-variant gradient, nearest2, the spectrum
(d = 16).  AUTO path (tensor kernels where the tables allow them) and the CUDA-core path.
usage: python scripts/fuzz_dims.py [n_cases] [seed0]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi

dev = torch.device('cuda:0')
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def rel_rows(a, b, live=None):
    a, b = a.flatten(1).double(), b.flatten(1).double()
    e = (a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)
    if live is not None:
        e = e[live]
    return e.max().item() if e.numel() else 0.0


def mk(c, M, T, lam, d, path):
    mt = MetricTensor(d, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(c.clone(), M.clone(), temperature=T, regularization=lam)
    return mt


def ref64(c, M, z, T, lam, d):
    c, M = c.to(dev).double(), M.to(dev).double()
    zz = z.double().clone().requires_grad_(True)
    outs, pys = [], []
    delta = c[None] - zz[:, None]
    w = torch.exp(-(delta ** 2).sum(-1) / T ** 2)
    ginv = torch.einsum('nk,kij->nij', w, M) + lam * torch.eye(d, device=dev, dtype=torch.float64)
    g = torch.linalg.inv(ginv)
    ld = -torch.linalg.slogdet(ginv).logabsdet
    grad = torch.autograd.grad(ld.sum(), zz)[0]
    v = torch.einsum('nk,kji,nkj->ni', w, M, delta)               # sum_k w_k M_k^T (c_k - z)
    py = torch.einsum('nij,ni->nj', g, v) / T ** 2                # G^T v
    return ginv.detach(), g.detach(), ld.detach(), grad.detach(), py.detach()


worst, bad = {}, 0
dims = [1, 2, 3, 5, 8, 10, 12, 16, 16, 16, 20, 24, 32, 33, 48, 64]
for case in range(cases):
    gen = torch.Generator().manual_seed(seed0 + case)
    r = lambda: torch.rand(1, generator=gen).item()
    d = dims[case % len(dims)]
    K = int(1 + r() ** 2 * (300 if d <= 32 else 80))
    n = int(1 + r() ** 2 * (300 if d <= 32 else 100))
    T = 10 ** (-0.7 + 1.4 * r()) * d ** 0.5 / 4            # scaled with sqrt(d) so that weights stay alive
    lam = [1e-3, 1e-2, 1e-1][int(r() * 3) % 3]
    sym = r() < 0.8
    c = torch.randn(K, d, generator=gen) + (r() < 0.3) * 3.0 * torch.randn(d, generator=gen)
    L = torch.tril(torch.randn(K, d, d, generator=gen)) * d ** -0.5
    M = (L @ L.transpose(1, 2)) * 10 ** (-1 + 2 * r())
    if not sym:
        M = M + 0.05 * torch.randn(K, d, d, generator=gen) * M.abs().mean()
    near = c[torch.randint(K, (n,), generator=gen)] + 0.3 * T * torch.randn(n, d, generator=gen) / d ** 0.5
    far = torch.randn(n, d, generator=gen)
    z = torch.where(torch.rand(n, 1, generator=gen) < 0.6, near, far).contiguous().to(dev)
    ginv, g, ld, grad, py = ref64(c, M, z, T, lam, d)
    cond = torch.linalg.cond(ginv).max().item()
    for path in ('auto', 'direct'):
        mt = mk(c, M, T, lam, d, path)
        tab = mt._tables(dev)
        tag = f'case {seed0 + case} d={d} K={K} n={n} T={T:.3f} lam={lam} sym={sym} cond={cond:.1e} path={path} impl={mt.kernel_info().get("kind", "?")}'
        ev = mt.evaluate(z, want_ginv=True, want_g=True, want_logdet=True, want_grad=True)
        gl = grad.norm(dim=1) > 1e-6 * grad.norm(dim=1).max()
        pl = py.norm(dim=1) > 1e-3 * py.norm(dim=1).max()
        pg, lad, sgn = _capi.pythae_eval(tab, z, path=mt._path())
        # tolerances scale with the conditioning of G^-1 wherever G is applied (fp32 inputs: eps * cond)
        amp = max(1.0, cond * 1e-2)
        errs = {'ginv': rel_rows(ev['ginv'], ginv) / 1e-5, 'g': rel_rows(ev['g'], g) / (1e-5 * amp),
                'logdet': ((ev['logdet_g'].double() - ld).abs() / (1 + ld.abs())).max().item() / 1e-4,
                'grad': rel_rows(ev['grad_logdet_g'], grad, gl) / (1e-4 * amp),
                'pythae': rel_rows(pg, py, pl) / (1e-4 * amp),
                'pythae_lad': ((lad.double() + ld).abs() / (1 + ld.abs())).max().item() / 1e-4}
        if K >= 2:
            idx, dist = _capi.nearest2(tab, z)
            dd = torch.cdist(z.double(), c.to(dev).double())
            top = dd.topk(2, dim=1, largest=False)
            # ties in fp32 may legitimately order differently from fp64: compare distances, not indices
            errs['nearest2'] = ((dist.double() - top.values).abs() / (1e-30 + top.values)).max().item() / 1e-5
        over = {k: v for k, v in errs.items() if not (v <= 1.0)}
        for k, v in errs.items():
            if not (v <= worst.get(k, (0.0, ''))[0]):
                worst[k] = (v, tag)
        if over:
            bad += 1
            print(tag, 'OVER (x tolerance)', {k: f'{v:.2f}' for k, v in over.items()}, flush=True)
print(f'{cases} cases x 2 paths, {bad} over the limits')
for k, (v, tag) in worst.items():
    print(f'worst {k:12s} {v:.3f} x tol  {tag}')
