"""Randomised cross-check of the d = 16 tensor path against the CUDA-core path (which the parity tests pin
against the oracle): random K, batch size, temperature, lambda, table offset and scale; every output of the
fused evaluation, the pythae variant, and the fused HMC trajectory against the per-step path.
usage: python scripts/fuzz_d16.py [n_cases] [seed0]"""
import contextlib, io, math, os, sys
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


def mk(c, M, T, lam, path):
    mt = MetricTensor(16, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(c.clone(), M.clone(), temperature=T, regularization=lam)
    return mt


worst = {}
bad = 0
for case in range(cases):
    g = torch.Generator().manual_seed(seed0 + case)
    r = lambda: torch.rand(1, generator=g).item()
    K = int(1 + r() ** 2 * 900)
    n = int(1 + r() ** 2 * 700) if case % 3 else 2049 + int(r() * 3000)     # every third case: the long-batch (split + bound) pythae path
    T = 10 ** (-1 + 1.8 * r())                      # 0.1 .. 6.3
    lam = [1e-3, 1e-2, 1e-1][int(r() * 3) % 3]
    off = (r() < 0.3) * 5.0 * torch.randn(16, generator=g)
    c = torch.randn(K, 16, generator=g) + off
    L = torch.tril(torch.randn(K, 16, 16, generator=g)) * 0.25
    M = (L @ L.transpose(1, 2)) * 10 ** (-2 + 3 * r())
    # half of the points near centroids (small T would otherwise leave only lambda I)
    near = c[torch.randint(K, (n,), generator=g)] + 0.3 * T * torch.randn(n, 16, generator=g) / 4
    far = torch.randn(n, 16, generator=g) + off
    z = torch.where(torch.rand(n, 1, generator=g) < 0.5, near, far).contiguous().to(dev)
    mt, md = mk(c, M, T, lam, 'auto'), mk(c, M, T, lam, 'direct')
    tab = mt._tables(dev)
    tag = f'case {seed0 + case}: K={K} n={n} T={T:.3f} lam={lam} mode={tab.weight_mode} auto={tab.tensor_auto}'
    if not (tab.tensor_capable and tab.tensor_auto):
        print(tag, 'SKIP (no tensor path)')
        continue
    a = mt.evaluate(z, want_ginv=True, want_g=True, want_logdet=True, want_grad=True)
    b = md.evaluate(z, want_ginv=True, want_g=True, want_logdet=True, want_grad=True)
    errs = {'ginv': rel_rows(a['ginv'], b['ginv']), 'g': rel_rows(a['g'], b['g']),
            'logdet': ((a['logdet_g'] - b['logdet_g']).abs() / (1 + b['logdet_g'].abs())).max().item()}
    if os.environ.get('FUZZ_FP64'):
        def ginv_ref(dtype):
            cz, Mz, zz = c.to(dev, dtype), M.to(dev, dtype), z.to(dtype)
            outs = []
            for lo in range(0, n, 64):
                delta = cz[None] - zz[lo:lo + 64][:, None]
                w = torch.exp(-(torch.norm(delta, dim=-1) ** 2) / (T ** 2))
                outs.append((Mz[None] * w[:, :, None, None]).sum(1) + lam * torch.eye(16, device=dev, dtype=dtype))
            return torch.cat(outs)
        gv64, gv32 = ginv_ref(torch.float64), ginv_ref(torch.float32)
        print(tag, 'G^-1 vs fp64: tensor %.2e  direct %.2e  reference-fp32 %.2e' % (
            rel_rows(a['ginv'], gv64), rel_rows(b['ginv'], gv64), rel_rows(gv32, gv64)), flush=True)
        gi64 = torch.linalg.inv(b['ginv'].double())              # truth for the inverse of the direct path's own G^-1
        g32 = torch.linalg.inv(b['ginv'])
        print(tag, 'G vs fp64 inverse of the same G^-1: direct (Gauss-Jordan) %.2e  torch fp32 inv %.2e | tensor (Cholesky of its own G^-1) %.2e'
              % (rel_rows(b['g'], gi64), rel_rows(g32, gi64), rel_rows(a['g'], torch.linalg.inv(a['ginv'].double()))), flush=True)
    gn = b['grad_logdet_g'].norm(dim=1)
    errs['grad'] = rel_rows(a['grad_logdet_g'], b['grad_logdet_g'], gn > 1e-6 * gn.max())
    pa, la, sa = _capi.pythae_eval(tab, z, path=_capi.PATH_TENSOR)
    pb, lb, sb = _capi.pythae_eval(md._tables(dev), z, path=_capi.PATH_DIRECT)
    pn = pb.norm(dim=1)
    errs['pythae'] = rel_rows(pa, pb, pn > 1e-3 * pn.max())
    errs['pythae_abs'] = ((pa - pb).norm(dim=1).max() / pn.max().clamp_min(1e-30)).item()
    errs['pythae_lad'] = ((la - lb).abs() / (1 + lb.abs())).max().item()
    # the same expression in fp64 (truth) and in the reference's fp32 eager arithmetic: our error has to stay in the
    # class of the reference's own rounding noise
    def pythae_ref(dtype):
        cz, Mz, zz = c.to(dev, dtype), M.to(dev, dtype), z.to(dtype)
        outs = []
        for lo in range(0, n, 64):
            zc = zz[lo:lo + 64]
            delta = cz[None] - zc[:, None]
            w = torch.exp(-(torch.norm(delta, dim=-1) ** 2) / (T ** 2))
            ginv = (Mz[None] * w[:, :, None, None]).sum(1) + lam * torch.eye(16, device=dev, dtype=dtype)
            gm = torch.linalg.inv(ginv)
            v = (delta[:, :, None, :] @ (Mz[None] * w[:, :, None, None])).sum(1)
            outs.append((gm.transpose(-1, -2) @ v.transpose(-1, -2) / (T ** 2)).squeeze(-1))
        return torch.cat(outs)
    t64, t32 = pythae_ref(torch.float64), pythae_ref(torch.float32)
    n64 = t64.norm(dim=1)
    lv = n64 > 1e-3 * n64.max()
    e_ref = rel_rows(t32, t64, lv)
    # exact gradient (variant D) in fp64: grad_z log det G = -(2/T^2) sum_k w_k tr(G M_k) (c_k - z)
    def grad_ref64():
        cz, Mz, zz = c.to(dev).double(), M.to(dev).double(), z.double()
        outs = []
        for lo in range(0, n, 64):
            delta = cz[None] - zz[lo:lo + 64][:, None]
            w = torch.exp(-(delta ** 2).sum(-1) / T ** 2)
            gm = torch.linalg.inv(torch.einsum('nk,kij->nij', w, Mz) + lam * torch.eye(16, device=dev, dtype=torch.float64))
            tr = torch.einsum('nij,kji->nk', gm, Mz)
            outs.append(-(2.0 / T ** 2) * torch.einsum('nk,nkj->nj', w * tr, delta))
        return torch.cat(outs)
    g64 = grad_ref64()
    gl64 = g64.norm(dim=1) > 1e-6 * g64.norm(dim=1).max()
    errs['grad64_tensor'] = rel_rows(a['grad_logdet_g'], g64, gl64)
    errs['grad64_direct'] = rel_rows(b['grad_logdet_g'], g64, gl64)
    if os.environ.get('FUZZ_FP64'):
        print(tag, 'grad vs fp64: tensor %.2e  direct %.2e' % (errs['grad64_tensor'], errs['grad64_direct']), flush=True)
    errs['pythae64_tensor'] = rel_rows(pa, t64, lv) / max(10 * e_ref, 5e-5)
    errs['pythae64_direct'] = rel_rows(pb, t64, lv) / max(10 * e_ref, 5e-5)
    if os.environ.get('FUZZ_FP64'):
        print(tag, 'pythae vs fp64: tensor %.2e  direct %.2e  reference-fp32 %.2e' % (
            rel_rows(pa, t64, lv), rel_rows(pb, t64, lv), e_ref), flush=True)
    # fused HMC trajectory vs per-step launches
    if _capi.hmc_fused_available(tab):
        iters, n_lf = 2, 5
        gam = torch.randn(iters, n, 16, generator=g).to(dev)
        acc = torch.rand(iters, n, generator=g).to(dev)
        scales = [1.0] * (iters * n_lf)
        z1, z2 = z.clone(), z.clone()
        r1 = _capi.hmc_run(tab, z1, gam, acc, n_lf, 0.03, 1.0, scales, want_stats=True)
        r2 = _capi.hmc_run(tab, z2, gam, acc, n_lf, 0.03, 1.0, scales, grad_mode=_capi.GRAD_MODULAR | _capi.HMC_NO_FUSION,
                           want_stats=True)
        al1, al2 = r1['stats'][2], r2['stats'][2]
        flips = (r1['stats'][3] != r2['stats'][3])
        tie = ((acc - al2).abs() < 1e-4 * (1 + al2.abs()))
        errs['hmc_flips_not_tie'] = float((flips & ~tie).sum().item())
        same = ~flips.any(dim=0)
        errs['hmc_z'] = ((z1 - z2)[same].abs().max().item() if same.any() else 0.0)
        fin = torch.isfinite(al2)
        errs['hmc_alpha'] = ((al1 - al2)[fin].abs().max().item() if fin.any() else 0.0)
    lim = {'ginv': 1e-5, 'g': 1e-4, 'logdet': 2e-4, 'grad': 2e-4, 'pythae': 1e-2, 'pythae_abs': 1e-2, 'pythae_lad': 2e-4,
           'pythae64_tensor': 1.0, 'pythae64_direct': 3.0, 'grad64_tensor': 2e-4, 'grad64_direct': 2e-4,
           'hmc_flips_not_tie': 0.5, 'hmc_z': 1e-3, 'hmc_alpha': 1e-3}
    over = {k: v for k, v in errs.items() if not (v <= lim[k])}
    for k, v in errs.items():
        if not (v <= worst.get(k, (0.0, ''))[0]):
            worst[k] = (v, tag)
    if over:
        bad += 1
        print(tag, 'OVER', {k: f'{v:.2e}' for k, v in over.items()}, flush=True)
print(f'{cases} cases, {bad} over the limits')
for k, (v, tag) in worst.items():
    print(f'worst {k:18s} {v:.3e}  {tag}')
