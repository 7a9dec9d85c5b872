"""Fused HMC trajectory kernel vs the per-step launches: time per MCMC iteration at several batch sizes.
usage: python scripts/time_hmc_fused.py [K] [n_lf] [temperature]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_hmc_streams, make_synthetic_metric

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_lf = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = float(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device('cuda:0')
sm = make_synthetic_metric(K, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    kw = sm.as_load_kwargs()
    if T is not None:
        kw['temperature'] = T
    mt.load_pretrained(**kw)
tab = mt._tables(dev)
print(f'K={K} n_lf={n_lf} T={float(mt.temperature):.2f} weight_mode={tab.weight_mode} fused available: {_capi.hmc_fused_available(tab)}')
# bring the GPU out of its idle clocks first (a cold B200 sits at 120 MHz and takes a while to boost: short
# kernels timed cold come out ~25 % slower)
_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(200):
    _a @ _a
torch.cuda.synchronize()
del _a
for n in (64, 256, 18944, 37888, 1 << 17, 1 << 20):
    z0, gam, acc = make_hmc_streams(n, 16, 1, seed=2)
    z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
    scales = [1.0] * n_lf
    out = []
    for mode in (_capi.GRAD_MODULAR, _capi.GRAD_MODULAR | _capi.HMC_NO_FUSION):
        work = None
        best = 1e9
        for rep in range(6):
            z = z0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            r = _capi.hmc_run(tab, z, gam, acc, n_lf, 0.03, 1.0, scales, mode, work=work)
            e1.record(); e1.synchronize()
            work = r['work']
            if rep:
                best = min(best, e0.elapsed_time(e1))
        out.append(best)
    print(f'n={n:8d}: fused {out[0]:9.3f} ms ({1e3 * out[0] / (n_lf + 1):8.1f} us/eval)   per-step {out[1]:9.3f} ms '
          f'({1e3 * out[1] / (n_lf + 1):8.1f} us/eval)   chain-steps/s fused {n * n_lf / out[0] * 1e3:.3e}')
