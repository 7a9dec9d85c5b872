"""Step time (forward + gradient, N = 2^20, K = 10k) for the CTA-pair / single-CTA forms of the two tensor kernels,
sustained (20 steps back to back) and bench-style (L2 flush between steps).  Run once per setting:
RLVAE_TC_PAIR=<0|1> RLVAE_TC_PAIR_GRAD=<0|1> python scripts/time_pair_modes.py"""
import contextlib, ctypes, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric

dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
z = make_points(1 << 20, 16, seed=1).to(dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
out = {}
for _ in range(3):
    out = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True, out=out)
lib = _capi.lib()
buf = (ctypes.c_float * 3)()
for name, do_flush, reps in (('sustained', False, 20), ('flushed', True, 8)):
    lib.rlvae_profile_begin(reps)
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        out = mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=True, out=out)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    f = g = 0.0
    for i in range(reps):
        lib.rlvae_profile_read(i, buf); f += buf[0]; g += buf[2]
    step = sum(a.elapsed_time(b) for a, b in evs) / reps
    print(f"PAIR={os.environ.get('RLVAE_TC_PAIR', '1')} PAIR_GRAD={os.environ.get('RLVAE_TC_PAIR_GRAD', '-')} {name:9s}: "
          f"step {step:6.2f} ms  forward {f / reps:5.2f}  gradient {g / reps:5.2f}")
    lib.rlvae_profile_end()
