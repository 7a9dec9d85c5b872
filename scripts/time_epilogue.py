"""Cost of the fused per-point epilogue of the forward kernel: packed G^-1 only / + Cholesky log det / + packed G / + expanded G^-1."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(**sm.as_load_kwargs())
tab = mt._tables(dev)
n = 1 << 20
z = make_points(n, 16, seed=1).to(dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def t(name, fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f'{name:50s} {sum(ts)/len(ts):7.3f} ms')
pk = torch.empty(n, 144, device=dev)
out = {}
t('packed G^-1 only', lambda: _capi.inverse_metric_packed(tab, z, pk))
def ev(**kw):
    global out
    out = _capi.metric_eval(tab, z, out=out, **kw)
t('+ Cholesky log det', lambda: ev(want_ginv=False, want_g=False, want_logdet=True, want_grad=False))
t('+ expanded G^-1 [N,16,16]', lambda: ev(want_ginv=True, want_g=False, want_logdet=True, want_grad=False))
t('+ packed G and gradient kernel', lambda: ev(want_ginv=True, want_g=False, want_logdet=True, want_grad=True))
