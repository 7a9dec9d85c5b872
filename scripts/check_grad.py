"""Quick GPU check of the tensor-path gradient kernel against the direct kernel + timing.
usage: python scripts/check_grad.py [n_time]"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric

dev = torch.device('cuda:0')
K = int(os.environ.get('K', 10000))
sm = make_synthetic_metric(K, 16, seed=0)

def mk(path):
    mt = MetricTensor(16, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(**sm.as_load_kwargs())
    return mt

z = make_points(8192 + 77, 16, seed=1).to(dev)
a = mk('direct').evaluate(z, want_g=True, want_grad=True)
b = mk('tensor').evaluate(z, want_g=True, want_grad=True)
torch.cuda.synchronize()
rel = lambda x, y: ((x - y).flatten(1).norm(dim=1) / y.flatten(1).norm(dim=1)).max().item()
print('grad rel err tensor vs direct', rel(b['grad_logdet_g'], a['grad_logdet_g']))
print('ginv rel', rel(b['ginv'], a['ginv']))
print('g rel', rel(b['g'], a['g']), 'logdet abs', (b['logdet_g'] - a['logdet_g']).abs().max().item())
# arbitrary (non-symmetric) U through the autograd backward
U = torch.randn(z.shape[0], 16, 16, device=dev)
ga = _capi.metric_grad(mk('direct')._tables(dev), z, U, 2.0 / 9.0, _capi.PATH_DIRECT)
gb = _capi.metric_grad(mk('tensor')._tables(dev), z, U, 2.0 / 9.0, _capi.PATH_TENSOR)
print('backward rel err', rel(gb, ga))
n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
zt = make_points(n, 16, seed=2).to(dev)
mt = mk('tensor')
out = {}
for want_grad in (False, True):
    for _ in range(2):
        out = mt.evaluate(zt, want_ginv=True, want_logdet=True, want_grad=want_grad, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        out = mt.evaluate(zt, want_ginv=True, want_logdet=True, want_grad=want_grad, out=out)
    e1.record(); e1.synchronize()
    print(f'n={n} want_grad={want_grad}: {e0.elapsed_time(e1) / 3:.3f} ms per eval')
