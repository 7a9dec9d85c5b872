import contextlib, io, os, sys
sys.path.insert(0, '/root/repo')
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric
dev = torch.device('cuda:0')
sm = make_synthetic_metric(10000, 16, seed=0)
for path in ('direct', 'tensor'):
    mt = MetricTensor(16, device=dev, kernel_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(**sm.as_load_kwargs())
    n = 1 << 17
    z = make_points(n, 16, seed=1).to(dev)
    for want_grad in (False, True):
        mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=want_grad); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); mt.evaluate(z, want_ginv=True, want_logdet=True, want_grad=want_grad); e1.record(); e1.synchronize()
        print(path, 'grad' if want_grad else 'fwd', f'{e0.elapsed_time(e1):.2f} ms for {n} points -> {n / e0.elapsed_time(e1) * 1e3:.3e} evals/s')
