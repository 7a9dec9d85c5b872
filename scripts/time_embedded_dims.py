"""latent dims other than 16 / 64 at scale: MetricTensor.evaluate on the zero-padded tensor path (kernel_path='auto')
against the native CUDA-core kernels (kernel_path='direct').  usage: python scripts/time_embedded_dims.py [K] [N]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor
from rlvae_b200.synthetic import make_points, make_synthetic_metric

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
dev = torch.device('cuda:0')


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for d in (8, 12, 32, 48):
    sm = make_synthetic_metric(K, d, seed=d)
    z = make_points(N, d, seed=1).to(dev)
    res = {}
    for path in ('auto', 'direct'):
        mt = MetricTensor(d, device=dev, kernel_path=path)
        with contextlib.redirect_stdout(io.StringIO()):
            mt.load_pretrained(sm.centroids.clone(), sm.metric_matrices.clone(), temperature=sm.temperature,
                               regularization=sm.regularization)
        n = N if path == 'auto' else min(N, 8192)
        zz = z[:n]
        ms = timed(lambda: mt.evaluate(zz, want_ginv=True, want_logdet=True, want_grad=True))
        res[path] = n / ms * 1e3
        if path == 'auto':
            impl = mt.kernel_info()['implementation']
    print(f'd={d:2d} K={K}: padded tensor path {res["auto"] / 1e6:8.3f} M evals/s   native CUDA-core path '
          f'{res["direct"] / 1e6:8.4f} M evals/s   x{res["auto"] / res["direct"]:.0f}   [{impl[:60]}]', flush=True)
