import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlvae_b200 import MetricTensor, _capi
from rlvae_b200.synthetic import make_points, make_synthetic_metric
from oracle import metric_oracle as O
dev = torch.device('cuda:0')
T = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
sm = make_synthetic_metric(10000, 16, seed=0)
t = (sm.centroids, sm.metric_matrices, T, sm.regularization)
c = t[0]
near = c[:384] + 0.3 * T * make_points(384, 16, seed=2) / 4.0
z = torch.cat([make_points(128, 16, seed=1), near]).contiguous()
mt = MetricTensor(16, device=dev)
with contextlib.redirect_stdout(io.StringIO()):
    mt.load_pretrained(c.clone(), t[1].clone(), temperature=T, regularization=t[3])
tab = mt._tables(dev)
print('mode', tab.weight_mode, 'lambda', t[3])
zd = z.to(dev)
ref = O.chunked(O.grad_pythae, z, *t, chunk=128).reshape(z.shape)
got, lad, sgn = _capi.pythae_eval(tab, zd, path=_capi.PATH_TENSOR)
gd, _, _ = _capi.pythae_eval(tab, zd, path=_capi.PATH_DIRECT)
err = (got.cpu() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)
errd = (gd.cpu() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)
rn = ref.norm(dim=1)
live = rn > 1e-6 * rn.max()
print('live', int(live.sum()), 'max err live tc %.3e direct %.3e' % (err[live].max(), errd[live].max()))
order = err.argsort(descending=True)[:12]
for i in order.tolist():
    print(i, 'err_tc %.3e err_direct %.3e |ref| %.3e' % (err[i], errd[i], rn[i]))
print('max |ref|', rn.max().item())
# pieces
w = O.centroid_weights(z, t[0], T)
print('sum w for worst rows', w[order].sum(1))
