"""Host-buffer entry point: metric evaluation for latents that live in (pinned) host memory.

This is the call ``bench.py`` times for its end-to-end number: every step copies the
step's latents host->device, runs the fused evaluation through the C ABI and copies
log det G and grad_z log det G back device->host.  The batch is cut into chunks that
are double-buffered over two CUDA streams so that the PCIe copies of chunk i+1 / i-1
overlap the kernels of chunk i.  G^{-1} itself stays on the device (1 KB per point;
consumers -- losses, samplers -- reduce it there).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _capi
from .metric_tensor import MetricTensor


class HostEvaluator:
    def __init__(self, metric: MetricTensor, chunk: int = 1 << 17, want_grad: bool = True,
                 keep_ginv: bool = False):
        self.metric = metric
        self.dev = metric.centroids.device
        self.chunk = int(chunk)
        self.want_grad = want_grad
        self.keep_ginv = keep_ginv
        d = metric.latent_dim
        self.d = d
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(2)]
        self.bufs = []
        need = int(_capi.lib().rlvae_metric_eval_workspace(self.chunk, d))
        for _ in range(2):
            self.bufs.append(dict(
                z=torch.empty((self.chunk, d), device=self.dev),
                ginv=torch.empty((self.chunk, d, d), device=self.dev),
                ld=torch.empty(self.chunk, device=self.dev),
                grad=torch.empty((self.chunk, d), device=self.dev),
                work=torch.empty(max(need, 1), device=self.dev, dtype=torch.uint8)))
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, z_host: torch.Tensor, logdet_host: torch.Tensor, grad_host: Optional[torch.Tensor] = None,
               ginv_dev: Optional[torch.Tensor] = None, ginv_host: Optional[torch.Tensor] = None):
        """The same evaluation WITHOUT the final synchronisation: returns (events, byte counts); the host buffers
        are valid once every event has completed (``HostEvaluator.wait(events)``).  Successive submissions queue
        behind each other on the evaluator's two streams (the device-side chunk buffers are reused in stream order),
        so the copies of batch i + 1 overlap the kernels of batch i -- a streaming caller keeps one batch in flight
        and reads batch i while batch i + 1 runs.  Give successive in-flight batches different host buffers."""
        io = self._run(z_host, logdet_host, grad_host, ginv_dev, ginv_host, sync=False)
        events = []
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            events.append(ev)
        return events, io

    @staticmethod
    def wait(events) -> None:
        for ev in events:
            ev.synchronize()

    def __call__(self, z_host: torch.Tensor, logdet_host: torch.Tensor,
                 grad_host: Optional[torch.Tensor] = None, ginv_dev: Optional[torch.Tensor] = None,
                 ginv_host: Optional[torch.Tensor] = None) -> Dict:
        """z_host [N,d] pinned fp32 -> fills logdet_host [N], grad_host [N,d] and, if given, ginv_host
        [N,d,d] (all pinned; G^{-1} is 4 d^2 bytes per point, so copying it out makes the call PCIe
        bound).  Returns byte counts.  Synchronises before returning."""
        return self._run(z_host, logdet_host, grad_host, ginv_dev, ginv_host, sync=True)

    def _run(self, z_host, logdet_host, grad_host, ginv_dev, ginv_host, sync: bool) -> Dict:
        for name, t in (('logdet_host', logdet_host), ('grad_host', grad_host), ('ginv_host', ginv_host)):
            if t is not None and (t.is_cuda or not t.is_pinned() or not t.is_contiguous() or t.dtype != torch.float32):
                raise RuntimeError(f'HostEvaluator: {name} must be a contiguous pinned fp32 host tensor')
        if z_host.is_cuda or not z_host.is_pinned():
            raise RuntimeError('HostEvaluator expects pinned host tensors')
        n = z_host.shape[0]
        tab = self.metric._tables(self.dev)
        path = self.metric._path()
        embedded = self.metric._embedded(self.dev) is not None
        cur = torch.cuda.current_stream(self.dev)
        h2d = d2h = 0
        for s in self.streams:
            s.wait_stream(cur)
        for i, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            m = hi - lo
            b, st = self.bufs[i & 1], self.streams[i & 1]
            with torch.cuda.stream(st):
                b['z'][:m].copy_(z_host[lo:hi], non_blocking=True)
                h2d += m * self.d * 4
                ginv = ginv_dev[lo:hi] if ginv_dev is not None else b['ginv'][:m]
                if embedded:
                    # latent_dim evaluated as a block of the zero-padded problem (MetricTensor._embedded): the
                    # module's own entry point pads, evaluates and slices
                    ev = self.metric.evaluate(b['z'][:m], want_ginv=True, want_logdet=True, want_grad=self.want_grad)
                    ginv.copy_(ev['ginv'])
                    ld_dev, grad_dev = ev['logdet_g'], ev['grad_logdet_g']
                else:
                    _capi.metric_eval(tab, b['z'][:m], want_ginv=True, want_g=False, want_logdet=True,
                                      want_grad=self.want_grad, path=path,
                                      out=dict(ginv=ginv, logdet_g=b['ld'][:m], grad_logdet_g=b['grad'][:m],
                                               work=b['work']))
                    ld_dev, grad_dev = b['ld'][:m], b['grad'][:m]
                logdet_host[lo:hi].copy_(ld_dev, non_blocking=True)
                d2h += m * 4
                if self.want_grad and grad_host is not None:
                    grad_host[lo:hi].copy_(grad_dev, non_blocking=True)
                    d2h += m * self.d * 4
                if ginv_host is not None:
                    ginv_host[lo:hi].copy_(ginv, non_blocking=True)
                    d2h += m * self.d * self.d * 4
        if sync:
            for s in self.streams:
                cur.wait_stream(s)
            cur.synchronize()
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        return dict(h2d_bytes=h2d, d2h_bytes=d2h)
