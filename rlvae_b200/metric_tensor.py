"""Drop-in ``MetricTensor`` backed by the rlvae_b200 CUDA library.

Host-side mirror of ref ``src/models/components/metric_tensor.py`` (class
``MetricTensor`` :21-275): same constructor, buffer names, method names, argument
meaning and error behaviour, so the models, samplers and losses of the
reference (``modular_rlvae.py:239-261``, ``loss_manager.py:106-113``,
``base_sampler.py:60-68``) can use it unchanged.  What differs is underneath:

* ``compute_inverse_metric``  -> ``rlvae_inverse_metric``  (tcgen05 split-fp16 / 3xTF32 kernels
  or the fp32 direct kernel; never materialises [N,K] or [N,K,d,d]);
* ``compute_metric`` / ``compute_log_det_metric`` -> ``rlvae_metric_eval`` (Cholesky fused into the
  tensor kernel's epilogue) or ``rlvae_batched_inverse`` (register/shuffle Gauss-Jordan);
* backward w.r.t. ``z`` -> ``rlvae_metric_grad`` (closed-form contraction), wired
  through ``torch.autograd.Function`` so ``loss_manager.py:110-142`` and
  ``riemannian_flow_vae.py:1030-1069`` can back-propagate through the metric.

There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import math
import warnings
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import _capi


def _pad_cols(z: torch.Tensor, dp: int) -> torch.Tensor:
    zp = torch.zeros((z.shape[0], dp), device=z.device, dtype=z.dtype)
    zp[:, :z.shape[1]] = z
    return zp


class _InverseMetricFn(torch.autograd.Function):
    """G^{-1}(z); backward dL/dz = (2/T^2) sum_k w_k <dL/dG^{-1}, M_k> (c_k - z)."""

    @staticmethod
    def forward(ctx, z, owner, path):
        zc = z.detach().contiguous()
        emb = owner._embedded(z.device) if hasattr(owner, '_embedded') else None
        if emb is not None:
            # latent_dim without a tensor kernel of its own: evaluated as a block of the zero-padded problem
            d = zc.shape[1]
            zp = _pad_cols(zc, emb.d)
            ctx.tab, ctx.path, ctx.embed_d = emb, path, d
            ctx.save_for_backward(zp)
            return _capi.inverse_metric(emb, zp, path)[:, :d, :d].contiguous()
        tab = owner._tables(z.device)
        ctx.tab, ctx.path, ctx.embed_d = tab, path, None
        ctx.save_for_backward(zc)
        return _capi.inverse_metric(tab, zc, path)

    @staticmethod
    def backward(ctx, grad_out):
        (zc,) = ctx.saved_tensors
        tab = ctx.tab
        u = grad_out.contiguous()
        if ctx.embed_d is not None:
            d = ctx.embed_d
            up = torch.zeros((u.shape[0], tab.d, tab.d), device=u.device, dtype=u.dtype)
            up[:, :d, :d] = u
            gz = _capi.metric_grad(tab, zc, up, 2.0 / (tab.temperature ** 2), ctx.path)
            return gz[:, :d].contiguous(), None, None
        gz = _capi.metric_grad(tab, zc, u, 2.0 / (tab.temperature ** 2), ctx.path)
        return gz, None, None


class _InverseFn(torch.autograd.Function):
    """A -> (A^{-1}, sign det A) (batched); backward dL/dA = -A^{-T} (dL/dA^{-1}) A^{-T}."""

    @staticmethod
    def forward(ctx, a):
        inv, _, sgn, _ = _capi.batched_inverse(a.detach(), want_inv=True, want_sign=True)
        ctx.save_for_backward(inv)
        ctx.mark_non_differentiable(sgn)
        return inv, sgn

    @staticmethod
    def backward(ctx, grad_out, _grad_sign):
        (inv,) = ctx.saved_tensors
        it = inv.transpose(-1, -2)
        return -(it @ grad_out @ it)


class _LogAbsDetFn(torch.autograd.Function):
    """A -> log|det A| (batched); backward dL/dA = g * A^{-T}."""

    @staticmethod
    def forward(ctx, a):
        inv, lad, _, _ = _capi.batched_inverse(a.detach(), want_inv=True, want_logabsdet=True)
        ctx.save_for_backward(inv)
        return lad

    @staticmethod
    def backward(ctx, grad_out):
        (inv,) = ctx.saved_tensors
        return grad_out[:, None, None] * inv.transpose(-1, -2)


class MetricTensor(nn.Module):
    """G^{-1}(z) = sum_k M_k exp(-||z - c_k||^2 / T^2) + lambda I,  G = (G^{-1})^{-1}.

    Buffers (names are part of the contract, ref metric_tensor.py:50-53):
    ``centroids [K,d]``, ``metric_matrices [K,d,d]``, ``temperature``, ``regularization``.
    ``kernel_path`` selects the implementation: 'auto' | 'direct' | 'tensor'.
    """

    _PATHS = {'auto': _capi.PATH_AUTO, 'direct': _capi.PATH_DIRECT, 'tensor': _capi.PATH_TENSOR}

    def __init__(self, latent_dim: int, temperature: float = 0.1, regularization: float = 0.01,
                 device: Optional[torch.device] = None, kernel_path: str = 'auto'):
        super().__init__()
        self.latent_dim = latent_dim
        self.device = device or torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.register_buffer('centroids', torch.empty(0, latent_dim))
        self.register_buffer('metric_matrices', torch.empty(0, latent_dim, latent_dim))
        self.register_buffer('temperature', torch.tensor(temperature))
        self.register_buffer('regularization', torch.tensor(regularization))
        self._is_loaded = False
        self._diagnostic_counter = 0
        self.kernel_path = kernel_path
        self.check_singular = True   # one [N]-sized device->host check per compute_metric call
        self._tab = None
        self._tab_key = None
        self._emb = None             # zero-padded tables (latent dims other than 16 / 64), see _embedded()
        self._emb_key = None

    # ------------------------------------------------------------------ loading / caches
    def load_pretrained(self, centroids: torch.Tensor, metric_matrices: torch.Tensor,
                        temperature: Optional[float] = None,
                        regularization: Optional[float] = None) -> None:
        """Install pretrained tables (accepts ``**MetricLoader.load_from_file(...)``).
        Shape errors are ``ValueError`` like ref metric_tensor.py:76-83."""
        if centroids.shape[1] != self.latent_dim:
            raise ValueError(f'Centroids dimension {centroids.shape[1]} != latent_dim {self.latent_dim}')
        if metric_matrices.shape[0] != centroids.shape[0]:
            raise ValueError(f'Number of metric matrices {metric_matrices.shape[0]} != number of '
                             f'centroids {centroids.shape[0]}')
        if tuple(metric_matrices.shape[1:]) != (self.latent_dim, self.latent_dim):
            raise ValueError(f'Metric matrix shape {tuple(metric_matrices.shape[1:])} != '
                             f'({self.latent_dim}, {self.latent_dim})')
        self.register_buffer('centroids', centroids.to(self.device))
        self.register_buffer('metric_matrices', metric_matrices.to(self.device))
        if temperature is not None:
            self.register_buffer('temperature', torch.tensor(temperature, device=self.device))
        if regularization is not None:
            self.register_buffer('regularization', torch.tensor(regularization, device=self.device))
        self._is_loaded = True
        self._tab = None
        self._emb_key = None
        print(f'✅ MetricTensor loaded: {len(centroids)} centroids, T={self.temperature.item():.3f}, '
              f'λ={self.regularization.item():.3f}')

    def _tables(self, device) -> _capi.Tables:
        """Packed device tables: a derived cache of the four buffers, rebuilt whenever
        they change (load_pretrained, load_state_dict, .to(), in-place edits)."""
        c, m = self.centroids, self.metric_matrices
        if c.device != device:
            raise RuntimeError(f'MetricTensor tables live on {c.device} but z is on {device}; '
                               'call .to(device) on the module first')
        key = (c.data_ptr(), c._version, m.data_ptr(), m._version, tuple(c.shape),
               self.temperature.data_ptr(), self.temperature._version,
               self.regularization.data_ptr(), self.regularization._version)
        if self._tab is None or self._tab_key != key:
            if self._tab is not None:
                # kernels may still read the old packed copy; an autograd graph may still hold the old
                # handle (ctx.tab), so it is NOT closed here -- the last reference frees it (__del__)
                torch.cuda.synchronize(device)
            self._tab = _capi.Tables(c.float(), m.float(), float(self.temperature.item()),
                                     float(self.regularization.item()))
            self._tab_key = key
        return self._tab

    # latent dims other than 16 / 64 have no tensor kernel of their own.  With K >= EMBED_MIN_CENTROIDS centroids the
    # metric is evaluated as the leading d x d block of the same problem zero-padded to 16 / 64 dims: distances and
    # weights are unchanged, G^{-1} = diag(G^{-1}_d, lambda I), so G, the gradient and the backward are the leading
    # blocks and log det G = log det G_padded + (dp - d) log lambda.  (The samplers keep the native tables: the HMC
    # quirks -- det + clamp(1e-10) -- do not commute with the padding.)
    EMBED_MIN_CENTROIDS = 512

    def _table_key(self):
        c, m = self.centroids, self.metric_matrices
        return (c.data_ptr(), c._version, m.data_ptr(), m._version, tuple(c.shape),
                self.temperature.data_ptr(), self.temperature._version,
                self.regularization.data_ptr(), self.regularization._version)

    def _embedded(self, device) -> Optional[_capi.Tables]:
        d = self.latent_dim
        if d in (16, 64) or d > 64 or self.kernel_path == 'direct' or not self._is_loaded:
            return None
        c, m = self.centroids, self.metric_matrices
        if c.device != device or c.shape[0] < self.EMBED_MIN_CENTROIDS:
            return None
        key = self._table_key()
        if self._emb_key != key:
            self._emb, self._emb_key = None, key
            lam = float(self.regularization.item())
            if lam > 0.0 and math.isfinite(lam):
                dp = 16 if d < 16 else 64
                cp = torch.zeros((c.shape[0], dp), device=device, dtype=torch.float32)
                cp[:, :d] = c
                mp = torch.zeros((c.shape[0], dp, dp), device=device, dtype=torch.float32)
                mp[:, :d, :d] = m
                tab = _capi.Tables(cp, mp, float(self.temperature.item()), lam)
                if tab.tensor_capable and (tab.tensor_auto or self.kernel_path == 'tensor'):
                    self._emb = tab
        return self._emb

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # buffers registered empty must take the checkpoint's shapes (ref keeps tables in state_dict)
        for name in ('centroids', 'metric_matrices'):
            k = prefix + name
            if k in state_dict and state_dict[k].shape != getattr(self, name).shape:
                setattr(self, name, torch.empty_like(state_dict[k], device=getattr(self, name).device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        if self.centroids.numel() > 0:
            self._is_loaded = True
        self._tab = None
        self._emb_key = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.device = self.centroids.device
        self._tab = None
        self._emb_key = None
        return out

    def _path(self) -> int:
        return self._PATHS[self.kernel_path]

    def _check_ready(self, z):
        if not self._is_loaded:
            raise RuntimeError('Metric tensor not loaded. Call load_pretrained() first.')
        if z.dim() != 2 or z.shape[1] != self.latent_dim:
            raise ValueError(f'z must be [batch, {self.latent_dim}], got {tuple(z.shape)}')
        if not z.is_cuda:
            raise RuntimeError('rlvae_b200.MetricTensor evaluates on CUDA only (no CPU fallback); '
                               f'got z on {z.device}')

    # ------------------------------------------------------------------ A2 / A3 / A4 / A20
    def compute_inverse_metric(self, z: torch.Tensor) -> torch.Tensor:
        """G^{-1}(z) [N,d,d]  (ref metric_tensor.py:98-137)."""
        self._check_ready(z)
        return _InverseMetricFn.apply(z.float(), self, self._path())

    def compute_metric(self, z: torch.Tensor) -> torch.Tensor:
        """G(z) = inv(G^{-1}(z)) [N,d,d]  (ref metric_tensor.py:139-160).  A singular G^{-1}
        is retried once with +1e-6 I, like the reference's LinAlgError handler."""
        if not (torch.is_grad_enabled() and z.requires_grad):
            # no graph to build: one fused evaluation (packed Cholesky for symmetric tables).  The
            # singularity check reads the [N] log-determinants the same kernel produces (a zero pivot
            # gives log|det G^{-1}| = -inf), not the [N,d,d] result.
            ev = self.evaluate(z, want_ginv=False, want_g=True, want_logdet=self.check_singular)
            if not self.check_singular or bool(torch.isfinite(ev['logdet_g']).all()):
                return ev['g']
        g_inv = self.compute_inverse_metric(z)
        g, sgn = _InverseFn.apply(g_inv)
        if self.check_singular and bool((sgn == 0).any()):   # exact zero pivot == LinAlgError
            warnings.warn('Metric tensor inversion failed: singular matrix. Adding regularization.')
            eye = torch.eye(self.latent_dim, device=z.device, dtype=g_inv.dtype).unsqueeze(0)
            g, _ = _InverseFn.apply(g_inv + 1e-6 * eye)
        return g

    def compute_log_det_metric(self, z: torch.Tensor) -> torch.Tensor:
        """log|det G(z)| [N]  (ref metric_tensor.py:162-182) = -log|det G^{-1}(z)|."""
        if not (torch.is_grad_enabled() and z.requires_grad):
            return self.evaluate(z, want_ginv=False, want_g=False, want_logdet=True)['logdet_g']
        return -_LogAbsDetFn.apply(self.compute_inverse_metric(z))

    def compute_riemannian_distance_squared(self, z1: torch.Tensor, z2: torch.Tensor) -> torch.Tensor:
        """(z1-z2)^T G((z1+z2)/2) (z1-z2) [N]  (ref metric_tensor.py:184-207)."""
        g_mid = self.compute_metric(0.5 * (z1 + z2))
        dz = z1 - z2
        return torch.einsum('bi,bij,bj->b', dz, g_mid, dz)

    # ------------------------------------------------------------------ fused / extended entry points
    def evaluate(self, z: torch.Tensor, want_ginv=True, want_g=False, want_logdet=True,
                 want_grad=False, out=None) -> Dict[str, Optional[torch.Tensor]]:
        """One fused 'metric eval' (no autograd): any subset of G^{-1}, G, log det G and the
        analytic grad_z log det G, through ``rlvae_metric_eval``."""
        self._check_ready(z)
        with torch.no_grad():
            emb = self._embedded(z.device)
            if emb is not None:
                d = self.latent_dim
                ev = _capi.metric_eval(emb, _pad_cols(z.float(), emb.d), want_ginv, want_g, want_logdet, want_grad,
                                       self._path(), None)
                blk = lambda t: None if t is None else t[:, :d, :d].contiguous()
                ld = ev['logdet_g']
                if ld is not None:      # log det G_d = log det G_padded + (dp - d) log lambda
                    ld = (ld.double() + (emb.d - d) * math.log(float(self.regularization.item()))).float()
                gr = ev['grad_logdet_g']
                return dict(ginv=blk(ev['ginv']), g=blk(ev['g']), logdet_g=ld,
                            grad_logdet_g=None if gr is None else gr[:, :d].contiguous(), work=None)
            return _capi.metric_eval(self._tables(z.device), z.float(), want_ginv, want_g, want_logdet,
                                     want_grad, self._path(), out)

    def compute_metric_spectrum(self, z: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Per-point spectrum of the metric, what the flow-analysis consumers compute per step with
        torch.linalg.eigvals / det (ref flow_analysis.py:104-126, manifold.py:79-101,
        modular_rlvae.py:434-457): eigenvalues of G^{-1} and G (ascending), condition number, traces,
        log det.  Symmetric tables with d == 16 run one fused pass (forward kernel + per-thread Jacobi);
        anything else composes compute_inverse_metric with torch.linalg on the device."""
        self._check_ready(z)
        with torch.no_grad():
            tab = self._tables(z.device)
            if tab.d == 16 and tab.symmetric:
                eig_inv, ld = _capi.metric_spectrum(tab, z.float(), True, self._path())
            else:
                g_inv = self.compute_inverse_metric(z)
                if tab.symmetric:
                    eig_inv = torch.linalg.eigvalsh(g_inv)
                else:
                    eig_inv = torch.sort(torch.linalg.eigvals(g_inv).real, dim=-1).values
                ld = -torch.linalg.slogdet(g_inv).logabsdet
            eig_g = torch.flip(1.0 / eig_inv, dims=[-1])
            return {'eigenvals_G_inv': eig_inv, 'eigenvals_G': eig_g,
                    'condition_number': eig_inv[:, -1] / eig_inv[:, 0],
                    'trace_G_inv': eig_inv.sum(-1), 'trace_G': eig_g.sum(-1),
                    'logdet_G': ld, 'det_G': torch.exp(ld), 'det_G_inv': torch.exp(-ld)}

    def compute_grad_log_det_metric(self, z: torch.Tensor) -> torch.Tensor:
        """Analytic grad_z log det G(z) [N,d] (north_star)."""
        return self.evaluate(z, want_ginv=False, want_logdet=False, want_grad=True)['grad_logdet_g']

    # ------------------------------------------------------------------ diagnostics (host side)
    def diagnose_metric_properties(self, z: torch.Tensor, verbose: bool = False) -> Dict[str, Any]:
        """Same dictionary as ref metric_tensor.py:209-261 (keys consumed by hybrid_rlvae.py:324-331)."""
        with torch.no_grad():
            g = self.compute_metric(z)
            g_inv = self.compute_inverse_metric(z)
            ev_g = torch.linalg.eigvals(g[0].cpu()).real
            ev_gi = torch.linalg.eigvals(g_inv[0].cpu()).real
            _, lad, sgn, _ = _capi.batched_inverse(g_inv, want_inv=False, want_logabsdet=True, want_sign=True)
            det_gi = sgn * torch.exp(lad)
            det_g = sgn * torch.exp(-lad)
            tr_g = torch.diagonal(g, dim1=-2, dim2=-1).sum(-1)
            tr_gi = torch.diagonal(g_inv, dim1=-2, dim2=-1).sum(-1)
            d = {
                'eigenvals_G_min': ev_g.min().item(), 'eigenvals_G_max': ev_g.max().item(),
                'eigenvals_G_mean': ev_g.mean().item(),
                'eigenvals_G_inv_min': ev_gi.min().item(), 'eigenvals_G_inv_max': ev_gi.max().item(),
                'eigenvals_G_inv_mean': ev_gi.mean().item(),
                'condition_number_G': (ev_g.max() / (ev_g.min() + 1e-8)).item(),
                'condition_number_G_inv': (ev_gi.max() / (ev_gi.min() + 1e-8)).item(),
                'det_G_mean': det_g.mean().item(), 'det_G_inv_mean': det_gi.mean().item(),
                'trace_G_mean': tr_g.mean().item(), 'trace_G_inv_mean': tr_gi.mean().item(),
                'batch_size': z.shape[0], 'n_centroids': len(self.centroids),
                'temperature': self.temperature.item(), 'regularization': self.regularization.item(),
            }
            if verbose:
                print('🔍 METRIC DIAGNOSTICS:')
                print(f"   G eigenvalues: min={d['eigenvals_G_min']:.3e}, max={d['eigenvals_G_max']:.3e}, "
                      f"mean={d['eigenvals_G_mean']:.3e}")
                print(f"   G condition number: {d['condition_number_G']:.2e}")
                print(f"   det(G): mean={d['det_G_mean']:.3e}")
                print(f"   trace(G): mean={d['trace_G_mean']:.3e}")
                print(f"   Batch size: {d['batch_size']}, Centroids: {d['n_centroids']}")
            return d

    def kernel_info(self) -> Dict[str, Any]:
        """Which implementation `kernel_path='auto'` resolves to for the loaded tables."""
        if not self._is_loaded or not self.centroids.is_cuda:
            return {'loaded': self._is_loaded, 'device': str(self.centroids.device)}
        emb = self._embedded(self.centroids.device)
        tab = emb if emb is not None else self._tables(self.centroids.device)
        tensor = tab.tensor_capable and tab.tensor_auto and self.kernel_path != 'direct'

        if self.kernel_path == 'tensor':
            tensor = tab.tensor_capable
        if not tensor:
            kind = 'direct CUDA-core kernels'
        elif tab.d == 64:
            kind = 'split-fp16 tcgen05 kernels, column-tiled (forward: 17 tiles of 128 packed columns; gradient: partial tiles + reduction)'
        elif tab.symmetric:
            kind = ('split-fp16 tcgen05 kernels, ' +
                    ('expanded-distance GEMM', 'exact-distance mode (small temperature)',
                     'hybrid mode (small temperature: expanded-distance GEMM + exact refinement of the live weights)')[tab.weight_mode])
        else:
            kind = '3xTF32 tcgen05 kernels (non-symmetric tables)'
        if emb is not None:
            kind += f' (latent_dim {self.latent_dim} evaluated as a block of the problem zero-padded to {emb.d} dims)'
        return {'loaded': True, 'device': str(self.centroids.device), 'latent_dim': self.latent_dim, 'n_centroids': tab.K,
                'symmetric_tables': tab.symmetric, 'tensor_path': bool(tensor), 'expanded_form_accurate': tab.expanded_ok,
                'implementation': kind, 'kernel_path': self.kernel_path}

    def is_loaded(self) -> bool:
        return self._is_loaded

    def get_config(self) -> Dict[str, Any]:
        return {
            'latent_dim': self.latent_dim,
            'temperature': self.temperature.item() if self._is_loaded else None,
            'regularization': self.regularization.item() if self._is_loaded else None,
            'n_centroids': len(self.centroids) if self._is_loaded else 0,
            'is_loaded': self._is_loaded,
        }
