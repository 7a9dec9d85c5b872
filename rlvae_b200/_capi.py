"""ctypes binding of librlvae_b200.so (the C ABI in include/rlvae_b200.h).

This is the only module that touches the shared library.  There is NO CPU
fallback: if the library is missing, or a call is made with non-CUDA tensors,
it raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``rlvae_b200.build.build()``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RLVAE_B200_LIB: another build of the same ABI (A/B timing of kernel variants on one box; never a fallback)
LIB_PATH = os.environ.get('RLVAE_B200_LIB') or os.path.join(_HERE, 'lib', 'librlvae_b200.so')

PATH_AUTO, PATH_DIRECT, PATH_TENSOR = 0, 1, 2
GRAD_MODULAR, GRAD_EXACT = 0, 1
HMC_NO_FUSION = 256      # flag OR-ed into grad_mode: per-step launches instead of the fused trajectory kernel

# every symbol include/rlvae_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ('rlvae_last_error', c_char_p, []),
    ('rlvae_abi_version', c_int, []),
    ('rlvae_launch_count', ctypes.c_longlong, [c_int]),
    ('rlvae_profile_begin', c_int, [c_int]),
    ('rlvae_profile_count', c_int, []),
    ('rlvae_profile_read', c_int, [c_int, POINTER(c_float)]),
    ('rlvae_profile_end', c_int, []),
    ('rlvae_tables_create', c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p]),
    ('rlvae_tables_destroy', c_int, [c_void_p]),
    ('rlvae_tables_info', c_int, [c_void_p, POINTER(c_int64)]),
    ('rlvae_inverse_metric_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_inverse_metric', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_inverse_metric_packed', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    ('rlvae_batched_inverse', c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_metric_grad', c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_int, c_void_p]),
    ('rlvae_metric_grad_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_metric_grad_ws', c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_metric_grad_pythae_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_metric_grad_pythae', c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_pythae_eval_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_pythae_eval', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_metric_eval_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_metric_eval', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_sym_eigvalsh', c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    ('rlvae_metric_spectrum', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ('rlvae_local_covariance', c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    ('rlvae_hmc_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_hmc_iteration', c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float,
                                    POINTER(c_float), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    ('rlvae_hmc_fused_available', c_int, [c_void_p, c_int, c_int]),
    ('rlvae_hmc_run_workspace', c_int64, [c_int64, c_int, c_int, c_int]),
    ('rlvae_hmc_run', c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_float,
                              POINTER(c_float), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_void_p]),
    ('rlvae_pythae_hmc_workspace', c_int64, [c_int64, c_int]),
    ('rlvae_pythae_hmc_run', c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_float,
                                     POINTER(c_float), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_void_p]),
    ('rlvae_hmc_refine', c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_int, c_void_p]),
    ('rlvae_nearest2', c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    ('rlvae_chol_apply', c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f'rlvae_b200: CUDA library not built ({LIB_PATH} missing). '
                'Run `python -c "import __graft_entry__ as g; g.build()"` first; there is no CPU fallback.')
        h = ctypes.CDLL(LIB_PATH)
        for name, res, args in _SIGNATURES:
            fn = getattr(h, name)       # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = lib().rlvae_last_error()
        raise RuntimeError(f'{what} failed (code {rc}): {msg.decode() if msg else "?"}')


def _stream(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name}: expected a torch.Tensor')
    if not t.is_cuda:
        raise RuntimeError(f'{name}: rlvae_b200 kernels need CUDA tensors (got {t.device}); '
                           'there is no CPU fallback')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    return t.contiguous()


def _req_z(tab: 'Tables', z: torch.Tensor, name: str = 'z') -> torch.Tensor:
    """A latent batch for tables ``tab``: contiguous fp32 [N, tab.d] on the tables' device.  Everything
    below this line is raw pointers, so a wrong latent_dim or device must fail HERE, not in a kernel."""
    z = _req(z, name)
    if z.dim() != 2 or z.shape[1] != tab.d:
        raise ValueError(f'{name} must be [batch, {tab.d}], got {tuple(z.shape)}')
    if z.device != tab.device:
        raise RuntimeError(f'{name} is on {z.device} but the metric tables live on {tab.device}')
    return z


def _req_out(t, name: str, shape, device, dtype=torch.float32):
    """A caller-supplied output buffer: exact shape, dtype, device, contiguous (kernels write through
    the raw pointer).  None passes through."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != device:
        raise RuntimeError(f'{name}: output buffer must be a CUDA tensor on {device}')
    if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f'{name}: output buffer must be contiguous {dtype} of shape {tuple(shape)}, '
                         f'got {t.dtype} {tuple(t.shape)} (contiguous={t.is_contiguous()})')
    return t


class Tables:
    """Owner of a ``rlvae_tables_t`` handle (packed device copies of the metric tables)."""

    def __init__(self, centroids: torch.Tensor, matrices: torch.Tensor, temperature: float,
                 regularization: float):
        c = _req(centroids, 'centroids')
        m = _req(matrices, 'metric_matrices')
        if c.dim() != 2 or m.dim() != 3 or m.shape != (c.shape[0], c.shape[1], c.shape[1]):
            raise ValueError(f'inconsistent table shapes: centroids {tuple(c.shape)}, matrices {tuple(m.shape)}')
        self.device = c.device
        self.K, self.d = int(c.shape[0]), int(c.shape[1])
        self.temperature, self.regularization = float(temperature), float(regularization)
        h = c_void_p()
        with torch.cuda.device(self.device):
            _check(lib().rlvae_tables_create(ctypes.byref(h), _ptr(c), _ptr(m), self.K, self.d,
                                             c_float(self.temperature), c_float(self.regularization),
                                             _stream(c)), 'rlvae_tables_create')
        self._h = h
        info = (c_int64 * 12)()
        _check(lib().rlvae_tables_info(self._h, info), 'rlvae_tables_info')
        self.Kpad, self.symmetric = int(info[2]), bool(info[3])
        self.tensor_capable, self.tensor_auto = bool(info[4]), bool(info[5])
        self.expanded_ok = bool(info[6])     # accuracy gate of the expanded distance form
        self.weight_mode = int(info[7])      # d = 16 symmetric: 0 expanded form, 1 exact differences, 2 hybrid
        self.psd_certified = bool(info[8])   # every M_k PSD and lambda > 0: G^{-1}(z) positive definite everywhere

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError('rlvae_b200: tables handle already destroyed')
        return self._h

    def close(self):
        if getattr(self, '_h', None) is not None and _lib is not None:
            _lib.rlvae_tables_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- thin call wrappers
def inverse_metric(tab: Tables, z: torch.Tensor, path: int = PATH_AUTO) -> torch.Tensor:
    z = _req_z(tab, z)
    n, d = z.shape
    out = torch.empty((n, d, d), device=z.device, dtype=torch.float32)
    need = int(lib().rlvae_inverse_metric_workspace(n, d))
    work = torch.empty(need, device=z.device, dtype=torch.uint8) if need > 0 else None
    with torch.cuda.device(z.device):
        _check(lib().rlvae_inverse_metric(tab.handle, _ptr(z), n, _ptr(out), _ptr(work), path, _stream(z)),
               'rlvae_inverse_metric')
    return out


def inverse_metric_packed(tab: Tables, z: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """G^{-1} in the packed symmetric [N,144] layout (symmetric tables, d == 16)."""
    z = _req_z(tab, z)
    n = z.shape[0]
    if out is None:
        out = torch.empty((n, 144), device=z.device, dtype=torch.float32)
    _req_out(out, 'out', (n, 144), z.device)
    with torch.cuda.device(z.device):
        _check(lib().rlvae_inverse_metric_packed(tab.handle, _ptr(z), n, _ptr(out), _stream(z)),
               'rlvae_inverse_metric_packed')
    return out


def batched_inverse(a: torch.Tensor, want_inv=True, want_logabsdet=False, want_sign=False,
                    want_diag=False, transpose=False):
    a = _req(a, 'a')
    if a.dim() != 3 or a.shape[1] != a.shape[2]:
        raise ValueError(f'a must be [batch, d, d], got {tuple(a.shape)}')
    n, d = a.shape[0], a.shape[-1]
    dev = a.device
    inv = torch.empty_like(a) if want_inv else None
    lad = torch.empty(n, device=dev) if want_logabsdet else None
    sgn = torch.empty(n, device=dev) if want_sign else None
    diag = torch.empty((n, d), device=dev) if want_diag else None
    with torch.cuda.device(dev):
        _check(lib().rlvae_batched_inverse(_ptr(a), n, d, _ptr(inv), _ptr(lad), _ptr(sgn), _ptr(diag),
                                           1 if transpose else 0, _stream(a)), 'rlvae_batched_inverse')
    return inv, lad, sgn, diag


def metric_grad(tab: Tables, z: torch.Tensor, u: torch.Tensor, scale: float, path: int = PATH_AUTO):
    z = _req_z(tab, z)
    u = _req(u, 'u')
    if tuple(u.shape) != (z.shape[0], tab.d, tab.d) or u.device != z.device:
        raise ValueError(f'u must be [{z.shape[0]}, {tab.d}, {tab.d}] on {z.device}, got {tuple(u.shape)} on {u.device}')
    out = torch.empty_like(z)
    need = int(lib().rlvae_metric_grad_workspace(z.shape[0], tab.d))
    work = torch.empty(need, device=z.device, dtype=torch.uint8) if need > 0 else None
    with torch.cuda.device(z.device):
        _check(lib().rlvae_metric_grad_ws(tab.handle, _ptr(z), _ptr(u), z.shape[0], c_float(scale), _ptr(out),
                                          _ptr(work), path, _stream(z)), 'rlvae_metric_grad')
    return out


def metric_grad_pythae(tab: Tables, z: torch.Tensor, g: torch.Tensor, path: int = PATH_AUTO):
    z = _req_z(tab, z)
    g = _req(g, 'g')
    if tuple(g.shape) != (z.shape[0], tab.d, tab.d) or g.device != z.device:
        raise ValueError(f'g must be [{z.shape[0]}, {tab.d}, {tab.d}] on {z.device}, got {tuple(g.shape)} on {g.device}')
    out = torch.empty_like(z)
    need = int(lib().rlvae_metric_grad_pythae_workspace(z.shape[0], tab.d))
    work = torch.empty(max(need, 1), device=z.device, dtype=torch.uint8)
    with torch.cuda.device(z.device):
        _check(lib().rlvae_metric_grad_pythae(tab.handle, _ptr(z), _ptr(g), z.shape[0], _ptr(out), _ptr(work),
                                              path, _stream(z)), 'rlvae_metric_grad_pythae')
    return out


def pythae_eval(tab: Tables, z: torch.Tensor, path: int = PATH_AUTO, work: Optional[torch.Tensor] = None):
    """-> (grad [N,d], logabsdet [N], sign [N]): variant-C gradient, log|det G^{-1}(z)| and its sign."""
    z = _req_z(tab, z)
    n = z.shape[0]
    grad = torch.empty_like(z)
    lad = torch.empty(n, device=z.device, dtype=torch.float32)
    sgn = torch.empty(n, device=z.device, dtype=torch.float32)
    need = int(lib().rlvae_pythae_eval_workspace(n, tab.d))
    if work is None or work.numel() < need or work.device != z.device or work.dtype != torch.uint8:
        work = torch.empty(max(need, 1), device=z.device, dtype=torch.uint8)
    with torch.cuda.device(z.device):
        _check(lib().rlvae_pythae_eval(tab.handle, _ptr(z), n, _ptr(grad), _ptr(lad), _ptr(sgn), _ptr(work), path,
                                       _stream(z)), 'rlvae_pythae_eval')
    return grad, lad, sgn


def metric_eval(tab: Tables, z: torch.Tensor, want_ginv=True, want_g=False, want_logdet=True,
                want_grad=False, path: int = PATH_AUTO, out=None):
    """-> dict(ginv, g, logdet_g, grad_logdet_g) (missing keys are None)."""
    z = _req_z(tab, z)
    n, d = z.shape
    dev = z.device
    out = out or {}
    ginv = _req_out(out.get('ginv'), 'out[ginv]', (n, d, d), dev) if want_ginv else None
    if want_ginv and ginv is None:
        ginv = torch.empty((n, d, d), device=dev)
    g = _req_out(out.get('g'), 'out[g]', (n, d, d), dev) if want_g else None
    if want_g and g is None:
        g = torch.empty((n, d, d), device=dev)
    ld = _req_out(out.get('logdet_g'), 'out[logdet_g]', (n,), dev) if want_logdet else None
    if want_logdet and ld is None:
        ld = torch.empty(n, device=dev)
    gr = _req_out(out.get('grad_logdet_g'), 'out[grad_logdet_g]', (n, d), dev) if want_grad else None
    if want_grad and gr is None:
        gr = torch.empty((n, d), device=dev)
    work = out.get('work')
    need = int(lib().rlvae_metric_eval_workspace(n, d))
    if work is not None and (not work.is_cuda or work.device != dev or work.dtype != torch.uint8
                             or not work.is_contiguous()):
        work = None
    if work is None or work.numel() < need:
        work = torch.empty(max(need, 1), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _check(lib().rlvae_metric_eval(tab.handle, _ptr(z), n, _ptr(ginv), _ptr(g), _ptr(ld), _ptr(gr),
                                       _ptr(work), path, _stream(z)), 'rlvae_metric_eval')
    return dict(ginv=ginv, g=g, logdet_g=ld, grad_logdet_g=gr, work=work)


def sym_eigvalsh(a: torch.Tensor) -> torch.Tensor:
    """Eigenvalues (ascending) of symmetric matrices a [N,d,d].  d == 16 runs the per-thread Jacobi
    kernel; other sizes use torch.linalg.eigvalsh on the device (a library call, not a CPU path)."""
    a = _req(a, 'a')
    n, d = a.shape[0], a.shape[-1]
    if d != 16:
        return torch.linalg.eigvalsh(a)
    out = torch.empty((n, d), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _check(lib().rlvae_sym_eigvalsh(_ptr(a), n, d, 0, _ptr(out), _stream(a)), 'rlvae_sym_eigvalsh')
    return out


def metric_spectrum(tab: Tables, z: torch.Tensor, want_logdet: bool = True, path: int = PATH_AUTO):
    """-> (eig(G^{-1}(z)) [N,16] ascending, log|det G| [N] or None); symmetric tables, d == 16."""
    z = _req_z(tab, z)
    n, d = z.shape
    eig = torch.empty((n, d), device=z.device, dtype=torch.float32)
    ld = torch.empty(n, device=z.device, dtype=torch.float32) if want_logdet else None
    work = torch.empty(max(int(lib().rlvae_metric_eval_workspace(n, d)), 1), device=z.device, dtype=torch.uint8)
    with torch.cuda.device(z.device):
        _check(lib().rlvae_metric_spectrum(tab.handle, _ptr(z), n, _ptr(eig), _ptr(ld), _ptr(work), path,
                                           _stream(z)), 'rlvae_metric_spectrum')
    return eig, ld


def local_covariance(latents: torch.Tensor, centroids: torch.Tensor, temperature: float) -> torch.Tensor:
    """cov [K,d,d]: weighted covariance of the latents around every centroid (metric construction)."""
    x = _req(latents, 'latents')
    c = _req(centroids, 'centroids')
    if x.dim() != 2 or c.dim() != 2 or x.shape[1] != c.shape[1]:
        raise ValueError(f'inconsistent shapes: latents {tuple(x.shape)}, centroids {tuple(c.shape)}')
    k, d = c.shape
    out = torch.empty((k, d, d), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _check(lib().rlvae_local_covariance(_ptr(x), x.shape[0], _ptr(c), k, d, c_float(temperature), _ptr(out),
                                            _stream(x)), 'rlvae_local_covariance')
    return out


def hmc_workspace(n: int, d: int, device) -> torch.Tensor:
    return torch.empty(max(int(lib().rlvae_hmc_workspace(n, d)), 1), device=device, dtype=torch.uint8)


def hmc_iteration(tab: Tables, z: torch.Tensor, gamma: torch.Tensor, acc: torch.Tensor, n_lf: int,
                  eps_lf: float, beta_zero_sqrt: float, scales, grad_mode: int = GRAD_MODULAR,
                  work: torch.Tensor | None = None, path: int = PATH_AUTO, want_stats: bool = False):
    """One MCMC iteration, in place on ``z``.  Returns (h0, h1, alpha, moves) if want_stats."""
    if not (z.is_cuda and z.is_contiguous() and z.dtype == torch.float32):
        raise RuntimeError('hmc_iteration: z must be a contiguous fp32 CUDA tensor (updated in place)')
    _req_z(tab, z)
    gamma = _req_z(tab, gamma, 'gamma')
    acc = _req(acc, 'acc')
    n, d = z.shape
    if gamma.shape[0] != n or tuple(acc.shape) != (n,) or acc.device != z.device:
        raise ValueError(f'hmc_iteration: gamma must be [{n}, {d}] and acc [{n}] on {z.device}')
    if len(scales) != n_lf:
        raise ValueError(f'hmc_iteration: need {n_lf} tempering scales, got {len(scales)}')
    if work is None:
        work = hmc_workspace(n, d, z.device)
    sc = (c_float * n_lf)(*[float(s) for s in scales])
    stats = [torch.empty(n, device=z.device) for _ in range(4)] if want_stats else [None] * 4
    with torch.cuda.device(z.device):
        _check(lib().rlvae_hmc_iteration(tab.handle, _ptr(z), _ptr(gamma), _ptr(acc), n, n_lf, c_float(eps_lf),
                                         c_float(beta_zero_sqrt), sc, grad_mode, _ptr(stats[0]), _ptr(stats[1]),
                                         _ptr(stats[2]), _ptr(stats[3]), _ptr(work), path, _stream(z)),
               'rlvae_hmc_iteration')
    return tuple(stats) if want_stats else None


def hmc_fused_available(tab: Tables, grad_mode: int = GRAD_MODULAR, path: int = PATH_AUTO) -> bool:
    """True when rlvae_hmc_iteration / rlvae_hmc_run take the single-launch trajectory kernel."""
    return bool(lib().rlvae_hmc_fused_available(tab.handle, grad_mode, path))


def hmc_fail_count(work: torch.Tensor, n: int, d: int) -> torch.Tensor:
    """View of the fused kernel's rounding-failure counter inside an HMC workspace (0-dim int32)."""
    off = int(lib().rlvae_hmc_workspace(n, d)) - 256
    return work[off:off + 4].view(torch.int32)[0]


def hmc_run(tab: Tables, z: torch.Tensor, gammas: torch.Tensor, accs: torch.Tensor, n_lf: int, eps_lf: float,
            beta_zero_sqrt: float, scales, grad_mode: int = GRAD_MODULAR, path: int = PATH_AUTO,
            want_stats: bool = False, want_trace: bool = False, work: torch.Tensor | None = None):
    """n_iters = gammas.shape[0] MCMC iterations in place on ``z`` -- ONE launch on the fused path.
    scales: n_iters * n_lf floats (host).  Returns dict(work, stats=(h0,h1,alpha,moves) each [I,n], trace [I,n,d])."""
    if not (z.is_cuda and z.is_contiguous() and z.dtype == torch.float32):
        raise RuntimeError('hmc_run: z must be a contiguous fp32 CUDA tensor (updated in place)')
    _req_z(tab, z)
    n, d = z.shape
    gammas = _req(gammas, 'gammas')
    accs = _req(accs, 'accs')
    iters = int(gammas.shape[0])
    if tuple(gammas.shape) != (iters, n, d) or tuple(accs.shape) != (iters, n) or gammas.device != z.device \
            or accs.device != z.device:
        raise ValueError(f'hmc_run: gammas must be [I, {n}, {d}] and accs [I, {n}] on {z.device}')
    if len(scales) != iters * n_lf:
        raise ValueError(f'hmc_run: need {iters * n_lf} tempering scales, got {len(scales)}')
    need = int(lib().rlvae_hmc_run_workspace(n, d, iters, n_lf))
    if work is None or work.numel() < need or work.device != z.device:
        work = torch.empty(max(need, 1), device=z.device, dtype=torch.uint8)
    sc = (c_float * max(len(scales), 1))(*[float(x) for x in scales])
    stats = [torch.empty((iters, n), device=z.device) for _ in range(4)] if want_stats else [None] * 4
    trace = torch.empty((iters, n, d), device=z.device) if want_trace else None
    with torch.cuda.device(z.device):
        _check(lib().rlvae_hmc_run(tab.handle, _ptr(z), _ptr(gammas), _ptr(accs), n, iters, n_lf, c_float(eps_lf),
                                   c_float(beta_zero_sqrt), sc, grad_mode, _ptr(stats[0]), _ptr(stats[1]),
                                   _ptr(stats[2]), _ptr(stats[3]), _ptr(trace), _ptr(work), path, _stream(z)),
               'rlvae_hmc_run')
    return dict(work=work, stats=tuple(stats) if want_stats else None, trace=trace)


def pythae_hmc_run(tab: Tables, z: torch.Tensor, gammas: torch.Tensor, accs: torch.Tensor, n_lf: int, eps_lf: float,
                   beta_zero_sqrt: float, scales, path: int = PATH_AUTO, want_stats: bool = False,
                   want_trace: bool = False, work: torch.Tensor | None = None):
    """pythae's RHVAESampler.hmc_sampling loop in place on ``z``: gammas.shape[0] MCMC iterations x n_lf leapfrog
    steps.  scales: n_iters * n_lf floats (host).  Returns dict(work, stats=(h0,h,alpha,moves) [I,n], trace [I,n,d])."""
    if not (z.is_cuda and z.is_contiguous() and z.dtype == torch.float32):
        raise RuntimeError('pythae_hmc_run: z must be a contiguous fp32 CUDA tensor (updated in place)')
    _req_z(tab, z)
    n, d = z.shape
    gammas = _req(gammas, 'gammas')
    accs = _req(accs, 'accs')
    iters = int(gammas.shape[0])
    if tuple(gammas.shape) != (iters, n, d) or tuple(accs.shape) != (iters, n) or gammas.device != z.device \
            or accs.device != z.device:
        raise ValueError(f'pythae_hmc_run: gammas must be [I, {n}, {d}] and accs [I, {n}] on {z.device}')
    if len(scales) != iters * n_lf:
        raise ValueError(f'pythae_hmc_run: need {iters * n_lf} tempering scales, got {len(scales)}')
    need = int(lib().rlvae_pythae_hmc_workspace(n, d))
    if work is None or work.numel() < need or work.device != z.device:
        work = torch.empty(max(need, 1), device=z.device, dtype=torch.uint8)
    sc = (c_float * max(len(scales), 1))(*[float(x) for x in scales])
    stats = [torch.empty((iters, n), device=z.device) for _ in range(4)] if want_stats else [None] * 4
    trace = torch.empty((iters, n, d), device=z.device) if want_trace else None
    with torch.cuda.device(z.device):
        _check(lib().rlvae_pythae_hmc_run(tab.handle, _ptr(z), _ptr(gammas), _ptr(accs), n, iters, n_lf,
                                          c_float(eps_lf), c_float(beta_zero_sqrt), sc, _ptr(stats[0]), _ptr(stats[1]),
                                          _ptr(stats[2]), _ptr(stats[3]), _ptr(trace), _ptr(work), path, _stream(z)),
               'rlvae_pythae_hmc_run')
    return dict(work=work, stats=tuple(stats) if want_stats else None, trace=trace)


def hmc_refine(tab: Tables, z: torch.Tensor, n_steps: int, step_size: float, path: int = PATH_AUTO):
    if not (z.is_cuda and z.is_contiguous() and z.dtype == torch.float32):
        raise RuntimeError('hmc_refine: z must be a contiguous fp32 CUDA tensor (updated in place)')
    _req_z(tab, z)
    n, d = z.shape
    work = hmc_workspace(n, d, z.device)
    with torch.cuda.device(z.device):
        _check(lib().rlvae_hmc_refine(tab.handle, _ptr(z), n, n_steps, c_float(step_size), _ptr(work), path,
                                      _stream(z)), 'rlvae_hmc_refine')
    return z


def nearest2(tab: Tables, mu: torch.Tensor):
    mu = _req_z(tab, mu, 'mu')
    n = mu.shape[0]
    idx = torch.empty((n, 2), device=mu.device, dtype=torch.int64)
    dist = torch.empty((n, 2), device=mu.device)
    with torch.cuda.device(mu.device):
        _check(lib().rlvae_nearest2(tab.handle, _ptr(mu), n, _ptr(idx), _ptr(dist), _stream(mu)), 'rlvae_nearest2')
    return idx, dist


def chol_apply(a: torch.Tensor, eps: torch.Tensor, jitter: float = 1e-6):
    a = _req(a, 'a')
    eps = _req(eps, 'eps')
    n, d = eps.shape
    if tuple(a.shape) != (n, d, d) or a.device != eps.device:
        raise ValueError(f'chol_apply: a must be [{n}, {d}, {d}] on {eps.device}, got {tuple(a.shape)} on {a.device}')
    out = torch.empty_like(eps)
    status = torch.empty(n, device=a.device, dtype=torch.int32)
    with torch.cuda.device(a.device):
        _check(lib().rlvae_chol_apply(_ptr(a), _ptr(eps), n, d, c_float(jitter), _ptr(out), _ptr(status),
                                      _stream(a)), 'rlvae_chol_apply')
    return out, status
