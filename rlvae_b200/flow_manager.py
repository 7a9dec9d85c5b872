"""``FlowManager`` (ref src/models/components/flow_manager.py:13-116) plus the IAF/MADE blocks it
needs, in stock PyTorch.

The reference builds its flows from the vendored pythae package
(src/lib/src/pythae/models/normalizing_flows/{iaf,made,layers}), which is not installed where
this framework runs, so the three small modules are re-implemented here with the SAME parameter
and buffer names -- a reference ``FlowManager.state_dict()`` loads unchanged -- and the same
arithmetic (tests/golden/flow_d16.npz).  The flows are small masked MLPs and stay on stock
PyTorch (out of scope for custom kernels).  What this file adds for
the hot path is ``metric_along_flow``: the per-timestep metric evaluation that the reference's
consumers run in a Python loop (src/visualizations/flow_analysis.py:104-126), as ONE fused CUDA
evaluation over the flattened [B*T, d] trajectory.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class MaskedLinear(nn.Linear):
    """nn.Linear whose weight is multiplied by a fixed 0/1 mask (pythae layers.py:15-33)."""

    def __init__(self, in_features: int, out_features: int, mask: torch.Tensor):
        super().__init__(in_features, out_features)
        self.register_buffer('mask', mask)

    def forward(self, x):
        return F.linear(x, self.mask * self.weight, self.bias)


def _sequential_degrees(dim: int, hidden: List[int]):
    """MADE 'sequential' degree assignment (pythae made_model.py:84-98)."""
    deg = {-1: torch.arange(1, dim + 1)}
    for i, h in enumerate(hidden):
        floor = min(int(torch.min(deg[i - 1])), dim - 1)
        d = np.ceil(np.arange(1, h + 1) * (dim - 1) / float(h + 1)).astype(np.int32)
        deg[i] = torch.from_numpy(np.maximum(floor, d))
    return deg


class MADE(nn.Module):
    """Masked autoencoder producing (mu, log_var) with log_var clamped to +-1.5.
    Layer layout and names follow pythae made_model.py:24-139: ``context_input_layer`` feeds
    ``net`` = [MaskedLinear, ReLU] * (len(hidden)-1) + [MaskedLinear -> 2*dim] directly (there is
    no activation between ``context_input_layer`` and ``net[0]``, as in the reference)."""

    def __init__(self, dim: int, hidden_sizes: List[int]):
        super().__init__()
        self.input_dim = dim
        deg = _sequential_degrees(dim, hidden_sizes)
        masks = [(deg[i].unsqueeze(-1) >= deg[i - 1].unsqueeze(0)).float() for i in range(len(hidden_sizes))]
        masks.append((deg[len(hidden_sizes) - 1].unsqueeze(0) < deg[-1].unsqueeze(-1)).float())
        sizes = list(hidden_sizes) + [dim]
        self.context_input_layer = MaskedLinear(dim, sizes[0], masks[0])
        layers: List[nn.Module] = []
        for n_in, n_out, m in zip(sizes[:-1], sizes[1:-1], masks[1:-1]):
            layers += [MaskedLinear(n_in, n_out, m), nn.ReLU()]
        layers.append(MaskedLinear(hidden_sizes[-1], 2 * dim, masks[-1].repeat(2, 1)))
        self.net = nn.Sequential(*layers)
        with torch.no_grad():
            self.net[-1].bias[dim:].fill_(-2.0)

    def forward(self, x):
        o = self.net(self.context_input_layer(x.reshape(x.shape[0], -1)))
        return o[:, :self.input_dim], torch.clamp(o[:, self.input_dim:], -1.5, 1.5)


class IAF(nn.Module):
    """Inverse autoregressive flow: ``n_blocks`` MADE layers, each applied with ``dim``
    sequential passes, followed by a feature flip (pythae iaf_model.py:50-83)."""

    def __init__(self, dim: int, hidden_size: int = 128, n_blocks: int = 2, n_hidden_in_made: int = 3):
        super().__init__()
        self.input_dim = dim
        self.net = nn.ModuleList([MADE(dim, [hidden_size] * n_hidden_in_made) for _ in range(n_blocks)])

    def forward(self, x):
        x = x.reshape(x.shape[0], -1)
        log_det = torch.zeros(x.shape[0], device=x.device)
        for made in self.net:
            y = torch.zeros_like(x)
            for i in range(self.input_dim):
                mu, s = made(y.clone())
                y[:, i] = (x[:, i] - mu[:, i]) * (-s[:, i]).exp()
                log_det = log_det - s[:, i]
            x = y.flip(dims=(1,))
        return SimpleNamespace(out=x, log_abs_det_jac=log_det)


class FlowManager(nn.Module):
    def __init__(self, latent_dim: int, n_flows: int = 8, flow_hidden_size: int = 256, flow_n_blocks: int = 2,
                 flow_n_hidden: int = 1, device: Optional[torch.device] = None):
        super().__init__()
        self.latent_dim = latent_dim
        self.n_flows = n_flows
        self.flow_hidden_size = flow_hidden_size
        self.flow_n_blocks = flow_n_blocks
        self.flow_n_hidden = flow_n_hidden
        self.device = device or torch.device('cpu')
        # The reference passes ``n_hidden=flow_n_hidden`` to IAFConfig, a field that does not
        # exist, so pydantic drops it and every MADE keeps the default depth of 3
        # (flow_manager.py:28 vs iaf_config.py:24).  Reproduced.
        self.flows = nn.ModuleList([IAF(latent_dim, flow_hidden_size, flow_n_blocks, 3) for _ in range(n_flows)])
        self.to(self.device)

    def apply_flows(self, z_seq: list, n_obs: int = None):
        """z_t = IAF_{t-1}(z_{t-1}); returns (latents, log|det J| per step).  With a single start
        latent and ``n_obs`` the sequence is rolled out, re-using the last flow beyond
        ``n_flows`` (flow_manager.py:45-56); otherwise one flow per provided step (:58-68)."""
        steps = n_obs if (n_obs is not None and len(z_seq) == 1) else len(z_seq)
        rollout = n_obs is not None and len(z_seq) == 1
        out, log_dets = [z_seq[0]], []
        for t in range(1, steps):
            flow = self.flows[t - 1] if (not rollout or t - 1 < len(self.flows)) else self.flows[-1]
            res = flow(out[-1])
            out.append(res.out)
            log_dets.append(res.log_abs_det_jac)
        return out, log_dets

    def invert_flows(self, z_seq: List[torch.Tensor]) -> List[torch.Tensor]:
        raise NotImplementedError('Invert flows is not implemented for IAF.')

    def get_log_det_jacobians(self, z_seq: List[torch.Tensor]) -> List[torch.Tensor]:
        return self.apply_flows(z_seq)[1]

    def get_flow_params(self) -> Dict[str, Any]:
        return {'latent_dim': self.latent_dim, 'n_flows': self.n_flows, 'flow_hidden_size': self.flow_hidden_size,
                'flow_n_blocks': self.flow_n_blocks, 'flow_n_hidden': self.flow_n_hidden}

    def diagnose_flows(self) -> Dict[str, Any]:
        return {'total_params': sum(p.numel() for p in self.parameters()), 'n_flows': self.n_flows}

    # ------------------------------------------------------------------ hot-path consumer (A19)
    @torch.no_grad()
    def metric_along_flow(self, metric_tensor, z0: torch.Tensor, n_obs: int, want_g: bool = False,
                          want_spectrum: bool = False):
        """Roll ``z0`` [B,d] through the flows and evaluate the metric at every step in one
        fused call: returns dict(z [B,T,d], log_det_jacobians [B,T-1], logdet_G [B,T],
        det_G [B,T], optionally G [B,T,d,d]) -- what flow_analysis.py:104-126 computes per t."""
        z_seq, lds = self.apply_flows([z0], n_obs=n_obs)
        z = torch.stack(z_seq, dim=1)
        b, t, d = z.shape
        ev = metric_tensor.evaluate(z.reshape(b * t, d).contiguous(), want_ginv=False, want_g=want_g,
                                    want_logdet=True, want_grad=False)
        ld = ev['logdet_g'].reshape(b, t)
        out = {'z': z, 'log_det_jacobians': torch.stack(lds, dim=1) if lds else z.new_zeros(b, 0),
               'logdet_G': ld, 'det_G': torch.exp(ld)}
        if want_g:
            out['G'] = ev['g'].reshape(b, t, d, d)
        if want_spectrum:   # eigenvalues / condition number / traces per (B, T) (manifold.py:86-93)
            sp = metric_tensor.compute_metric_spectrum(z.reshape(b * t, d).contiguous())
            for k in ('eigenvals_G_inv', 'eigenvals_G'):
                out[k] = sp[k].reshape(b, t, d)
            for k in ('condition_number', 'trace_G', 'trace_G_inv'):
                out[k] = sp[k].reshape(b, t)
        return out
