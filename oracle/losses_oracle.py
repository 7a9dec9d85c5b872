"""CPU restatement of the reference's metric-dependent LOSSES.

TEST INFRASTRUCTURE ONLY (see oracle/metric_oracle.py).  Each function takes the metric as a
duck-typed object / callable exactly like the reference does, so that the tests can run the SAME
arithmetic once with the CPU oracle metric (pinned against goldens produced by the real reference,
oracle/make_golden_losses.py) and once with the CUDA drop-in ``rlvae_b200.MetricTensor`` -- loss value
and the gradients w.r.t. mu and log_var must agree.
"""
from __future__ import annotations

import torch

from . import metric_oracle as O


class OracleMetric:
    """The three MetricTensor methods the losses call, on the CPU oracle."""

    def __init__(self, centroids, matrices, temperature, regularization):
        self.t = (centroids, matrices, temperature, regularization)

    def to(self, device):            # loss_manager.py:106-107 calls .to(target_device) on it
        return self

    def compute_inverse_metric(self, z):
        return O.inverse_metric(z, *self.t)

    def compute_metric(self, z):
        return O.metric(z, *self.t)

    def compute_log_det_metric(self, z):
        return O.log_det_metric(z, *self.t)


def modular_riemannian_kl(mu, log_var, z_samples, metric_tensor):
    """LossManager.compute_riemannian_kl_loss, ref src/models/components/loss_manager.py:75-146 (main
    path).  Reference quirk kept: ``mu_diff.unsqueeze(1) * (G_inv mu)`` broadcasts [B,1,d] x [B,d] to
    [B,B,d], so term2 is a [B,B] matrix (row i: mu_i . G_inv_j mu_j) and the mean runs over B^2 entries."""
    g_inv_mu = metric_tensor.compute_inverse_metric(mu)                      # :110
    log_det = metric_tensor.compute_log_det_metric(mu)                       # :113
    batch, latent_dim = mu.shape
    sigma_post = torch.diag_embed(torch.exp(log_var))                        # :124
    term1 = torch.sum(torch.diagonal(torch.bmm(g_inv_mu, sigma_post), dim1=-2, dim2=-1), dim=-1)   # :127
    mu_diff = mu - torch.zeros_like(mu)                                      # :130
    term2 = torch.sum(mu_diff.unsqueeze(1) * torch.bmm(g_inv_mu, mu_diff.unsqueeze(-1)).squeeze(-1), dim=-1)  # :131
    term3 = log_det                                                          # :134
    term4 = -latent_dim                                                      # :137
    return torch.mean(0.5 * (term1 + term2 + term3 + term4))                 # :140-142


def monolith_metric_kl(mu, log_var, z_samples, G):
    """RiemannianFlowVAE.compute_riemannian_metric_kl_loss, ref src/models/riemannian_flow_vae.py:1004-1077:
    0.5 * mean_b (z-mu)^T G(z) (z-mu)   (``G`` = the model's callable z -> [B,d,d])."""
    g_z = G(z_samples)                                                       # :1030
    diff = z_samples - mu                                                    # :1049
    quad = torch.bmm(torch.bmm(diff.unsqueeze(1), g_z), diff.unsqueeze(-1)).squeeze(-1).squeeze(-1)   # :1054-1057
    return 0.5 * quad.mean()                                                 # :1069


def monolith_riemannian_kl(mu, log_var, z_sample, G):
    """RiemannianFlowVAE.compute_riemannian_kl_loss, ref src/models/riemannian_flow_vae.py:1328-1394 (main
    path).  ``quadratic_term`` is [B,1] after the single squeeze (:1360), so the sum broadcasts to [B,B]
    like the reference; det (not slogdet) clamped to [1e-10, 1e10] (:1365-1367)."""
    log_var_clamped = torch.clamp(log_var, -10.0, 10.0)                      # :1349
    g_z = G(z_sample)                                                        # :1352
    trace_term = torch.sum(torch.diagonal(g_z, dim1=-2, dim2=-1) * torch.exp(log_var), dim=1)       # :1360
    quad = torch.bmm(mu.unsqueeze(1), torch.bmm(g_z, mu.unsqueeze(-1))).squeeze(-1)                  # :1364
    det_g = torch.clamp(torch.linalg.det(g_z), min=1e-10, max=1e10)          # :1369-1370
    log_det_prior = torch.log(det_g)
    log_det_post = torch.sum(log_var_clamped, dim=1)                         # :1372
    kl = 0.5 * (trace_term + quad - mu.shape[1] + log_det_prior - log_det_post)   # :1378
    return kl.mean()
