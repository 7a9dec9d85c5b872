"""CPU oracle for the RlVAE metric-evaluation + sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rlvae_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  It is a restatement of the
reference's *algorithm* (same tensor expressions, same materialised
intermediates, same fp32 op order) in plain CPU torch, so that

* parity tests have a checker that runs where ``/root/reference`` does not
  exist (the GPU box), and
* the CPU baseline times the same work the reference's eager code does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
modules from ``/root/reference`` (possible only in the build container), runs
them on seeded inputs and commits inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those
vectors (and, when ``/root/reference`` is present, against the live reference).
The reference's own tests pin almost nothing numerically: the
closed form of the identity-metric fixture (tests/test_modular_components.py:68-73)
is checked too.

Every function cites the reference lines it restates
(paths relative to the reference checkout).
"""
from __future__ import annotations

import torch


# --------------------------------------------------------------------------- #
# A2-A4, A20: src/models/components/metric_tensor.py
# --------------------------------------------------------------------------- #
def centroid_weights(z, centroids, temperature):
    """w[n,k] = exp(-||z_n - c_k||^2 / T^2).  metric_tensor.py:115-119
    (direct difference, square, sum over the last axis, divide by T**2)."""
    delta = centroids[None, :, :] - z[:, None, :]
    sq = (delta ** 2).sum(dim=-1)
    return torch.exp(-sq / (temperature ** 2))


def inverse_metric(z, centroids, matrices, temperature, regularization):
    """G^{-1}(z) = sum_k w_k M_k + lambda I.  metric_tensor.py:98-137.
    Materialises [N,K,d,d] exactly like the reference (lines 124-128)."""
    w = centroid_weights(z, centroids, temperature)
    stacked = w[:, :, None, None] * matrices[None]
    ginv = stacked.sum(dim=1)
    d = z.shape[1]
    eye = torch.eye(d, dtype=z.dtype, device=z.device)
    return ginv + (regularization * eye)[None].expand(z.shape[0], -1, -1)


def metric(z, centroids, matrices, temperature, regularization):
    """G = inv(G^{-1}).  metric_tensor.py:139-160 (LinAlgError retry with 1e-6 I)."""
    ginv = inverse_metric(z, centroids, matrices, temperature, regularization)
    try:
        return torch.linalg.inv(ginv)
    except torch.linalg.LinAlgError:
        d = z.shape[1]
        return torch.linalg.inv(ginv + 1e-6 * torch.eye(d, dtype=z.dtype)[None])


def log_det_metric(z, centroids, matrices, temperature, regularization):
    """log|det G(z)| = slogdet(inv(G^{-1})).logabsdet.  metric_tensor.py:162-182."""
    return torch.linalg.slogdet(
        metric(z, centroids, matrices, temperature, regularization)).logabsdet


def riemannian_distance_squared(z1, z2, centroids, matrices, temperature, regularization):
    """(z1-z2)^T G((z1+z2)/2) (z1-z2).  metric_tensor.py:184-207."""
    mid = 0.5 * (z1 + z2)
    g = metric(mid, centroids, matrices, temperature, regularization)
    dz = z1 - z2
    return torch.einsum('bi,bij,bj->b', dz, g, dz)


# --------------------------------------------------------------------------- #
# A5-A10: src/models/samplers/hmc_sampler.py
# --------------------------------------------------------------------------- #
def hmc_log_pi(z, centroids, matrices, temperature, regularization):
    """0.5*log(clamp(det G^{-1}, 1e-10)).  hmc_sampler.py:26-30 (det, not slogdet)."""
    ginv = inverse_metric(z, centroids, matrices, temperature, regularization)
    det = torch.linalg.det(ginv).clamp(min=1e-10)
    return 0.5 * torch.log(det)


def hmc_grad_modular(z, centroids, matrices, temperature, regularization):
    """Variant A, the 'gradient' RiemannianHMCSampler integrates.
    hmc_sampler.py:33-68: diag(-0.5 * G^T @ ((-2/T^2) sum_k w_k M_k)^T).
    Note it re-derives the weights with torch.norm(...)**2 (line 49) rather
    than the sum of squares used inside G."""
    g = metric(z, centroids, matrices, temperature, regularization)
    delta = centroids[None] - z[:, None]
    sq = torch.norm(delta, dim=-1) ** 2
    w = torch.exp(-sq / (temperature ** 2))
    summed = (matrices[None] * w[:, :, None, None]).sum(dim=1)
    term = (-2 / (temperature ** 2)) * summed
    out = -0.5 * g.transpose(-2, -1) @ term.transpose(-2, -1)
    return out.diagonal(dim1=-2, dim2=-1)


def hmc_grad_modular_closed_form(z, centroids, matrices, temperature, regularization):
    """Closed form of variant A: (1 - lambda * G_ii) / T^2.
    This is the expression the CUDA path evaluates; kept here so the tests can
    show it equals the literal restatement above to rounding."""
    g = metric(z, centroids, matrices, temperature, regularization)
    return (1.0 - regularization * g.diagonal(dim1=-2, dim2=-1)) / (temperature ** 2)


def grad_log_sqrt_det_ginv_exact(z, centroids, matrices, temperature, regularization):
    """Variant D: grad_z 0.5*log det G^{-1} = (1/T^2) sum_k w_k tr(G M_k) (c_k - z).
    Equals autograd through metric_tensor.py:98-137 + hmc_sampler.py:26-30
    (the path hmc_sampler.py:183-204 differentiates).  grad_z log det G = -2x this."""
    w = centroid_weights(z, centroids, temperature)
    g = metric(z, centroids, matrices, temperature, regularization)
    tr = torch.einsum('nij,kij->nk', g, matrices)      # <G_n, M_k>  (tr(G M^T))
    u = w * tr
    diff = centroids[None] - z[:, None]
    return (u[:, :, None] * diff).sum(1) / (temperature ** 2)


def grad_log_det_metric_autograd(z, centroids, matrices, temperature, regularization):
    """north_star's 'grad_z log det G' through autograd on the reference expressions."""
    zz = z.clone().detach().requires_grad_(True)
    ld = log_det_metric(zz, centroids, matrices, temperature, regularization)
    return torch.autograd.grad(ld.sum(), zz)[0]


def grad_pythae(z, centroids, matrices, temperature, regularization):
    """Variant C (pythae RHVAESampler), src/lib/src/pythae/samplers/manifold_sampler/
    rhvae_sampler.py:160-187:  (1/T^2) * G^T @ sum_k w_k M_k^T (c_k - z)  -> [N,d]."""
    g = metric(z, centroids, matrices, temperature, regularization)
    delta = centroids[None] - z[:, None]
    w = torch.exp(-(torch.norm(delta, dim=-1) ** 2) / (temperature ** 2))
    # [N,K,1,d] @ [N,K,d,d] -> (c-z)^T M_k  i.e. M_k^T (c-z)
    v = (delta[:, :, None, :] @ (matrices[None] * w[:, :, None, None])).sum(1)  # [N,1,d]
    out = g.transpose(-1, -2) @ v.transpose(-1, -2) / (temperature ** 2)
    return out.squeeze(-1)


def metric_backward(z, centroids, matrices, temperature, grad_ginv):
    """Backward of A2 for an arbitrary upstream gradient U = dL/dG^{-1}:
    dL/dz = (2/T^2) sum_k w_k <U, M_k> (c_k - z)   (SURVEY.md §2.1 row K4)."""
    w = centroid_weights(z, centroids, temperature)
    ip = torch.einsum('nij,kij->nk', grad_ginv, matrices)
    diff = centroids[None] - z[:, None]
    return (2.0 / temperature ** 2) * ((w * ip)[:, :, None] * diff).sum(1)


def tempering(k, n_steps, beta_zero_sqrt):
    """1/beta_k, beta_k=(1-1/b0)(k/K)^2+1/b0.  hmc_sampler.py:98-102."""
    beta_k = ((1 - 1 / beta_zero_sqrt) * (k / n_steps) ** 2) + 1 / beta_zero_sqrt
    return 1 / beta_k


def hmc_sample(tables, z0, gammas, accs, n_lf, eps_lf, beta_zero=1.0,
               z_forced=None, record=None):
    """RiemannianHMCSampler.sample with the RNG draws injected.
    hmc_sampler.py:104-165.  ``z0`` replaces the randn of line 114, ``gammas[i]``
    the randn_like of line 122 and ``accs[i]`` the rand of line 158.

    ``z_forced[i]`` (optional) overrides the chain state at the start of MCMC
    iteration i (teacher forcing).  ``record`` (optional dict)
    receives per-iteration lists H0, H, alpha, moves, z.
    The tempering state ``beta_sqrt_old`` is deliberately NOT reset per MCMC
    iteration (line 116 is outside the loop)."""
    c, M, T, lam = tables
    eps = torch.tensor([eps_lf], dtype=z0.dtype)
    b0 = torch.tensor([beta_zero], dtype=z0.dtype).sqrt()
    beta_old = b0
    z_prev = z0.clone()
    z = z0.clone()
    for i in range(gammas.shape[0]):
        if z_forced is not None:
            z_prev = z_forced[i].clone()
            z = z_prev.clone()
        rho = gammas[i] / b0
        H0 = -hmc_log_pi(z, c, M, T, lam) + 0.5 * torch.norm(rho, dim=1) ** 2
        for k in range(n_lf):
            g = -hmc_grad_modular(z, c, M, T, lam)
            rho_half = rho - (eps / 2) * g
            z = z + eps * rho_half
            g = -hmc_grad_modular(z, c, M, T, lam)
            rho_full = rho_half - (eps / 2) * g
            beta_new = tempering(k + 1, n_lf, b0)
            rho = (beta_old / beta_new) * rho_full
            beta_old = beta_new
        H = -hmc_log_pi(z, c, M, T, lam) + 0.5 * torch.norm(rho, dim=1) ** 2
        alpha = torch.exp(-H) / (torch.exp(-H0) + 1e-10)
        alpha = torch.clamp(alpha, 0, 1)
        moves = (accs[i] < alpha).to(z.dtype).reshape(-1, 1)
        z = moves * z + (1 - moves) * z_prev
        z_prev = z.clone()
        if record is not None:
            for name, val in (('H0', H0), ('H', H), ('alpha', alpha),
                              ('moves', moves.reshape(-1)), ('z', z.clone())):
                record.setdefault(name, []).append(val)
    return z


def hmc_refine(mu, log_var, eps, tables, n_steps=3, step_size=0.01):
    """sample_riemannian_latents(method='hmc'): z = mu + eps*sigma, then three
    steps z <- z + 0.01 * (-grad_func(z)).  hmc_sampler.py:236-257."""
    c, M, T, lam = tables
    z = mu + eps * torch.exp(0.5 * log_var)
    for _ in range(n_steps):
        z = z + step_size * (-hmc_grad_modular(z, c, M, T, lam))
    return z


def hmc_sample_posterior(mu, log_var, eps0, gammas, tables, n_iter=20, n_lf=5, step=0.01):
    """sample_posterior, hmc_sampler.py:167-214: energy = -log_pi + Gaussian term,
    gradient by autograd, no accept/reject; note the position update sign
    ``z - 0.01*rho`` (line 210)."""
    c, M, T, lam = tables

    def grad_energy(zz):
        zz = zz.clone().detach().requires_grad_(True)
        e = -hmc_log_pi(zz, c, M, T, lam) + 0.5 * torch.sum(
            (zz - mu) * torch.exp(-log_var) * (zz - mu), dim=1)
        return torch.autograd.grad(e.sum(), zz)[0]

    z = (mu + eps0 * torch.exp(0.5 * log_var)).detach()
    for i in range(n_iter):
        rho = gammas[i] * 0.1
        for _ in range(n_lf):
            rho = rho - (step / 2) * grad_energy(z)
            z = (z - step * rho).detach()
            rho = rho - (step / 2) * grad_energy(z)
    return z


# --------------------------------------------------------------------------- #
# A14-A17: src/models/samplers/riemannian_sampler.py
# --------------------------------------------------------------------------- #
def nearest2(mu, centroids):
    """Euclidean distances to all centroids and the two smallest.
    riemannian_sampler.py:58-67 / 125-131.  Returns (idx [N,2], dist [N,2])."""
    dist = torch.norm(mu[:, None] - centroids[None], dim=-1)
    _, idx = torch.topk(dist, k=2, dim=-1, largest=False)
    return idx, torch.gather(dist, 1, idx)


def chol_apply(a, eps, jitter=1e-6):
    """L @ eps with L = cholesky(A + 1e-6 I).  riemannian_sampler.py:83-84."""
    d = a.shape[-1]
    L = torch.linalg.cholesky(a + jitter * torch.eye(d, dtype=a.dtype))
    return torch.einsum('bij,bj->bi', L, eps)


def sample_enhanced(mu, log_var, eps, tables):
    """riemannian_sampler.py:41-103 with eps injected (main path, no fallbacks)."""
    c, M, T, lam = tables
    idx, d2 = nearest2(mu, c)
    wts = 1.0 / (d2 + 1e-8)
    wts = wts / wts.sum(dim=-1, keepdim=True)
    virt = wts[:, 0:1] * c[idx[:, 0]] + wts[:, 1:2] * c[idx[:, 1]]
    ginv = inverse_metric(virt, c, M, T, lam)
    et = chol_apply(ginv, eps)
    sig = torch.exp(0.5 * log_var)
    return mu + et * sig * 0.15 + eps * sig * (1.0 - 0.15)


def sample_geodesic(mu, log_var, eps, t_geo, tables):
    """riemannian_sampler.py:105-181 with eps (line 116) and t (line 138) injected."""
    c, M, T, lam = tables
    idx, _ = nearest2(mu, c)
    c1, c2 = c[idx[:, 0]], c[idx[:, 1]]
    zg = (1 - t_geo) * c1 + t_geo * c2
    direction = c2 - c1
    direction = direction / (torch.norm(direction, dim=-1, keepdim=True) + 1e-8)
    off = mu - zg
    par = torch.sum(off * direction, dim=-1, keepdim=True) * direction
    g = torch.linalg.inv(inverse_metric(zg, c, M, T, lam))
    ep = chol_apply(g, eps)
    return zg + 0.3 * ep * torch.exp(0.5 * log_var) + (1.0 - 0.3) * (mu - zg) + 0.1 * par


def sample_basic(mu, log_var, eps, tables):
    """riemannian_sampler.py:183-220."""
    c, M, T, lam = tables
    sig = torch.exp(0.5 * log_var)
    zs = mu + eps * sig
    et = chol_apply(inverse_metric(zs, c, M, T, lam), eps)
    return mu + et * sig * 0.1 + eps * sig * (1.0 - 0.1)


def sample_geodesic_prior(idx1, idx2, t, eps, tables):
    """riemannian_sampler.py:242-288 with randint/rand/randn injected."""
    c, M, T, lam = tables
    zg = (1 - t) * c[idx1] + t * c[idx2]
    et = chol_apply(inverse_metric(zg, c, M, T, lam), eps)
    return zg + 0.1 * et


# --------------------------------------------------------------------------- #
# chunked drivers (the reference cannot hold [N,K,d,d] for large N: BASELINE.md §2)
# --------------------------------------------------------------------------- #
def chunked(fn, z, *args, chunk=512):
    outs = [fn(z[s:s + chunk], *args) for s in range(0, z.shape[0], chunk)]
    return torch.cat(outs, 0)


def eval_ginv_logdet_grad(z, centroids, matrices, temperature, regularization):
    """One 'metric eval' of BASELINE.json's metric as the reference would do it:
    G^{-1} (A2), log det G (A4, re-evaluates G^{-1}), and grad_z log det G by
    autograd."""
    ginv = inverse_metric(z, centroids, matrices, temperature, regularization)
    zz = z.clone().detach().requires_grad_(True)
    ld = log_det_metric(zz, centroids, matrices, temperature, regularization)
    grad = torch.autograd.grad(ld.sum(), zz)[0]
    return ginv, ld.detach(), grad


# ---------------------------------------------------------------------------------------------
# Metric construction
def build_local_metrics(all_mus, centroids, temperature, regularization):
    """ref scripts/train_and_extract_vanilla_vae.py:204-226: per centroid, normalised Gaussian
    weights over all encoded latents, weighted mean, weighted covariance + reg I, then the
    minimum-eigenvalue lift to 1e-6.  Same op order as the reference loop."""
    d = all_mus.shape[1]
    out = []
    for c in centroids:
        dists = torch.norm(all_mus - c, dim=1)
        weights = torch.exp(-dists ** 2 / (temperature ** 2))
        weights = weights / (weights.sum() + 1e-8)
        mean = (weights.unsqueeze(1) * all_mus).sum(dim=0)
        diffs = all_mus - mean.unsqueeze(0)
        metric = torch.einsum('n,ni,nj->ij', weights, diffs, diffs) + regularization * torch.eye(d, dtype=all_mus.dtype)
        min_eig = torch.linalg.eigvals(metric).real.min().item()
        if min_eig < 1e-6:
            metric = metric + (1e-6 - min_eig) * torch.eye(d, dtype=all_mus.dtype)
        out.append(metric)
    return torch.stack(out, dim=0)


# ---------------------------------------------------------------------------------------------
# A8: pythae RHVAESampler.hmc_sampling (variant C), what OfficialRHVAESampler.sample_prior runs
def rhvae_log_pi(z, c, M, T, lam):
    """rhvae_sampler.py:157-158: log(sqrt(det G^{-1}) + 1e-10)."""
    return torch.log(torch.sqrt(torch.det(inverse_metric(z, c, M, T, lam))) + 1e-10)


def rhvae_hmc_sample(tables, idx0, gammas, accs, n_lf, eps_lf, beta_zero=1.0, record=None):
    """src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py:98-148 with the RNG draws
    injected: ``idx0`` replaces the randint of line 100 (z0 = centroids[idx0]), ``gammas[i]`` the
    randn_like of line 107, ``accs[i]`` the rand of line 141.  alpha is NOT clamped and has no
    epsilon (line 140); the tempering state is not reset between MCMC iterations (line 104)."""
    c, M, T, lam = tables
    b0 = float(beta_zero) ** 0.5
    beta_old = b0
    z0 = c[idx0]
    z = z0
    n, d = z.shape
    for i in range(gammas.shape[0]):
        rho = gammas[i] / b0
        H0 = -rhvae_log_pi(z, c, M, T, lam) + 0.5 * torch.norm(rho, dim=1) ** 2
        for k in range(n_lf):
            g = -grad_pythae(z, c, M, T, lam).reshape(n, d)
            rho_ = rho - (eps_lf / 2) * g
            z = z + eps_lf * rho_
            g = -grad_pythae(z, c, M, T, lam).reshape(n, d)
            rho__ = rho_ - (eps_lf / 2) * g
            beta_new = 1.0 / (((1 - 1 / b0) * ((k + 1) / n_lf) ** 2) + 1 / b0)
            rho = (beta_old / beta_new) * rho__
            beta_old = beta_new
        H = -rhvae_log_pi(z, c, M, T, lam) + 0.5 * torch.norm(rho, dim=1) ** 2
        alpha = torch.exp(-H) / torch.exp(-H0)
        moves = (accs[i] < alpha).to(torch.int).reshape(n, 1)
        z = z * moves + (1 - moves) * z0
        z0 = z
        if record is not None:
            for name, val in (('H0', H0), ('H', H), ('alpha', alpha), ('moves', moves.reshape(-1)), ('z', z.clone())):
                record.setdefault(name, []).append(val)
    return z
