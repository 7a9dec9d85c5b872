"""Metric construction on the GPU: the step that produces the tables the hot path consumes.

Host-side mirror of ref ``scripts/train_and_extract_vanilla_vae.py:187-252`` (local weighted
covariance around every centroid -> ``M_matrices``; the result dictionary is what
``MetricLoader.load_from_file`` / ``MetricTensor.load_pretrained`` accept) and of the RHVAE
accumulation ``M = L L^T`` (ref ``src/lib/src/pythae/models/rhvae/rhvae_model.py:172-178,381-385``).

The reference loops over centroids in Python with O(K) passes over all N latents; here one CUDA
launch (``rlvae_local_covariance``) streams the latents once per centroid CTA, and the
minimum-eigenvalue lift uses the batched eigenvalue kernel.  Centroid *selection* -- the reference
standardises the latents and runs ``sklearn_extra.cluster.KMedoids(n_clusters, random_state=42,
max_iter=1000, init='k-medoids++')`` (ref :187-197; method 'alternate') -- is ``select_centroids_kmedoids``:
the same algorithm (k-medoids++ seeding by D^2 sampling, then alternate assignment / medoid update until the
medoids stop moving) as batched device tensor operations.  PARITY UNPINNED for that one function:
sklearn_extra is not installed here and its seeding consumes a numpy RandomState stream, so the selected
indices cannot be compared with the reference's; the tests check the algorithm's invariants instead.
There is no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import _capi


def build_local_metrics(all_mus: torch.Tensor, centroids: torch.Tensor, temperature: float = 0.1,
                        regularization: float = 0.01, min_eigenvalue: float = 1e-6) -> torch.Tensor:
    """M [K,d,d] = weighted covariance of ``all_mus`` [N,d] around every centroid + reg I, lifted so
    that its smallest eigenvalue is at least ``min_eigenvalue`` (ref lines 204-226)."""
    if not (all_mus.is_cuda and centroids.is_cuda):
        raise RuntimeError('rlvae_b200.metric_builder runs on CUDA only (no CPU fallback)')
    d = all_mus.shape[1]
    cov = _capi.local_covariance(all_mus.float(), centroids.float(), float(temperature))
    eye = torch.eye(d, device=cov.device, dtype=cov.dtype)
    m = cov + regularization * eye
    m = 0.5 * (m + m.transpose(1, 2)) if d > 16 else m      # the d <= 16 kernel writes exactly symmetric output
    min_eig = _capi.sym_eigvalsh(m)[:, 0]
    lift = torch.clamp(min_eigenvalue - min_eig, min=0.0)
    lift = torch.where(min_eig < min_eigenvalue, lift, torch.zeros_like(lift))
    return m + lift[:, None, None] * eye


def select_centroids_kmedoids(all_mus: torch.Tensor, n_centroids: int = 50, max_iter: int = 1000, seed: int = 42,
                              standardize: bool = True, return_info: bool = False):
    """Indices [n_centroids] of the medoids of ``all_mus`` [N,d] (ref :187-197: StandardScaler, then k-medoids with
    k-medoids++ seeding, 'alternate' updates, Euclidean distance).  Everything runs on the latents' device; the
    [N,N] distance matrix is formed once (N is the number of encoded training latents: thousands)."""
    if not all_mus.is_cuda:
        raise RuntimeError('rlvae_b200.metric_builder runs on CUDA only (no CPU fallback)')
    x = all_mus.float()
    n = x.shape[0]
    if not 1 <= n_centroids <= n:
        raise ValueError(f'n_centroids must be in [1, {n}]')
    if standardize:                                   # sklearn StandardScaler: population std, zero std -> 1
        std = x.std(dim=0, unbiased=False)
        x = (x - x.mean(dim=0)) / torch.where(std > 0, std, torch.ones_like(std))
    dist = torch.cdist(x, x)                          # [N,N]
    gen = torch.Generator(device=x.device).manual_seed(seed)
    # ---- k-medoids++ seeding: first medoid uniformly, then proportional to the squared distance to the chosen set
    med = torch.empty(n_centroids, dtype=torch.long, device=x.device)
    med[0] = torch.randint(0, n, (1,), generator=gen, device=x.device)
    closest = dist[med[0]] ** 2
    for i in range(1, n_centroids):
        p = closest / closest.sum().clamp_min(1e-30)
        med[i] = torch.multinomial(p, 1, generator=gen) if float(closest.sum()) > 0 else torch.randint(
            0, n, (1,), generator=gen, device=x.device)
        closest = torch.minimum(closest, dist[med[i]] ** 2)
    # ---- alternate: assign to the nearest medoid, move every medoid to the member minimising the in-cluster
    # distance sum, until nothing moves
    costs = []
    it = 0
    for it in range(max_iter):
        d_med = dist[:, med]                          # [N,K]
        label = d_med.argmin(dim=1)
        costs.append(float(d_med.gather(1, label[:, None]).sum()))
        member = torch.nn.functional.one_hot(label, n_centroids).to(dist.dtype)      # [N,K]
        in_cluster = dist @ member                    # [N,K]: sum of distances from point i to the members of cluster k
        in_cluster = torch.where(member.bool(), in_cluster, torch.full_like(in_cluster, float('inf')))
        new_med = in_cluster.argmin(dim=0)            # best member per cluster
        empty = member.sum(dim=0) == 0
        new_med = torch.where(empty, med, new_med)
        # keep the current medoid unless the candidate is strictly better (sklearn_extra does the same)
        cur = in_cluster[med, torch.arange(n_centroids, device=x.device)]
        best = in_cluster[new_med, torch.arange(n_centroids, device=x.device)]
        new_med = torch.where(best < cur, new_med, med)
        if torch.equal(new_med, med):
            break
        med = new_med
    if return_info:
        d_med = dist[:, med]
        label = d_med.argmin(dim=1)
        return med, {'labels': label, 'inertia': float(d_med.gather(1, label[:, None]).sum()), 'n_iter': it + 1,
                     'cost_history': costs}
    return med


def build_metric_data(all_mus: torch.Tensor, centroids: Optional[torch.Tensor] = None,
                      centroid_indices: Optional[torch.Tensor] = None, temperature: float = 0.1,
                      regularization: float = 0.01) -> Dict[str, Any]:
    """The dictionary the reference saves to ``metric.pt`` (ref lines 241-248)."""
    if centroids is None:
        if centroid_indices is None:
            raise ValueError('pass centroids or centroid_indices (e.g. select_centroids_kmedoids(all_mus, n))')
        centroids = all_mus[centroid_indices]
    m = build_local_metrics(all_mus, centroids, temperature, regularization)
    return {'centroids': centroids, 'M_matrices': m, 'temperature': torch.tensor(temperature),
            'regularization': torch.tensor(regularization), 'latent_dim': int(all_mus.shape[1]),
            'n_centroids': int(centroids.shape[0])}


def metric_statistics(m: torch.Tensor) -> Dict[str, float]:
    """The statistics the reference prints after construction (ref lines 228-239)."""
    ev = _capi.sym_eigvalsh(m)
    cond = ev[:, -1] / (ev[:, 0] + 1e-10)
    logdet = torch.log(ev.clamp_min(1e-38)).sum(-1)
    return {'min_eigenvalue': ev[:, 0].min().item(), 'max_eigenvalue': ev[:, -1].max().item(),
            'mean_condition_number': cond.mean().item(),
            'determinant_range': (torch.exp(logdet.min()).item(), torch.exp(logdet.max()).item())}


def matrices_from_cholesky_factors(l_factors: torch.Tensor) -> torch.Tensor:
    """M_k = L_k L_k^T (RHVAE parametrisation, ref rhvae_model.py:172-178): a plain batched GEMM."""
    return l_factors @ l_factors.transpose(-1, -2)
