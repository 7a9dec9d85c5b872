"""Golden vectors for the metric-construction step, produced by executing the REAL reference lines
(scripts/train_and_extract_vanilla_vae.py:204-226 -- the per-centroid weighted-covariance loop) on
seeded synthetic latents.  TEST INFRASTRUCTURE ONLY; needs /root/reference (build container).

    python -m oracle.make_golden_builder

The reference code is a script body, not a function, so the loop is sliced out of the file by line
number (asserting on its first and last line) and exec'd with `all_mus`, `centroids`, `temperature`,
`regularization`, `latent_dim` bound to our inputs; `tqdm` is replaced by the identity."""
from __future__ import annotations

import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference/scripts/train_and_extract_vanilla_vae.py'


def reference_loop_source():
    lines = open(REF).read().split('\n')
    seg = lines[203:226]                       # 1-based 204..226
    assert seg[0].strip().startswith('for i, c in enumerate(tqdm(centroids))'), seg[0]
    assert seg[-1].strip().startswith('M_matrices = torch.stack(M_matrices, dim=0)'), seg[-1]
    return textwrap.dedent('\n'.join(seg))


def run_reference(all_mus, centroids, temperature, regularization):
    ns = {'torch': torch, 'tqdm': (lambda x: x), 'all_mus': all_mus, 'centroids': centroids,
          'temperature': temperature, 'regularization': regularization, 'latent_dim': all_mus.shape[1],
          'M_matrices': [], 'print': (lambda *a, **k: None)}
    exec(reference_loop_source(), ns)
    return ns['M_matrices']


def make_latents(n, d, n_clusters, seed):
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(n_clusters, d, generator=g)
    which = torch.randint(0, n_clusters, (n,), generator=g)
    spread = 0.05 + 0.25 * torch.rand(n_clusters, 1, generator=g)
    aniso = 0.3 + torch.rand(n_clusters, d, generator=g)
    x = centers[which] + spread[which] * aniso[which] * torch.randn(n, d, generator=g)
    idx = torch.randperm(n, generator=g)[: 3 * n_clusters]
    return x, x[idx].clone()


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, (n, d, ncl, T, reg, seed) in {
            'builder_d16_T01': (3000, 16, 12, 0.1, 0.01, 0),       # the reference's own T = 0.1, reg = 0.01
            'builder_d16_T05': (3000, 16, 12, 0.5, 0.01, 1),
            'builder_d2_T03': (800, 2, 6, 0.3, 0.01, 2),
            'builder_d32_T10': (1500, 32, 8, 1.0, 0.01, 3)}.items():
        x, c = make_latents(n, d, ncl, seed)
        M = run_reference(x, c, T, reg)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), latents=x.numpy(), centroids=c.numpy(),
                            temperature=np.float64(T), regularization=np.float64(reg), M=M.numpy())
        print('wrote', name, tuple(M.shape), 'min eig', torch.linalg.eigvalsh(M.double()).min().item())


if __name__ == '__main__':
    main()
