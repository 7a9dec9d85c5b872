"""Golden vectors at the BENCHMARK's table size (K = 10,000, d = 16: BASELINE.json configs[1]/[2]),
produced by the REAL reference code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_k10k

* ``hmc_d16_k10k``: ``RiemannianHMCSampler.sample`` (ref src/models/samplers/hmc_sampler.py:104-165)
  on 256 chains x 2 MCMC iterations x 20 leapfrog steps with the random draws recorded, plus the
  per-iteration Hamiltonians / acceptance ratios / decisions / states of the oracle's restatement of
  the same loop -- stored only after checking that the oracle's final state equals the reference's.
* ``losses_d16_k10k``: see make_golden_losses.py (kept separate: other reference modules).

The tables are NOT stored: tests regenerate them from the same seed
(rlvae_b200.synthetic.make_synthetic_metric(10000, 16, seed=0)).
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import metric_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import _quiet, _save  # noqa: E402
from rlvae_b200.synthetic import make_synthetic_metric  # noqa: E402


def hmc_k10k(name='hmc_d16_k10k', n=256, mcmc=2, n_lf=20, eps=0.03, beta_zero=1.0, seed=23):
    sm = make_synthetic_metric(10000, 16, seed=0)
    c, M, T, lam = sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization
    mods = ref_loader.modules()
    mt = ref_loader.make_ref_metric(c, M, T, lam)
    model = ref_loader.RefModel(mt)
    with _quiet():
        s = mods['hmc_sampler'].RiemannianHMCSampler(model, mcmc_steps_nbr=mcmc, n_lf=n_lf, eps_lf=eps,
                                                     beta_zero=beta_zero)
    torch.manual_seed(seed)
    t0 = time.time()
    with ref_loader.RecordingRNG() as rec:
        zf = s.sample(n)
    print(f'reference sample({n}) x {mcmc} x {n_lf}: {time.time() - t0:.1f} s')
    assert [k for k, _ in rec.draws] == ['randn'] + ['randn_like', 'rand'] * mcmc
    z0 = rec.draws[0][1]
    gam = torch.stack([rec.draws[1 + 2 * i][1] for i in range(mcmc)])
    acc = torch.stack([rec.draws[2 + 2 * i][1] for i in range(mcmc)])
    orec = {}
    t0 = time.time()
    zo = O.hmc_sample((c, M, T, lam), z0, gam, acc, n_lf, eps, beta_zero, record=orec)
    print(f'oracle chain: {time.time() - t0:.1f} s; max |oracle - reference| = {(zo - zf).abs().max().item():.3e}')
    assert torch.equal(zo, zf), 'the oracle restatement must reproduce the reference chain exactly'
    _save(name, z0=z0, gamma=gam, acc=acc, n_lf=np.int64(n_lf), eps_lf=np.float64(eps),
          beta_zero=np.float64(beta_zero), z_final=zf, table_seed=np.int64(0), n_centroids=np.int64(10000),
          rec_H0=torch.stack(orec['H0']), rec_H=torch.stack(orec['H']), rec_alpha=torch.stack(orec['alpha']),
          rec_moves=torch.stack(orec['moves']), rec_z=torch.stack(orec['z']))


def main():
    warnings.simplefilter('ignore')
    assert ref_loader.available(), 'needs /root/reference'
    hmc_k10k()


if __name__ == '__main__':
    main()
