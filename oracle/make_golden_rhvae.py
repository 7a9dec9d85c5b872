"""Golden vectors for the pythae-variant HMC (A8): runs the REAL
src/lib/src/pythae/samplers/manifold_sampler/rhvae_sampler.py::RHVAESampler.hmc_sampling (the loop
behind OfficialRHVAESampler.sample_prior) on a stand-in `self` that carries exactly the attributes
the method touches, with every RNG draw recorded.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python -m oracle.make_golden_rhvae
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from rlvae_b200.synthetic import make_synthetic_metric  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def main():
    ref_loader.flow_modules()        # installs the sklearn_extra / imageio stubs and the pythae path
    rs = importlib.import_module('pythae.samplers.manifold_sampler.rhvae_sampler')
    RHVAESampler = rs.RHVAESampler
    for name, (K, n, steps, n_lf, eps, beta0, seed) in {
            'rhvae_hmc_d16_k120': (120, 24, 3, 4, 0.03, 1.0, 0),
            'rhvae_hmc_d16_k120_beta03': (120, 32, 2, 6, 0.6, 0.3, 1)}.items():
        sm = make_synthetic_metric(K, 16, seed=seed)
        mt = ref_loader.make_ref_metric(sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
        model = ref_loader.RefModel(mt)
        fake = types.SimpleNamespace(
            model=model, device=torch.device('cpu'), mcmc_steps_nbr=steps, n_lf=torch.tensor([n_lf]),
            eps_lf=torch.tensor([eps]), beta_zero_sqrt=torch.tensor([beta0]).sqrt(),
            log_pi=RHVAESampler.log_sqrt_det_G_inv, grad_func=RHVAESampler.grad_log_prop)
        torch.manual_seed(100 + seed)
        with ref_loader.RecordingRNG() as rng:
            z = RHVAESampler.hmc_sampling(fake, n)
        kinds = [k for k, _ in rng.draws]
        assert kinds == ['randint'] + ['randn_like', 'rand'] * steps, kinds
        idx0 = rng.draws[0][1]
        gammas = torch.stack([rng.draws[1 + 2 * i][1] for i in range(steps)])
        accs = torch.stack([rng.draws[2 + 2 * i][1] for i in range(steps)])
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), centroids=sm.centroids.numpy(),
                            matrices=sm.metric_matrices.numpy(), temperature=np.float64(sm.temperature),
                            regularization=np.float64(sm.regularization), idx0=idx0.numpy(),
                            gamma=gammas.numpy(), acc=accs.numpy(), n_lf=np.int64(n_lf),
                            eps_lf=np.float64(eps), beta_zero=np.float64(beta0), z_final=z.numpy())
        print('wrote', name, tuple(z.shape), 'moved', int((z != sm.centroids[idx0]).any(dim=1).sum()), 'of', n)


if __name__ == '__main__':
    main()
