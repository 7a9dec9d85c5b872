from .base_sampler import BaseRiemannianSampler, MetricModel, tables_for
from .hmc_sampler import RiemannianHMCSampler
from .rhvae_sampler import RHVAEStyleHMCSampler
from .riemannian_sampler import WorkingRiemannianSampler

__all__ = ['BaseRiemannianSampler', 'MetricModel', 'RiemannianHMCSampler', 'RHVAEStyleHMCSampler', 'WorkingRiemannianSampler',
           'tables_for']
