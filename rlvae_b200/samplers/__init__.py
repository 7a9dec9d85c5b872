from .base_sampler import BaseRiemannianSampler, MetricModel, tables_for
from .hmc_sampler import RiemannianHMCSampler
from .rhvae_sampler import OfficialRHVAESampler, RHVAEStyleHMCSampler
from .riemannian_sampler import WorkingRiemannianSampler

__all__ = ['BaseRiemannianSampler', 'MetricModel', 'RiemannianHMCSampler', 'RHVAEStyleHMCSampler', 'OfficialRHVAESampler', 'WorkingRiemannianSampler',
           'tables_for']
