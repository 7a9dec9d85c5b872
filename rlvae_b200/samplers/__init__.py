from .base_sampler import BaseRiemannianSampler, MetricModel, tables_for
from .hmc_sampler import RiemannianHMCSampler
from .riemannian_sampler import WorkingRiemannianSampler

__all__ = ['BaseRiemannianSampler', 'MetricModel', 'RiemannianHMCSampler', 'WorkingRiemannianSampler',
           'tables_for']
