"""rlvae_b200 -- B200-native (sm_100a) metric evaluation + sampling for RlVAE.

Drop-in names for the reference's hot-path API: ``MetricTensor``,
``MetricLoader``, ``BaseRiemannianSampler``, ``RiemannianHMCSampler``,
``WorkingRiemannianSampler``, ``FlowManager``.  See DESIGN.md.
"""
from .flow_manager import FlowManager
from .metric_loader import MetricLoader
from .metric_tensor import MetricTensor
from .samplers import (BaseRiemannianSampler, MetricModel, OfficialRHVAESampler, RHVAEStyleHMCSampler,
                       RiemannianHMCSampler, WorkingRiemannianSampler)

__all__ = ['MetricTensor', 'MetricLoader', 'BaseRiemannianSampler', 'MetricModel',
           'RiemannianHMCSampler', 'RHVAEStyleHMCSampler', 'OfficialRHVAESampler', 'WorkingRiemannianSampler',
           'FlowManager']
__version__ = '0.1.0'
