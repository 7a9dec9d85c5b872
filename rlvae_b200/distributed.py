"""Multi-GPU plumbing for the metric path: batch sharding only.

Every point / chain is independent given the tables, so the N axis is cut into
contiguous slices ``[r*N/W, (r+1)*N/W)``, the tables are replicated with one
broadcast at load time, and there is NO collective inside ``compute_*`` or inside
the leapfrog loop.  Collectives appear only where a caller wants results on every
rank (``all_gather``) or a scalar reduced (``all_reduce``).  The reference itself is
single-GPU (conf/training/*.yaml: devices=1), so this module has no reference
counterpart; it works on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank ``rank``; sizes differ by at most one row block."""
    return (rank * n) // world, ((rank + 1) * n) // world


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def broadcast_tables(metric_tensor, src: int = 0, group=None) -> None:
    """Replicate centroids / matrices / temperature / regularization from ``src``."""
    world, _ = world_info(group)
    if world == 1:
        return
    for name in ('centroids', 'metric_matrices', 'temperature', 'regularization'):
        t = getattr(metric_tensor, name)
        dist.broadcast(t, src=src, group=group)
    metric_tensor._tab = None


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row slices (produced with ``shard_bounds``) on every rank."""
    world, rank = world_info(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def sharded_apply(fn: Callable[[torch.Tensor], Dict[str, torch.Tensor]], z_global: torch.Tensor,
                  gather: bool = True, group=None) -> Dict[str, torch.Tensor]:
    """Run ``fn`` (e.g. ``MetricTensor.evaluate``) on this rank's slice of ``z_global``;
    optionally all_gather every returned [n_r, ...] tensor back to [N, ...]."""
    world, rank = world_info(group)
    n = z_global.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    out = fn(z_global[lo:hi].contiguous())
    if not gather or world == 1:
        return out
    return {k: (all_gather_rows(v, n, group) if isinstance(v, torch.Tensor) and v.dim() >= 1
                and v.shape[0] == hi - lo else v) for k, v in out.items()}


def all_reduce_scalar(x: torch.Tensor, op=None, group=None) -> torch.Tensor:
    """Sum (default) of a scalar loss / acceptance count over ranks."""
    world, _ = world_info(group)
    if world > 1:
        dist.all_reduce(x, op=op or dist.ReduceOp.SUM, group=group)
    return x


def shard_hmc_streams(z0: torch.Tensor, gammas: torch.Tensor, accs: torch.Tensor, group=None):
    """Rows of the globally generated HMC draws that belong to this rank, so results do not
    depend on the world size."""
    world, rank = world_info(group)
    lo, hi = shard_bounds(z0.shape[0], world, rank)
    return z0[lo:hi].contiguous(), gammas[:, lo:hi].contiguous(), accs[:, lo:hi].contiguous()
