"""Multi-GPU layout of the ``AO-v0`` step path: environments shard, nothing else.

Environments are independent (no shared mutable state, tables are read-only and replicated), so
rank r of G owns the contiguous block ``shard_range(num_envs, r, G)`` and the step path has NO
collective.  The only exchange is the end-of-episode gather of per-env returns (a few KB,
latency-bound): ``gather_episode_stats`` -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(num_envs: int, rank: int, world_size: int):
    """Contiguous block of global env ids owned by ``rank``: (first, count).  The first
    ``num_envs % world_size`` ranks take one extra env."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    base, extra = divmod(num_envs, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_episode_stats(local_returns, group=None):
    """All-reduce {sum, sum of squares, count, min, max} of per-env episode returns.

    ``local_returns``: 1-D torch tensor on the rank's device.  Returns a dict of Python floats
    identical on every rank.  Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    r = local_returns.to(torch.float64)
    s = torch.stack([r.sum(), (r * r).sum(), torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device)])
    mn, mx = r.min().reshape(1), r.max().reshape(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    total, sq, cnt = (float(v) for v in s.cpu())
    mean = total / cnt
    var = max(sq / cnt - mean * mean, 0.0)
    return dict(count=int(cnt), mean=mean, std=var ** 0.5, min=float(mn.cpu()), max=float(mx.cpu()))
