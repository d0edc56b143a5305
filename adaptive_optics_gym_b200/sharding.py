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
    """Reduce {sum, sum of squares, count, min, max} of per-env episode returns over all ranks (one all-gather).

    ``local_returns``: 1-D torch tensor on the rank's device.  Returns a dict of Python floats
    identical on every rank.  Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    r = local_returns.to(torch.float64)
    # ONE collective: every rank contributes its five partial statistics (an all-gather of 40 bytes per rank); the
    # final reduction over ranks runs on the host after a single device->host copy
    part = torch.stack([r.sum(), (r * r).sum(), torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device),
                        r.min(), r.max()])
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        parts = [torch.empty(5, dtype=torch.float64, device=r.device) for _ in range(world)]
        dist.all_gather(parts, part, group=group)
        allp = torch.stack(parts)
    else:
        allp = part.unsqueeze(0)
    h = allp.cpu()
    total, sq, cnt = (float(h[:, k].sum()) for k in range(3))
    mean = total / cnt
    var = max(sq / cnt - mean * mean, 0.0)
    return dict(count=int(cnt), mean=mean, std=var ** 0.5, min=float(h[:, 3].min()), max=float(h[:, 4].max()))
