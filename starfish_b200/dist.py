"""Walker sharding across GPUs: walkers are independent, so the ensemble is split into
contiguous, equal-as-possible shards, one per rank, and the ONLY exchange per step is an all-gather of
the log-likelihood scalars.  `torch.distributed` is plumbing here (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_walkers: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the walkers owned by ``rank``; the first ``n % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_walkers, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_walkers: int, world: int):
    return [shard_range(n_walkers, r, world)[1] - shard_range(n_walkers, r, world)[0] for r in range(world)]


def gather_lnl(local, n_walkers: int, group=None):
    """All-gather the per-rank lnL shards into the full ``[n_walkers]`` vector on every rank.

    ``local`` is a 1-D tensor (CUDA with NCCL, CPU with gloo) holding this rank's shard in
    ``shard_range`` order.  Equal shards use one ``all_gather_into_tensor``; ragged shards pad to the
    largest shard first.  Without an initialised process group this is the identity.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_walkers, world)
    if local.numel() != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has the wrong length")
    mx = max(sizes)
    if min(sizes) == mx:
        out = torch.empty(n_walkers, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    buf = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + sizes[r]] for r in range(world)])
