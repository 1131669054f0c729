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


def init_engine_comm(engine, group=None):
    """Give ``engine`` (a LikelihoodEngine) its own NCCL communicator, created INSIDE libsfb200 (sfb_comm_init):
    rank 0 makes the unique id, ``torch.distributed`` only carries those 128 bytes to the other ranks once at
    set-up.  After this ``engine.allgather_lnl`` is the per-step exchange and no torch collective is on the path.
    Without an initialised process group a single-rank communicator is created."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        engine.comm_init(0, 1, engine.comm_unique_id())
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    engine.comm_init(rank, world, box[0])
