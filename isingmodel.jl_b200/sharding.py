"""Replica sharding over the GPUs of one box: chains are independent, so rank g owns a
contiguous block of replicas and no data-path collective exists.  torch.distributed is plumbing only:
barrier + max-over-ranks timing + gathering per-replica observables to rank 0."""
from __future__ import annotations

import numpy as np


def replica_range(R: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the replicas owned by ``rank`` when R replicas are split over ``world`` ranks as evenly
    as possible (the first R % world ranks own one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(R, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_replica_values(local: np.ndarray, R: int, group=None) -> np.ndarray | None:
    """All ranks contribute their per-replica values (first axis = local replicas, in replica order);
    rank 0 gets the concatenation in global replica order, the others ``None``."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = [replica_range(R, r, world)[1] - replica_range(R, r, world)[0] for r in range(world)]
    loc = np.ascontiguousarray(local, dtype=np.float64)
    tail = loc.shape[1:]
    pad = max(counts)
    buf = torch.zeros((pad,) + tail, dtype=torch.float64, device=dev)
    buf[: loc.shape[0]] = torch.from_numpy(loc).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)], axis=0)


def max_over_ranks(value: float, group=None) -> float:
    """Max of a scalar over all ranks (device times are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
