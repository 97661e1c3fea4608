"""Multi-GPU plumbing for the replica-parallel path (inference / NMS shard by image, no data-path collective;
SURVEY §8e): rank-local sharding of an image list and the max-over-ranks reduction of device timings that
bench.py reports.  torch.distributed is plumbing only (nccl on GPUs, gloo in the CPU tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous, disjoint, exhaustive split of [0, n_items) — rank r gets the r-th slice (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, device="cpu") -> float:
    """The slowest rank defines the job's time (bench contract: max over ranks)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def broadcast_model_state(model, src: int = 0) -> int:
    """What DistributedDataParallel does at construction (train.py:214 `DDP(model, ...)`): every parameter and every
    buffer of rank `src` replaces the other ranks' copies, so replicas seeded per rank (init_seeds(1 + RANK),
    train.py:100) start from identical weights and BatchNorm statistics.  In place (`.data.copy_`-free: the broadcast
    writes the storages the device plans point at).  Returns the number of tensors sent."""
    rank, ws = world()
    if ws == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
            n += 1
    return n
