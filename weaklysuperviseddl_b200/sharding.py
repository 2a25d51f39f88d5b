"""Image sharding across the GPUs of one box: every stage of the hot path is independent per
image, so ranks never exchange pixels.  One process per GPU; torch.distributed is only used to agree on counters
(near-threshold pixels, masks written) and, in the training configuration, for PyTorch's own DDP all-reduce.

Reference: the per-image loops of PsuedoMasks.py:47-76 and LayerCAM.py:96-120 (single GPU, one image at a time)."""
from __future__ import annotations

import os
from typing import Iterable, Iterator, List, Sequence, Tuple, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def rank_world() -> Tuple[int, int]:
    """(rank, world) from the initialised process group, else from torchrun's environment, else (0, 1)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Round-robin shard: item i belongs to rank i % world (BASELINE config 3: `i % world == rank`)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return range(rank, n_items, world)


def shard_batches(loader: Iterable[T], rank: int, world: int) -> Iterator[Tuple[int, T]]:
    """Yields (global batch index, batch) for the batches this rank owns."""
    for i, batch in enumerate(loader):
        if i % world == rank:
            yield i, batch


def chunk(indices: Sequence[int], size: int) -> List[Sequence[int]]:
    """Resident chunks of at most `size` images (config 3 keeps ~128 images of hooks on the device at a time)."""
    if size < 1:
        raise ValueError("chunk size must be positive")
    return [indices[i:i + size] for i in range(0, len(indices), size)]


def all_reduce_counters(values: Sequence[int], device=None) -> List[int]:
    """Sum of small integer counters over ranks (identity without a process group).  The only collective of
    pseudo-mask generation: bookkeeping, not data path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(v) for v in values]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def max_over_ranks(ms: float, device=None) -> float:
    """Elapsed time of a multi-GPU step = the slowest rank's device time (bench.py's timing rule)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
