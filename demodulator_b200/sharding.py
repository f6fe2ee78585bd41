"""Channel-parallel sharding across the GPUs of one box.

Units (channels / files / chunk-aligned ranges of one stream) are independent: there is no
data-path collective.  torch.distributed is used only for the start barrier and for reducing
the per-rank device time to its maximum.
"""
from __future__ import annotations


def assign_channels(n_channels: int, world: int, rank: int) -> list[int]:
    """Static round-robin: channel c runs on rank c % world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_channels, world))


def split_stream(n_chunks: int, world: int, rank: int, batch_chunks: int) -> list[tuple[int, int]]:
    """One long stream: round-robin batches of `batch_chunks` whole chunks; returns this rank's
    (first_chunk, n_chunks) ranges in stream order.  Chunks never straddle ranks, so no overlap
    or state hand-off is needed (the reference clears all filter memory per chunk)."""
    out = []
    first, b = 0, 0
    while first < n_chunks:
        n = min(batch_chunks, n_chunks - first)
        if b % world == rank:
            out.append((first, n))
        first += n
        b += 1
    return out


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a python float over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
