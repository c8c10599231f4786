"""Multi-GPU sharding: independent streams / frame batches go to ranks round-robin, with no collective on
the hot path; one all-reduce gathers the per-rank BER/CRC counters.  The reference has
no counterpart (single host, GNU Radio threads); this only scales the batch dimension."""
from __future__ import annotations

from typing import List, Sequence

COUNTER_NAMES = ("frames", "messages", "crc_ok", "payload_bytes_ok", "samples", "triggers", "vectors", "reserved")


def streams_of_rank(n_streams: int, world: int, rank: int) -> List[int]:
    """Stream s is processed by rank s mod world."""
    return [s for s in range(n_streams) if s % world == rank]


def split_frames(n_frames: int, world: int, rank: int):
    """Contiguous frame range [lo, hi) of a long stream cut at frame boundaries on the transmit side."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_counters(counters, group=None):
    """SUM all-reduce of the int64[8] counter tensor (NCCL for cuda tensors, gloo for cpu tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def as_dict(counters: Sequence[int]):
    return {k: int(v) for k, v in zip(COUNTER_NAMES, counters)}
