"""Instance sharding across GPUs: every OCP instance is independent, so the instance
axis is cut into contiguous blocks, one per rank, with no exchange step on the solve path."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous block of `total` instances (block sizes differ by at most 1)"""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def mixed_split(total: int) -> dict:
    """BASELINE config 5: a mixed batch split 1/3 omni4, 1/3 diff, 1/3 tric (remainder to omni4 first)"""
    base, rem = divmod(total, 3)
    names = ("omni4", "diff", "tric")
    return {n: base + (1 if i < rem else 0) for i, n in enumerate(names)}
