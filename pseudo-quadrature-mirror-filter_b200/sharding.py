"""Row sharding for multi-GPU runs.  Every (clip, channel) row and every stream of the PQMF path is independent
(SURVEY.md 8e), so a job is split by contiguous row ranges, one process per GPU, with NO collective on the data path."""
from __future__ import annotations

from typing import Tuple


def shard_rows(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the rows owned by `rank`: contiguous, disjoint, covering, sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size) or n_rows < 0:
        raise ValueError(f"bad shard request: n_rows={n_rows} world_size={world_size} rank={rank}")
    base, extra = divmod(n_rows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
