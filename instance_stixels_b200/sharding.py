"""Frame-level data parallelism: frames are independent, so a stream is cut into contiguous blocks,
one per rank (one process per GPU); there is no collective in the data path, only an optional
result gather.  The reference has no counterpart (single process, single GPU)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, last) of the frames rank `rank` owns: contiguous blocks, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def batches(first: int, last: int, max_batch: int):
    """Cut a shard into batches the library accepts (isx_initialize(max_batch))."""
    for b in range(first, last, max_batch):
        yield b, min(b + max_batch, last)


def gather_frames(local: Sequence[np.ndarray], n_frames: int, dist=None, dst: int = 0) -> List[np.ndarray] | None:
    """Gather per-frame result arrays (compact stixel lists, variable length) in frame order on `dst`.
    `dist` is torch.distributed (any backend) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world if rank == dst else None
    dist.gather_object(list(local), gathered, dst=dst)
    if rank != dst:
        return None
    out: List[np.ndarray] = []
    for r in range(world):
        out.extend(gathered[r])
    if len(out) != n_frames:
        raise RuntimeError(f"gathered {len(out)} frames, expected {n_frames}")
    return out


def compact_sections(sections: np.ndarray) -> np.ndarray:
    """[C][200] Section array -> flat array of the used entries plus their column (what travels)."""
    term = sections["type"] == -1
    n = np.where(term.any(axis=1), term.argmax(axis=1), sections.shape[1])
    mask = np.arange(sections.shape[1])[None, :] < n[:, None]
    cols = np.broadcast_to(np.arange(sections.shape[0])[:, None], sections.shape)[mask]
    out = np.zeros(int(mask.sum()), dtype=[("column", "<i4"), ("section", sections.dtype)])
    out["column"] = cols
    out["section"] = sections[mask]
    return out
