"""Multi-GPU plumbing for the merged map: one process per GPU (torch.distributed), no
collective on the detect/pose path, ONE all-gather of the fixed-size per-rank map tables followed by
the same deterministic merge on every rank."""
from __future__ import annotations

import numpy as np


def allgather_tables(table: np.ndarray, dist, device=None) -> np.ndarray:
    """table: uint8[table_bytes] (fid_map_export).  Returns uint8[world, table_bytes] in rank order.
    Backend nccl (device='cuda') or gloo (device=None, CPU tests)."""
    import torch

    world = dist.get_world_size()
    mine = torch.from_numpy(np.ascontiguousarray(table, np.uint8))
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return torch.stack(out).cpu().numpy()


def shard_frames(n_frames: int, rank: int, world: int):
    """Contiguous frame shard of a stream for this rank (frames are independent)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)
