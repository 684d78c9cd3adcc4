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


class MapExchange:
    """Stream-ordered map exchange for one rank: export -> ONE NCCL all-gather -> merge, all
    enqueued on the map's own CUDA stream, so the host never blocks and the exchange overlaps detection.
    The buffers are torch tensors (torch is the device-memory / collective plumbing); the export and the
    merge are the library's kernels.  The merged view lives in the library (FiducialSlam.merged_entries)."""

    def __init__(self, slam, dist, device):
        import torch

        self.slam, self.dist, self.torch = slam, dist, torch
        self.world = dist.get_world_size() if dist is not None else 1
        nb = slam.table_bytes
        self.send = torch.empty(nb, dtype=torch.uint8, device=device)
        self.recv = torch.empty(nb * self.world, dtype=torch.uint8, device=device)
        self.stream = torch.cuda.ExternalStream(slam.cuda_stream(), device=device)

    def step(self, instance=0):
        """Enqueue one exchange epoch (returns at once).  Kernel launches of ours: 2 (export, merge)."""
        torch = self.torch
        self.slam.export_async(self.send.data_ptr(), instance)
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                self.dist.all_gather_into_tensor(self.recv, self.send)
        else:
            with torch.cuda.stream(self.stream):
                self.recv.copy_(self.send, non_blocking=True)
        self.slam.merge_device_async(self.recv.data_ptr(), self.world)
        return 2
