"""Multi-GPU plumbing for the frame-parallel path.

Frames are independent, so the only communication is a one-time broadcast of the four calibration
tables from rank 0 (which parsed the files and built the tables on its host); afterwards every rank
prepares its own contiguous shard of the sequence with no collective.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) is the transport; the tables land in tensors that the device context
adopts in place (mdc_ctx_create_from_device_tables)."""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of the frame sequence for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_calibration(fov, photo, device, src: int = 0):
    """Rank `src` passes its host models (api.UndistorterFOV / api.PhotometricUndistorter); the other
    ranks pass None.  Returns (dims, (remap_x, remap_y, ginv, vinv)) with the tables as tensors on
    `device` on every rank.  A missing table (invalid object) is broadcast as None."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    meta = torch.zeros(8, dtype=torch.int64, device=device)
    if rank == src:
        iw, ih = fov.getInputDims()
        ow, oh = fov.getOutputDims()
        meta[:] = torch.tensor([iw, ih, ow, oh, int(fov.isValid()), int(photo.validGamma), int(photo.validVignette), 0])
    dist.broadcast(meta, src=src)
    iw, ih, ow, oh, has_fov, has_g, has_v, _ = [int(v) for v in meta.tolist()]

    def bcast(n, host_array):
        t = torch.empty(n, dtype=torch.float32, device=device)
        if rank == src:
            t.copy_(torch.from_numpy(np.ascontiguousarray(host_array, np.float32)))
        dist.broadcast(t, src=src)
        return t

    rx = ry = g = v = None
    if has_fov:
        tabs = fov.remap_tables() if rank == src else (None, None)
        rx, ry = bcast(ow * oh, tabs[0]), bcast(ow * oh, tabs[1])
    if has_g:
        g = bcast(256, photo.getGInv() if rank == src else None)
    if has_v:
        v = bcast(iw * ih, photo.vignette_maps()[1] if rank == src else None)
    return (iw, ih, ow, oh), (rx, ry, g, v)
