"""Multi-GPU plumbing: frame-sharded frame preparation and the pixel-sharded responseCalib loop.

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


# ------------------------------------------------------------------------------------------------------------------------
# Pixel-sharded responseCalib.  Every pass of the calibrator is per-pixel work
# (main_responseCalib.cpp:317-346); only the G-step's 2 x 256 accumulators (:290-299) and rmse's {error, count}
# pair (:50-69) are global.  Each rank holds its slice of every image and the loop below is main()'s loop (:281-362)
# with one all-reduce per reduction.  `ops` does the per-slice work: api.Context on GPUs, a numpy stand-in in the
# world-size-2 gloo test — the host logic is the same object either way.

def shard_pixels(npix: int, rank: int, world: int, align: int = 128):
    """Contiguous [begin, end) of the pixel range for `rank`: balanced in units of `align` pixels (slices that are a multiple of
    16 pixels keep the bulk-copy streaming kernels; 128 = one warp task), the last rank takes the ragged tail."""
    units = (npix + align - 1) // align
    b, e = shard_range(units, rank, world)
    return min(b * align, npix), min(e * align, npix)


def response_calib_sharded(ops, data_local, t, nits, E_local, G, group=None):
    """main()'s optimisation loop on this rank's pixel slice.  data_local uint8 [n, npix_local], t float64 [n],
    E_local float64 [npix_local] (out), G float64 [256] (out, identical on every rank).  Returns the log [nits, 4] =
    (rmse after G-step, after E-step, after rescale, sample count) — global values, identical on every rank."""
    import torch
    import torch.distributed as dist

    dev = data_local.device
    scale4 = torch.zeros(4, dtype=torch.int64, device=dev)        # bit patterns of max |E|, max |t|, bad-t flag: all-reduce MAX
    limbs = torch.zeros(768, dtype=torch.int64, device=dev)       # the bins' fixed-point sums as 3 x 43-bit limbs: all-reduce SUM
    special = torch.zeros(256, dtype=torch.float64, device=dev)   # fp64 sums of non-finite products: all-reduce SUM
    gnum = torch.zeros(256, dtype=torch.int64, device=dev)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    # the counts depend on the images only: all-reduced once, then carried over (bulk-copy path only, like mdc_response_calib)
    reusable = data_local.shape[1] % 16 == 0 and data_local.data_ptr() % 16 == 0 and getattr(ops, "can_reuse_counts", True)
    if distributed:       # the ranks must agree on it, or they would disagree on which all-reduces happen
        flag = torch.tensor([1 if reusable else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        reusable = bool(flag.item())

    def all_reduce(x, op=None):
        if distributed:
            dist.all_reduce(x, op=op or dist.ReduceOp.SUM, group=group)

    def rmse():
        ops.rc_rmse_accumulate(data_local, t, G, E_local, acc)
        all_reduce(acc)
        a = acc.tolist()
        return 1e5 * float(np.sqrt(a[0] / a[1])) if a[1] else float("nan"), a[1]

    ops.rc_einit(data_local, E_local)
    G.zero_()
    log = np.zeros((nits, 4), np.float64)
    for it in range(nits):
        reuse = it > 0 and reusable
        # G-step with sums that are exact across ranks (G, and with it E, does not depend on the number of ranks)
        ops.rc_gstep_scale(E_local, t, scale4)
        all_reduce(scale4, dist.ReduceOp.MAX if distributed else None)
        ops.rc_gstep_accumulate_exact(data_local, t, E_local, scale4, limbs, special, gnum, reuse)
        all_reduce(limbs)
        all_reduce(special)
        if not reuse:
            all_reduce(gnum)
        ops.rc_gstep_finish_exact(scale4, limbs, special, gnum, G)
        log[it, 0] = rmse()[0]
        ops.estep(data_local, t, G, E_local)
        log[it, 1] = rmse()[0]
        ops.rc_rescale(E_local, G)            # factor = 255/G[255] from the replicated G: same on every rank, no collective
        log[it, 2], log[it, 3] = rmse()
    return log


# ------------------------------------------------------------------------------------------------------------------------
# Native NCCL path (include/mdc_b200_nccl.h, libmdc_b200_nccl.so): what a C++ host uses — the launcher's process group is
# only the side channel that carries the 128-byte NCCL unique id from rank 0 to its peers.

class NativeComm:
    """An ncclComm_t owned by libmdc_b200_nccl.so for this rank: table broadcast at init (mdc_ctx_create_broadcast) and
    the pixel-sharded calibrator loop (mdc_response_calib_sharded) run on it, entirely in C++."""

    def __init__(self, device: int, group=None):
        import ctypes as C
        import os
        import torch
        import torch.distributed as dist
        from . import _lib
        path = os.path.join(_lib.PKG, "lib", "libmdc_b200_nccl.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing (needs the system NCCL headers at build time)")
        self.lib = L = C.CDLL(path)
        L.mdc_nccl_unique_id.argtypes = [C.c_char_p]
        L.mdc_nccl_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.mdc_nccl_comm_destroy.argtypes = [C.c_void_p]
        L.mdc_ctx_create_broadcast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.mdc_response_calib_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        self.rank, self.world, self.device = dist.get_rank(group), dist.get_world_size(group), device
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(L.mdc_nccl_unique_id(uid), "mdc_nccl_unique_id")
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
        on_gpu = dist.get_backend(group) == "nccl"
        if on_gpu:
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, src=0, group=group)
        uid = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        h = C.c_void_p()
        _lib.check(L.mdc_nccl_comm_create(uid, self.world, self.rank, device, C.byref(h)), "mdc_nccl_comm_create")
        self.h = h
        self.version = int(L.mdc_nccl_version())

    def create_context(self, fov, photo):
        """mdc_ctx_create_broadcast: rank 0 passes its host models, the others None; every rank gets a context owning its tables."""
        import ctypes as C
        from . import _lib, api
        h = C.c_void_p()
        _lib.check(self.lib.mdc_ctx_create_broadcast(self.h, self.rank, 0, self.device, fov._h if fov is not None else None,
                                                     photo._h if photo is not None else None, C.byref(h)), "mdc_ctx_create_broadcast")
        ctx = api.Context.__new__(api.Context)
        ctx._h, ctx.device, ctx.fov, ctx.photo, ctx._keep = h, self.device, fov, photo, None
        return ctx

    def response_calib_sharded(self, ctx, data_local, t, nits, E_local, G):
        """The whole pixel-sharded loop in C++ over this communicator (two ncclAllReduce kinds per iteration)."""
        import ctypes as C
        import torch
        from . import _lib
        torch.cuda.current_stream(data_local.device).synchronize()
        log = np.zeros((max(nits, 1), 4), np.float64)
        _lib.check(self.lib.mdc_response_calib_sharded(ctx._h, self.h, self.device, C.c_void_p(data_local.data_ptr()), data_local.shape[0],
                                                       data_local.shape[1], C.c_void_p(t.data_ptr()), nits, C.c_void_p(E_local.data_ptr()),
                                                       C.c_void_p(G.data_ptr()), log.ctypes.data_as(C.POINTER(C.c_double))),
                   "mdc_response_calib_sharded")
        return log[:nits]

    def close(self):
        if getattr(self, "h", None):
            self.lib.mdc_nccl_comm_destroy(self.h)
            self.h = None
