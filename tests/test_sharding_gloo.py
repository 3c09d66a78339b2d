"""CPU, world_size 2, gloo: the N>1 host logic — rank 0 builds the calibration tables, broadcasts
them, rank 1 receives bit-identical tables; the frame shards tile the sequence exactly."""
import os
import socket
import sys

import numpy as np
import pytest

import gstep_fixed_point_spec as fx
from conftest import ROOT
from mono_dataset_code_b200 import sharding


def test_shard_ranges_tile_the_sequence():
    for n in (0, 1, 7, 256, 10000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, cam, pcalib, vig, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mono_dataset_code_b200 import api, sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fov = photo = None
    if rank == 0:
        fov = api.UndistorterFOV(cam)
        photo = api.PhotometricUndistorter(pcalib, vig, 320, 240)
    dims, tabs = sh.broadcast_calibration(fov, photo, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), dims=np.array(dims), rx=tabs[0].numpy(), ry=tabs[1].numpy(),
             g=tabs[2].numpy(), v=tabs[3].numpy(), shard=np.array(sh.shard_range(1001, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_table_broadcast_world2_gloo(dataset_dir, tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    files = dataset_dir((320, 240, 300, 200, "crop", (0.35, 0.44, 0.49, 0.5, 0.93)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, files["camera"], files["pcalib"], files["vignette"], str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert list(a["dims"]) == [320, 240, 300, 200] == list(b["dims"])
    for k in ("rx", "ry", "g", "v"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    from mono_dataset_code_b200 import api
    ref = api.UndistorterFOV(files["camera"]).remap_tables()
    assert np.array_equal(b["rx"].view(np.uint32), ref[0].view(np.uint32))
    assert list(a["shard"]) == [0, 501] and list(b["shard"]) == [501, 1001]


# ------------------------------------------------------------------------------------------------ pixel-sharded responseCalib
class NumpyCalibOps:
    """CPU stand-in for api.Context's calibrator passes (torch CPU tensors in, numpy inside): lets the world-size-2 gloo test run the
    SAME host loop (sharding.response_calib_sharded) that drives the CUDA kernels on GPUs.  Test infrastructure."""
    can_reuse_counts = True

    def __init__(self, port):
        self.port = port

    def rc_einit(self, data, E):
        E.copy_(__import__("torch").from_numpy(self.port.einit(data.numpy())))

    # the G-step in fixed point, from the numpy specification of what the CUDA kernels compute (tests/gstep_fixed_point_spec.py)
    def rc_gstep_scale(self, E, t, scale4):
        import torch
        scale4.copy_(torch.from_numpy(fx.scale_words(E.numpy(), t.numpy())))

    def rc_gstep_accumulate_exact(self, data, t, E, scale4, limbs, special, gnum, reuse_counts):
        import torch
        l, sp, cnt = fx.accumulate(data.numpy(), t.numpy(), E.numpy(), scale4.numpy())
        limbs.copy_(torch.from_numpy(l))
        special.copy_(torch.from_numpy(sp))
        if not reuse_counts:
            gnum.copy_(torch.from_numpy(cnt))

    def rc_gstep_finish_exact(self, scale4, limbs, special, gnum, G):
        import torch
        G.copy_(torch.from_numpy(fx.finish(scale4.numpy(), limbs.numpy(), special.numpy(), gnum.numpy())))

    def estep(self, data, t, G, E):
        E.copy_(__import__("torch").from_numpy(self.port.estep(data.numpy(), t.numpy(), G.numpy())))

    def rc_rmse_accumulate(self, data, t, G, E, acc):
        d = data.numpy()
        r = G.numpy()[d] - t.numpy()[:, None] * E.numpy()[None, :]
        ok = (d != 255) & np.isfinite(r)
        acc[0] = float(np.sum(r[ok] * r[ok] * 1e-10))
        acc[1] = float(np.count_nonzero(ok))

    def rc_rescale(self, E, G):
        e, g = E.numpy(), G.numpy()          # views: scaled in place
        return self.port.rescale(e, g)


def _calib_worker(rank, world, port_no, data_path, out_dir, nits):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from mono_dataset_code_b200 import sharding as sh
    from oracle import loader
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(data_path)
    data, t = z["data"], z["t"]
    lo, hi = sh.shard_pixels(data.shape[1], rank, world, align=16)
    local = torch.from_numpy(np.ascontiguousarray(data[:, lo:hi]))
    E = torch.zeros(hi - lo, dtype=torch.float64)
    G = torch.zeros(256, dtype=torch.float64)
    log = sh.response_calib_sharded(NumpyCalibOps(loader.PortOracle()), local, torch.from_numpy(t), nits, E, G)
    np.savez(os.path.join(out_dir, f"calib{rank}.npz"), lo=lo, hi=hi, E=E.numpy(), G=G.numpy(), log=log)
    dist.barrier()
    dist.destroy_process_group()


def test_pixel_sharded_response_calib_world2_gloo(port, tmp_path):
    """the calibrator's loop with the image stack split by pixel range over 2 ranks (gloo) reproduces the unsharded
    loop: identical G on both ranks, E slices that tile the unsharded E, the same rmse log (sums regrouped: <= 1e-10 relative)."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    n, npix, nits = 24, 1000, 3                                    # 1000 pixels: ranks get 512 and 488 (ragged, not a multiple of 16)
    t = np.linspace(0.5, 12.0, n)
    scene = rng.uniform(2.0, 30.0, npix)
    data = np.clip(scene[None, :] * t[:, None] ** 0.9 + rng.normal(0, 1.5, (n, npix)), 0, 255).astype(np.uint8)
    data[:, 17] = 255                                              # a pixel that is saturated in every image: E = 0/0
    np.savez(tmp_path / "stack.npz", data=data, t=t)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    mp.spawn(_calib_worker, args=(2, port_no, str(tmp_path / "stack.npz"), str(tmp_path), nits), nprocs=2, join=True)
    a, b = np.load(tmp_path / "calib0.npz"), np.load(tmp_path / "calib1.npz")
    assert (int(a["lo"]), int(a["hi"]), int(b["lo"]), int(b["hi"])) == (0, 512, 512, 1000)
    assert np.array_equal(a["G"], b["G"], equal_nan=True) and np.array_equal(a["log"], b["log"])
    # the unsharded loop (oracle restatement of main_responseCalib.cpp:281-362)
    E = port.einit(data)
    exp_log = np.zeros((nits, 4))
    for it in range(nits):
        G = port.gstep(data, t, E)
        exp_log[it, 0] = port.rmse(data, t, G, E)[0]
        E = port.estep(data, t, G)
        exp_log[it, 1] = port.rmse(data, t, G, E)[0]
        port.rescale(E, G)
        exp_log[it, 2:] = port.rmse(data, t, G, E)
    got_E = np.concatenate([a["E"], b["E"]])
    # exact sums across ranks: the same loop on ONE rank (no process group) gives the same bits of G and E
    import torch
    E1, G1 = torch.zeros(npix, dtype=torch.float64), torch.zeros(256, dtype=torch.float64)
    from oracle import loader
    sharding.response_calib_sharded(NumpyCalibOps(loader.PortOracle()), torch.from_numpy(data), torch.from_numpy(t), nits, E1, G1)
    assert np.array_equal(a["G"], G1.numpy(), equal_nan=True) and np.array_equal(got_E, E1.numpy(), equal_nan=True)
    np.testing.assert_allclose(a["G"], G, rtol=1e-10, atol=0)
    np.testing.assert_allclose(got_E, E, rtol=1e-10, atol=0, equal_nan=True)
    assert np.isnan(got_E[17]) and np.isnan(E[17])
    np.testing.assert_allclose(a["log"], exp_log, rtol=1e-9)
    assert a["log"][-1, 3] == exp_log[-1, 3]                      # sample counts are exact


def test_pixel_shards_tile_the_image():
    for npix in (0, 1, 127, 128, 1000, 1000 * 1000, 1920 * 1080):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_pixels(npix, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == npix
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(b % 128 == 0 or b == npix for b, _ in spans)
