"""CPU, world_size 2, gloo: the N>1 host logic — rank 0 builds the calibration tables, broadcasts
them, rank 1 receives bit-identical tables; the frame shards tile the sequence exactly."""
import os
import socket
import sys

import numpy as np
import pytest

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
