"""Native C++ multi-GPU init (include/mdc_b200_nccl.h): the helper library builds and exports its entry point (CPU);
on a 2-GPU box a C++ program broadcasts the tables with NCCL and the non-root rank reproduces the oracle (GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_bits_equal
from mono_dataset_code_b200 import synthetic as S

LIBDIR = os.path.join(ROOT, "mono_dataset_code_b200", "lib")
NCCL_LIB = os.path.join(LIBDIR, "libmdc_b200_nccl.so")
HAVE_NCCL = os.path.exists("/usr/include/nccl.h")


@pytest.mark.skipif(not HAVE_NCCL, reason="system NCCL headers absent")
def test_nccl_helper_builds_and_compiles_against_headers(tmp_path):
    assert os.path.exists(NCCL_LIB), "build() should have produced libmdc_b200_nccl.so"
    r = subprocess.run(["nm", "-D", "--defined-only", NCCL_LIB], capture_output=True, text=True)
    assert "mdc_ctx_create_broadcast" in r.stdout
    exe = compile_check(tmp_path)
    assert os.path.exists(exe)


def compile_check(tmp_path):
    exe = str(tmp_path / "nccl_bcast_check")
    nvcc_inc = "/usr/local/cuda/include"
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++11", "-O2", "-pthread",
           "-I" + os.path.join(ROOT, "include"), "-I" + nvcc_inc, os.path.join(ROOT, "tests", "cpp", "nccl_bcast_check.cpp"), "-o", exe,
           "-L" + LIBDIR, "-lmdc_b200_nccl", "-lmdc_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl",
           "-Wl,-rpath," + LIBDIR + ":/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_NCCL, reason="system NCCL headers absent")
def test_native_nccl_broadcast_two_gpus(port, tmp_path):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    iw, ih, ow, oh, n = 320, 240, 288, 200, 5
    d = tmp_path / "seq"
    files = S.write_dataset_dir(str(d), iw, ih, ow, oh, "crop")
    frames = np.stack([S.frame(i, iw, ih, ["uniform", "speckle", "gradient"][i % 3]) for i in range(n)])
    (tmp_path / "frames.bin").write_bytes(frames.tobytes())
    exe = compile_check(tmp_path)
    r = subprocess.run([exe, str(d) + "/", str(tmp_path / "frames.bin"), str(n), str(tmp_path / "out.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    out = np.fromfile(tmp_path / "out.bin", np.float32).reshape(n, ow * oh)
    f = port.fov_from_file(files["camera"])
    rx, ry = f.tables()
    ginv, _ = port.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
    _, vinv = port.vignette_maps(files["vignette_pixels"])
    for i in range(n):
        assert_bits_equal(out[i], port.get_image(rx, ry, iw, ih, ginv, vinv.reshape(-1), frames[i], 1, 1, 1, 1), f"rank-1 frame {i}")
