"""tools/responseCalib_b200.cpp — the C++ host of the GPU calibrator: builds and links against include/ + libmdc_b200.so (CPU), and on a
GPU reproduces what the REFERENCE PROGRAM wrote for the same sequence (tests/golden/programs/response_calib.npz, produced by
main_responseCalib.cpp compiled unmodified): pcalib.txt and log.txt."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from mono_dataset_code_b200 import synthetic as S

LIBDIR = os.path.join(ROOT, "mono_dataset_code_b200", "lib")


def build_tool(tmp_path):
    exe = str(tmp_path / "responseCalib_b200")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++11", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", os.path.join(ROOT, "tools", "responseCalib_b200.cpp"), "-o", exe, "-L" + LIBDIR, "-lmdc_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + LIBDIR + ":/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_calibrator_host_program_builds(tmp_path):
    exe = build_tool(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "usage: responseCalib_b200" in r.stdout


@pytest.mark.gpu
def test_calibrator_host_program_reproduces_the_reference_program(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "programs", "response_calib.npz"))
    w, h, nits, leak = int(g["w"]), int(g["h"]), int(g["nits"]), int(g["leak_padding"])
    seq = tmp_path / "seq"
    os.makedirs(seq / "images")
    with open(seq / "times.txt", "w") as t:
        for i, (f, e) in enumerate(zip(g["frames"], g["exposures"])):
            S.write_pgm(str(seq / "images" / f"{i:05d}.pgm"), f.reshape(h, w))
            t.write(f"{i} {100.0 + 0.05 * i:.6f} {float(e):.9g}\n")
    exe = build_tool(tmp_path)
    work = tmp_path / "work"
    work.mkdir()
    r = subprocess.run([exe, str(seq) + "/", f"iterations={nits}", f"leakPadding={leak}"], cwd=work, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert f"loaded {len(g['frames'])} images" in r.stdout and "init RMSE" in r.stdout and "resc RMSE" in r.stdout
    G = np.array(open(work / "photoCalibResult" / "pcalib.txt").read().split(), dtype=np.float64)
    log = np.loadtxt(work / "photoCalibResult" / "log.txt", ndmin=2)          # it, n, num, rmse
    ref = g["G"]
    fin = np.isfinite(ref)
    assert G.shape == (256,) and np.array_equal(np.isfinite(G), fin)
    assert np.max(np.abs(G[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-300)) < 1e-9
    assert log.shape == (nits, 4) and np.array_equal(log[:, 2], g["log_num"])
    assert np.max(np.abs(log[:, 3] - g["log_rmse"]) / g["log_rmse"]) < 1e-9
