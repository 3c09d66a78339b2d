"""Pins the responseCalib restatements (oracle/port: E-init, leak padding, G-step, E-step, rescale, rmse) against the REFERENCE
PROGRAM itself: main_responseCalib.cpp compiled unmodified (oracle/_ref/responseCalib_ref, `make -C oracle ref`) is run on a small
PGM sequence; the G it writes to photoCalibResult/pcalib.txt (15 significant digits), the per-iteration rmse / sample count of
photoCalibResult/log.txt and the 16-bit irradiance plots it hands to cv::imwrite are compared with the same loop composed from the
restatement.  The GPU kernels are then compared with the restatement in tests/test_gpu_parity.py."""
import os
import subprocess

import numpy as np
import pytest

from mono_dataset_code_b200 import synthetic as S
from oracle.loader import PortOracle

EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "responseCalib_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/responseCalib_ref not built (needs /root/reference)")

W, H, N = 56, 40, 18


def write_sequence(d, rng):
    """An exposure sweep of a fixed scene through a gamma response, with noise and a saturating blob."""
    S.write_dataset_dir(str(d), W, H, W, H, "crop")
    os.makedirs(d / "images")
    scene = rng.uniform(4.0, 90.0, W * H)
    scene[5 * W + 7: 5 * W + 12] = 400.0                       # saturates at most exposures -> leak padding matters
    times = np.geomspace(0.2, 6.0, N)
    frames = []
    with open(d / "times.txt", "w") as t:
        for i, ti in enumerate(times):
            irr = np.clip(scene * ti, 0, 255.0)
            img = np.clip(np.rint(255.0 * (irr / 255.0) ** (1 / 2.2) + rng.normal(0, 1.0, W * H)), 0, 255).astype(np.uint8)
            frames.append(img)
            S.write_pgm(str(d / "images" / f"{i:05d}.pgm"), img.reshape(H, W))
            t.write(f"{i} {100.0 + 0.05 * i:.6f} {ti:.7f}\n")
    exposures = np.array([np.float32(f"{ti:.7f}") for ti in times], dtype=np.float32).astype(np.float64)   # parsed with %f into a float
    return np.stack(frames), exposures


def read_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    rows, cols, typ = np.frombuffer(raw[:12].tobytes(), dtype=np.int32)
    return rows, cols, typ, raw[12:]


@pytest.mark.parametrize("nits,leak", [(4, 2), (3, 0)])
def test_restated_loop_reproduces_the_reference_program(tmp_path, nits, leak):
    rng = np.random.default_rng(100 + nits)
    seq = tmp_path / "seq"
    seq.mkdir()
    frames, t = write_sequence(seq, rng)
    work = tmp_path / "work"
    (work / "dump").mkdir(parents=True)
    env = dict(os.environ, MDC_SHIM_DUMP_DIR=str(work / "dump"))
    r = subprocess.run([EXE, str(seq) + "/", f"iterations={nits}", f"leakPadding={leak}"], cwd=work, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert f"loaded {N} images" in r.stdout
    G_ref = np.array(open(work / "photoCalibResult" / "pcalib.txt").read().split(), dtype=np.float64)
    log_ref = np.loadtxt(work / "photoCalibResult" / "log.txt", ndmin=2)          # it, n, num, rmse
    assert G_ref.shape == (256,) and log_ref.shape == (nits, 4)

    port = PortOracle()
    data = np.stack([port.leak_padding(f, W, H, leak) for f in frames])
    E = port.einit(data)
    G = np.zeros(256)
    log = []
    for it in range(nits):
        G = port.gstep(data, t, E)
        E = port.estep(data, t, G)
        port.rescale(E, G)
        r_ = port.rmse(data, t, G, E)
        log.append([it, N, r_[1], r_[0]])
    log = np.array(log)
    fin = np.isfinite(G_ref)
    assert np.array_equal(np.isfinite(G), fin)
    assert np.max(np.abs(G[fin] - G_ref[fin]) / np.maximum(np.abs(G_ref[fin]), 1e-300)) < 5e-14      # 15 printed digits
    assert np.array_equal(log[:, 2], log_ref[:, 2])                                                  # sample counts: exact
    assert np.max(np.abs(log[:, 3] - log_ref[:, 3]) / log_ref[:, 3]) < 5e-14                         # rmse after rescale
    # last irradiance plot: E after the last E-step, before the rescale
    rows, cols, typ, px = read_dump(work / "dump" / f"E-{nits}16.png.raw")
    assert (rows, cols, typ) == (H, W, 2)
    e16 = px.view(np.uint16).astype(np.float64)
    G_last = port.gstep(data, t, E_prev(port, data, t, nits))      # E/G as they were when the plot was made
    E_plot = port.estep(data, t, G_last)
    lo, hi = np.nanmin(E_plot), np.nanmax(E_plot)
    expect = np.floor(255.0 * 255.0 * (E_plot - lo) / (hi - lo))
    ok = np.isfinite(E_plot)
    assert np.max(np.abs(e16[ok] - expect[ok])) <= 1.0, "irradiance plot differs by more than one 16-bit step"


def E_prev(port, data, t, nits):
    """E at the start of the last iteration (restatement)."""
    E = port.einit(data)
    for _ in range(nits - 1):
        G = port.gstep(data, t, E)
        E = port.estep(data, t, G)
        port.rescale(E, G)
    return E
