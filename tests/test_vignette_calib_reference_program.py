"""Pins the vignetteCalib restatements (oracle/port: oport_vc_*) against the REFERENCE PROGRAM: main_vignetteCalib.cpp compiled
unmodified (oracle/_ref/vignetteCalib_ref, `make -C oracle ref`) against stand-in aruco / OpenCV / Eigen headers.  The stand-ins
take the place of the marker detector (one marker per image) and of cv::findHomography (homographies read from a file); everything
else — the reader, getImage, distortCoordinates, the image preparation and the optimisation loop — is the reference's own code.
The program is run on a small PGM sequence; what it writes (vignetteCalibResult/log.txt and the images it hands to cv::imwrite)
is compared with a replay composed from the restatement.  The GPU kernels are compared with the restatement in
tests/test_vignette_calib.py."""
import os
import subprocess

import numpy as np
import pytest

from mono_dataset_code_b200 import synthetic as S
from oracle.loader import PortOracle, RefOracle, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "vignetteCalib_ref")
pytestmark = pytest.mark.skipif(not (os.path.exists(EXE) and ref_available()), reason="oracle/_ref not built (needs /root/reference)")

WI, HI, GW, GH, N, ITS = 160, 120, 200, 200, 9, 6          # GW*GH*4 bytes > glibc's mmap threshold: the program's
                                                            # uninitialised planeColor (new float[], :381) is then zero pages


def make_sequence(d, rng):
    files = S.write_dataset_dir(str(d), WI, HI, WI, HI, "crop")
    os.replace(files["vignette_pgm"], files["vignette"])          # the stand-in imread only understands PGM (by content)
    os.makedirs(d / "images")
    yy, xx = np.mgrid[0:HI, 0:WI].astype(np.float64)
    vig = 1.0 - 0.5 * ((xx - WI / 2) ** 2 + (yy - HI / 2) ** 2) / ((WI / 2) ** 2 + (HI / 2) ** 2)
    frames, exposures, Hs = [], [], []
    with open(d / "times.txt", "w") as t:
        for i in range(N):
            scene = 120 + 50 * np.sin(xx * 0.11 + 0.4 * i) * np.cos(yy * 0.09 - 0.3 * i)
            expo = 0.8 + 0.15 * i
            img = np.clip(np.rint(scene * vig * (0.7 + 0.05 * i) + rng.normal(0, 1.5, scene.shape)), 0, 255).astype(np.uint8)
            img[40 + i, 30:34] = 255 if i % 3 == 0 else img[40 + i, 30:34]      # a few hard edges
            frames.append(img.ravel())
            exposures.append(np.float32(f"{expo:.6f}"))
            S.write_pgm(str(d / "images" / f"{i:05d}.pgm"), img)
            t.write(f"{i} {10.0 + 0.1 * i:.6f} {expo:.6f}\n")
            ang, s = 0.25 * np.sin(i), 14.0 + 1.5 * i                           # plane units -> rectified pixels
            Hs.append([s * np.cos(ang), -s * np.sin(ang), WI / 2 + 6 * np.cos(i), s * np.sin(ang), s * np.cos(ang), HI / 2 + 5 * np.sin(i),
                       0.004 * np.cos(i), -0.003 * np.sin(i), 1.0])
    np.savetxt(d / "homographies.txt", np.array(Hs), fmt="%.17g")
    return files, np.stack(frames), np.array(exposures, np.float32), np.array(Hs)


def read_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    rows, cols, typ = np.frombuffer(raw[:12].tobytes(), dtype=np.int32)
    return int(rows), int(cols), int(typ), raw[12:]


def to_u16(v):
    """cv::Mat(float) * 254.9 * 254.9 -> convertTo(CV_16U), as the stand-in evaluates it."""
    a = (v.astype(np.float64) * 254.9).astype(np.float32)
    a = (a.astype(np.float64) * 254.9).astype(np.float32).astype(np.float64)
    r = np.rint(a)
    return np.where(np.isnan(r), 0, np.clip(r, 0, 65535)).astype(np.uint16)


def test_replay_of_the_reference_program(tmp_path):
    rng = np.random.default_rng(77)
    seq = tmp_path / "seq"
    seq.mkdir()
    files, frames, expo, Hs = make_sequence(seq, rng)
    work = tmp_path / "work"
    (work / "dump").mkdir(parents=True)
    env = dict(os.environ, MDC_SHIM_DUMP_DIR=str(work / "dump"), MDC_SHIM_HOMOGRAPHIES=str(seq / "homographies.txt"))
    r = subprocess.run([EXE, str(seq) + "/", f"iterations={ITS}", f"patternX={GW}", f"patternY={GH}"], cwd=work, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    log_ref = np.loadtxt(work / "vignetteCalibResult" / "log.txt", ndmin=2)            # it, n, R, sqrtf(E/R) of the vignette step
    assert log_ref.shape == (ITS, 4) and np.all(log_ref[:, 1] == N)

    # ---- replay with the restatement (+ the reference's own distortCoordinates / unMapImage through oracle/_ref)
    port, ref = PortOracle(), RefOracle()
    ref.register_image(files["vignette"], files["vignette_pixels"])
    fov = ref.fov(files["camera"])
    photo = ref.photo(files["pcalib"], files["vignette"], WI, HI)
    mean_exposure = np.float32(0)
    for e in expo:
        mean_exposure = np.float32(mean_exposure + e)                                  # :225-227
    mean_exposure = np.float32(mean_exposure / np.float32(N))
    images, p2x, p2y = [], [], []
    for i in range(N):
        X, Y = port.vc_plane_maps(Hs[i], GW, GH)
        X, Y = fov.distort(X, Y)                                                       # :284
        raw = photo.unmap(frames[i], True, False, False)                               # getImage(i, false, true, false, false), :264
        images.append(port.vc_prepare_image(raw, WI, HI, mean_exposure, expo[i]))
        port.vc_mask_maps(X, Y, WI, HI)
        p2x.append(X)
        p2y.append(Y)
    images, p2x, p2y = np.stack(images), np.stack(p2x), np.stack(p2y)
    assert np.isfinite(p2x).mean() > 0.3
    plane = np.zeros(GW * GH, np.float32)
    vig = np.ones(WI * HI, np.float32)
    log = []
    for it in range(ITS):
        oth2 = 15 * 15 if it >= ITS // 2 else 10000 * 10000
        plane, _, _, _ = port.vc_plane_step(images, p2x, p2y, WI, HI, vig, plane, oth2)
        vig, _, _, vs = port.vc_vignette_step(images, p2x, p2y, WI, HI, plane, vig, oth2)
        log.append([it, N, vs[1], np.sqrt(np.float32(vs[0] / vs[1]))])
    log = np.array(log, dtype=np.float64)
    assert np.array_equal(log[:, 2], log_ref[:, 2]), (log[:, 2], log_ref[:, 2])         # residual-term counts: exact
    assert np.max(np.abs(log[:, 3] - log_ref[:, 3]) / log_ref[:, 3]) < 1e-6             # sqrtf(E/R)

    rows, cols, typ, px = read_dump(work / "dump" / "vignette.png.raw")
    assert (rows, cols, typ) == (HI, WI, 2)
    assert np.array_equal(px.view(np.uint16), to_u16(vig)), "vignette.png of the last iteration"
    rows, cols, typ, px = read_dump(work / "dump" / "vignetteSmoothed.png.raw")
    assert np.array_equal(px.view(np.uint16), to_u16(port.vc_smooth(vig, WI, HI, 4))), "vignetteSmoothed.png"
    # plane.png (displayImage, :72-92): 8UC3, NaN -> (0,0,255), else 255*(I-vmin)/(vmax-vmin) truncated
    rows, cols, typ, px = read_dump(work / "dump" / "plane.png.raw")
    assert (rows, cols, typ) == (GH, GW, 16)
    rgb = px.reshape(-1, 3)
    fin = np.isfinite(plane)
    vmin, vmax = np.float32(plane[fin].min()), np.float32(plane[fin].max())
    c = (np.float32(255) * (plane - vmin) / (vmax - vmin))
    assert np.array_equal(rgb[fin, 0], c[fin].astype(np.uint8)) and np.array_equal(rgb[~fin], np.tile(np.array([0, 0, 255], np.uint8), ((~fin).sum(), 1)))
