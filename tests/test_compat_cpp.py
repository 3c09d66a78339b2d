"""C++ drop-in surface (include/compat/*.h): compiles against the stand-in OpenCV/Eigen/libzip
headers (CPU), the reference's own main_playbackDataset.cpp links against it unchanged (CPU, only
where /root/reference exists), and on the GPU box the classes reproduce the oracle bit for bit."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_bits_equal
from mono_dataset_code_b200 import synthetic as S

LIBDIR = os.path.join(ROOT, "mono_dataset_code_b200", "lib")
CXXFLAGS = ["-std=c++0x", "-O2", "-DNDEBUG", "-I" + os.path.join(ROOT, "include", "compat"), "-I" + os.path.join(ROOT, "include"),
            "-I" + os.path.join(ROOT, "oracle", "shim")]
LDFLAGS = ["-L" + LIBDIR, "-lmdc_b200", "-Wl,-rpath," + LIBDIR]


def compile_cpp(src, exe):
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", *CXXFLAGS, src,
           os.path.join(ROOT, "oracle", "shim", "shim_impl.cpp"), "-o", exe, *LDFLAGS]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


@pytest.fixture(scope="module")
def compat_exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("compat")
    return compile_cpp(os.path.join(ROOT, "tests", "cpp", "compat_check.cpp"), str(d / "compat_check"))


def test_compat_headers_compile_and_link(compat_exe):
    assert os.path.exists(compat_exe)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/main_playbackDataset.cpp"), reason="reference sources not present")
@pytest.mark.parametrize("program", ["main_playbackDataset.cpp", "main_responseCalib.cpp"])
def test_reference_programs_link_unchanged(program, tmp_path):
    """playDataset and responseCalib are the link-compat consumers: their unmodified
    sources must compile and link against include/compat + libmdc_b200.so.  The file is copied next to nothing
    else so that its `#include "BenchmarkDatasetReader.h"` resolves to ours.  (vignetteCalib needs aruco.)"""
    src = tmp_path / program
    shutil.copy("/root/reference/src/" + program, src)
    compile_cpp(str(src), str(tmp_path / "prog"))


def write_sequence(d, iw, ih, ow, oh, n):
    files = S.write_dataset_dir(str(d), iw, ih, ow, oh, "crop", vignette_zeros=True)
    os.makedirs(d / "images")
    frames = []
    kinds = ["uniform", "speckle", "gradient"]
    with open(d / "times.txt", "w") as t:
        for i in range(n):
            fr = S.frame(i, iw, ih, kinds[i % 3])
            frames.append(fr)
            S.write_pgm(str(d / "images" / f"{i:05d}.pgm"), fr.reshape(ih, iw))
            t.write(f"{i} {1000.5 + i * 0.04:.6f} {1.5 + i}\n")
    return files, np.stack(frames)


@pytest.mark.gpu
def test_compat_classes_match_oracle_on_gpu(compat_exe, port, tmp_path):
    iw, ih, ow, oh, n = 320, 240, 288, 200, 3
    d = tmp_path / "seq"
    files, frames = write_sequence(d, iw, ih, ow, oh, n)
    out = tmp_path / "out.bin"
    r = subprocess.run([compat_exe, str(d) + "/", str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    buf = open(out, "rb").read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v

    def take_f32(count):
        nonlocal pos
        a = np.frombuffer(buf, np.float32, count, pos).copy()
        pos += 4 * count
        return a

    f = port.fov_from_file(files["camera"])
    rx, ry = f.tables()
    ginv, _ = port.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
    _, vinv = port.vignette_maps(files["vignette_pixels"])
    vinv = vinv.reshape(-1)
    assert take("<i")[0] == n
    for i in range(n):
        for flags in range(16):
            w, h = take("<ii")
            rectify, g, v, k = flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, (flags >> 3) & 1
            assert (w, h) == ((ow, oh) if rectify else (iw, ih))
            assert_bits_equal(take_f32(w * h), port.get_image(rx, ry, iw, ih, ginv, vinv, frames[i], rectify, g, v, k),
                              f"C++ getImage id={i} flags={flags}")
    assert_bits_equal(take_f32(ow * oh), port.undistort(rx, ry, iw, frames[0]), "C++ undistort<uchar>")
    fl = port.unmap(ginv, vinv, frames[0], 1, 1, 1)
    assert_bits_equal(take_f32(ow * oh), port.undistort(rx, ry, iw, fl), "C++ unMapImage+undistort<float>")
    assert (take_f32(ow * oh) == 5.0).all(), "wrong pixel count must leave the output untouched"
    K = take_f32(18)
    assert_bits_equal(K[:9].reshape(3, 3), f.K()[0], "getK_rect")
    assert_bits_equal(K[9:].reshape(3, 3), f.K()[1], "getK_org")
    assert take_f32(1)[0] == np.float32(S.TUM_CALIB[4])
    take_f32(5)
    assert_bits_equal(take_f32(256), ginv, "getGInv")
    ts, ex = take("<df")
    assert ts == 1000.54 and ex == 2.5
    xs, ys = take_f32(3), take_f32(3)
    ex_, ey_ = f.distort(np.array([0.0, 10.5, ow], np.float32), np.array([0.0, 20.25, oh], np.float32))
    assert_bits_equal(xs, ex_, "distortCoordinates x")
    assert_bits_equal(ys, ey_, "distortCoordinates y")


@pytest.mark.parametrize("zipped", [False, True], ids=["folder", "zip"])
def test_native_sequence_reader_mode_of_the_compat_header(tmp_path, zipped):
    """MDC_NATIVE_SEQUENCE_READER: DatasetReader without libzip / cv::imread — file order, times.txt, JPEG / PNG / PGM frames."""
    import zipfile
    cv2 = pytest.importorskip("cv2")
    exe = compile_cpp(os.path.join(ROOT, "tests", "cpp", "native_reader_check.cpp"), str(tmp_path / "native_reader_check"))
    iw, ih, n = 96, 64, 7
    d = tmp_path / "seq"
    S.write_dataset_dir(str(d), iw, ih, iw, ih, "crop")
    stage = d / ("stage" if zipped else "images")
    os.makedirs(stage)
    expect = []
    for i in range(n):
        fr = S.frame(i, iw, ih, ["uniform", "speckle", "gradient"][i % 3]).reshape(ih, iw)
        if i % 3 == 0:
            ok, enc = cv2.imencode(".jpg", fr, [cv2.IMWRITE_JPEG_QUALITY, 85])
            (stage / f"{i:05d}.jpg").write_bytes(enc.tobytes())
            expect.append(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
        elif i % 3 == 1:
            S.write_png_gray(str(stage / f"{i:05d}.png"), fr)
            expect.append(fr)
        else:
            S.write_pgm(str(stage / f"{i:05d}.pgm"), fr)
            expect.append(fr)
    if zipped:
        with zipfile.ZipFile(d / "images.zip", "w", zipfile.ZIP_DEFLATED) as z:
            for nm in sorted(os.listdir(stage), reverse=True):
                z.write(stage / nm, nm)
    with open(d / "times.txt", "w") as t:
        for i in range(n):
            t.write(f"{i} {5.25 + i:.6f} {0.75 * (i + 1):.4f}\n")
    out = tmp_path / "dump.bin"
    r = subprocess.run([exe, str(d) + "/", str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    blob = out.read_bytes()
    (count,) = struct.unpack_from("i", blob, 0)
    assert count == n
    pos = 4
    for i in range(n):
        stamp, exposure, rows, cols = struct.unpack_from("=dfii", blob, pos)
        pos += struct.calcsize("=dfii")
        assert stamp == float(f"{5.25 + i:.6f}") and exposure == np.float32(f"{0.75 * (i + 1):.4f}")
        assert (rows, cols) == (ih, iw)
        px = np.frombuffer(blob, np.uint8, rows * cols, pos).reshape(rows, cols)
        pos += rows * cols
        assert np.array_equal(px, expect[i]), f"frame {i}"
