"""CPU: product host code (table builders, file parsers, PNG/PGM decode, C-ABI surface) against the
C restatement and — where /root/reference or a prebuilt oracle/_ref exists — the reference itself."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import CALIBS, BIG_CALIBS, ROOT, assert_bits_equal
from mono_dataset_code_b200 import api, _lib, synthetic as S


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "mdc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mdc_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 40
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert names == set(_lib.SIGNATURES), (names ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in _lib.lib.mdc_version()


@pytest.mark.parametrize("name", list(CALIBS) + list(BIG_CALIBS))
@pytest.mark.parametrize("float_math", [False, True])
def test_fov_model_vs_port(name, float_math, port, dataset_dir):
    files = dataset_dir(name)
    u = api.UndistorterFOV(files["camera"], float_math=float_math)
    p = port.fov_from_file(files["camera"], float_math)
    assert u.isValid()
    assert u.getInputDims() == (p.in_w, p.in_h) and u.getOutputDims() == (p.out_w, p.out_h)
    rx, ry = u.remap_tables()
    px, py = p.tables()
    assert_bits_equal(rx, px, "remapX")
    assert_bits_equal(ry, py, "remapY")
    assert_bits_equal(u.getK_rect(), p.K()[0], "Krect")
    assert_bits_equal(u.getK_org(), p.K()[1], "Korg")
    rng = np.random.default_rng(9)
    x = rng.uniform(-30, p.out_w + 30, 2000).astype(np.float32)
    y = rng.uniform(-30, p.out_h + 30, 2000).astype(np.float32)
    ex, ey = p.distort(x, y)
    u.distortCoordinates(x, y)
    assert_bits_equal(x, ex, "distort x")
    assert_bits_equal(y, ey, "distort y")


@pytest.mark.parametrize("name", ["c1_crop_640", "full_blackpx", "odd_sizes"])
def test_fov_model_vs_reference(name, ref, dataset_dir):
    files = dataset_dir(name)
    u = api.UndistorterFOV(files["camera"])
    r = ref.fov(files["camera"])
    rx, ry = r.tables()
    assert_bits_equal(u.remap_tables()[0], rx, "remapX")
    assert_bits_equal(u.remap_tables()[1], ry, "remapY")
    assert_bits_equal(u.getK_rect(), r.K()[0], "Krect")
    assert u.getOmega() == r.omega()
    assert_bits_equal(u.getOriginalCalibration(), r.original_calibration(), "getOriginalCalibration")


def test_fov_from_params_equals_file(dataset_dir):
    files = dataset_dir("tum_explicit")
    a = api.UndistorterFOV(files["camera"])
    calib = np.array([float(v) for v in open(files["camera"]).readline().split()], np.float32)
    b = api.UndistorterFOV(params=(calib, 1280, 1024, _lib.FOV_EXPLICIT, [0.4, 0.53, 0.5, 0.5, 0], 640, 480))
    assert_bits_equal(a.remap_tables()[0], b.remap_tables()[0], "remapX")
    c = api.UndistorterFOV(params=(calib, 1280, 1024, _lib.FOV_CROP, None, 640, 480))
    assert c.isValid() and not np.array_equal(c.remap_tables()[0], a.remap_tables()[0])


def test_invalid_camera_files(tmp_path, capfd):
    cam = tmp_path / "camera.txt"
    for text in [S.camera_txt(640, 480, 640, 480, "none"), "1 2 3\n640 480\ncrop\n640 480\n",
                 S.camera_txt(640, 480, 640, 480, "crop ").replace("crop \n", "crop \n"),   # trailing blank: not 'crop'
                 "0.3 0.4 0.5 0.5 0.9\n640 480\ncrop\nxx\n"]:
        cam.write_text(text)
        u = api.UndistorterFOV(str(cam))
        assert not u.isValid(), text
        assert u.remap_tables() == (None, None)
        out = np.full(10, 3.0, np.float32)
        u.undistort(np.zeros(10, np.uint8), out)      # invalid object: output untouched, no CUDA needed
        assert (out == 3.0).all()
    u = api.UndistorterFOV(str(tmp_path / "missing.txt"))
    assert not u.isValid() and u.status == 2
    assert "Failed to read camera calibration" in capfd.readouterr().out
    # sizes from the file drive allocations: absurd ones are refused instead of letting an allocation throw across the C ABI
    for text in ["0.3 0.4 0.5 0.5 0.9\n640 480\ncrop\n2000000000 2000000000\n", "0.3 0.4 0.5 0.5 0.9\n2000000000 2000000000\ncrop\n640 480\n",
                 "0.3 0.4 0.5 0.5 0.9\n640 480\ncrop\n65536 65536\n"]:
        cam.write_text(text)
        u = api.UndistorterFOV(str(cam))
        assert not u.isValid() and u.status == 3, text      # MDC_ERR_FORMAT


@pytest.mark.parametrize("depth", [8, 16])
def test_photo_model_vs_port_and_cv2(depth, port, dataset_dir):
    files = dataset_dir("odd_sizes", vignette_depth=depth, vignette_zeros=True)
    iw, ih = 333, 217
    p = api.PhotometricUndistorter(files["pcalib"], files["vignette"], iw, ih)
    assert p.validGamma and p.validVignette and p.status == 0
    ginv, g = port.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
    vm, vmi = port.vignette_maps(files["vignette_pixels"])
    assert_bits_equal(p.getGInv(), ginv, "GInv")
    assert_bits_equal(p.getG(), g, "G")
    assert_bits_equal(p.vignette_maps()[0], vm.reshape(-1), "vignetteMap")
    assert_bits_equal(p.vignette_maps()[1], vmi.reshape(-1), "vignetteMapInv")
    assert np.isinf(p.vignette_maps()[1]).sum() == 4
    # the PNG decoder against the real OpenCV decoder, and the PGM path
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.imread(files["vignette"], cv2.IMREAD_UNCHANGED), files["vignette_pixels"])
    q = api.PhotometricUndistorter(files["pcalib"], files["vignette_pgm"], iw, ih)
    assert_bits_equal(q.vignette_maps()[1], vmi.reshape(-1), "vignetteMapInv (pgm)")
    # from arrays
    a = api.PhotometricUndistorter("", "", iw, ih, arrays=(np.loadtxt(files["pcalib"], dtype=np.float32), files["vignette_pixels"]))
    assert_bits_equal(a.vignette_maps()[1], vmi.reshape(-1), "vignetteMapInv (arrays)")


def test_png_decoder_filters_against_cv2(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for dt in (np.uint8, np.uint16):
        img = (rng.integers(0, 60000 if dt == np.uint16 else 250, (37, 53)) + 1).astype(dt)
        img[5:20, 5:30] = np.arange(25, dtype=dt)[None, :] * 3 + 7      # smooth region -> Sub/Up/Paeth filters get chosen
        path = str(tmp_path / f"v_{np.dtype(dt).name}.png")
        cv2.imwrite(path, img)                                            # libpng picks per-row filters
        p = api.PhotometricUndistorter("", "", 53, 37, arrays=(S.ginv_raw().astype(np.float32), img))
        q = api.PhotometricUndistorter(str(tmp_path / "pc.txt"), path, 53, 37) if False else None
        pc = tmp_path / "pc.txt"
        pc.write_text(S.pcalib_txt())
        q = api.PhotometricUndistorter(str(pc), path, 53, 37)
        assert q.validVignette
        assert_bits_equal(q.vignette_maps()[0], p.vignette_maps()[0], "decoded vignette")


def test_invalid_photo_files(tmp_path, dataset_dir, capfd):
    files = dataset_dir("c1_crop_640")
    # empty names
    p = api.PhotometricUndistorter("", "", 640, 480)
    assert not p.validGamma and not p.validVignette and p.getGInv() is None and p.getG() is None
    # wrong vignette size: gamma stays valid
    p = api.PhotometricUndistorter(files["pcalib"], files["vignette"], 320, 240)
    assert p.validGamma and not p.validVignette and p.getGInv() is not None
    assert "Invalid vignette image size" in capfd.readouterr().out
    # missing vignette file behaves like an empty image
    p = api.PhotometricUndistorter(files["pcalib"], str(tmp_path / "nope.png"), 640, 480)
    assert p.validGamma and not p.validVignette
    # 255 entries / non-monotone
    bad = tmp_path / "bad.txt"
    bad.write_text(" ".join("%g" % v for v in S.ginv_raw()[:255]) + "\n")
    assert not api.PhotometricUndistorter(str(bad), files["vignette"], 640, 480).validGamma
    v = S.ginv_raw()
    v[10] = v[9]
    bad.write_text(S.pcalib_txt(v))
    assert not api.PhotometricUndistorter(str(bad), files["vignette"], 640, 480).validGamma
    # RGB png is refused (vignette invalid), not mis-decoded
    cv2 = pytest.importorskip("cv2")
    rgb = str(tmp_path / "rgb.png")
    cv2.imwrite(rgb, np.zeros((480, 640, 3), np.uint8))
    assert not api.PhotometricUndistorter(files["pcalib"], rgb, 640, 480).validVignette


def test_argument_validation_of_the_device_entry_points_needs_no_gpu(dataset_dir, capfd):
    """Bad arguments are rejected before any CUDA call, with the C ABI's status codes (include/mdc_b200.h:28-35)."""
    import ctypes as C
    from mono_dataset_code_b200 import _lib
    L = _lib.lib
    null = C.c_void_p()
    st = (C.c_double * 2)()
    # no context
    assert L.mdc_vc_plane_step(null, null, null, null, 1, 4, 4, 8, 8, null, null, 1.0, 1, st) == 1
    assert L.mdc_vc_vignette_step(null, null, null, null, 1, 4, 4, 8, 8, null, null, 1.0, 1, st) == 1
    assert L.mdc_vc_smooth(null, null, 8, 8, 4, null) == 1
    assert L.mdc_vignette_calib(null, null, null, null, 1, 4, 4, 8, 8, 2, 15, 1, null, null, null, None) == 1
    assert b"bad argument" in L.mdc_last_error()
    assert L.mdc_estep(null, null, 1, 16, null, null, null, null) == 1
    assert L.mdc_rc_gstep(null, null, 1, 16, null, null, null, null) == 1
    # distortCoordinates on the device: NULL model / NULL arrays are argument errors, an invalid model prints like the reference
    assert L.mdc_fov_distort_coordinates_device(null, null, null, 0, 0, null) == 1
    files = dataset_dir("c1_crop_640")
    u = api.UndistorterFOV(files["camera"])
    assert L.mdc_fov_distort_coordinates_device(u._h, null, null, 5, 0, null) == 1
    bad = C.c_void_p()
    L.mdc_fov_create(b"/nonexistent/camera.txt", C.byref(bad))
    capfd.readouterr()
    assert L.mdc_fov_distort_coordinates_device(bad, C.c_void_p(16), C.c_void_p(16), 5, 0, null) == 4       # MDC_ERR_INVALID_OBJECT
    assert "ERROR: invalid UndistorterFOV!" in capfd.readouterr().out
    L.mdc_fov_destroy(bad)
    assert L.mdc_atanf_device(null, null, 3, 0, null) == 1


def test_public_header_is_plain_c(tmp_path):
    """include/mdc_b200.h is the FFI boundary: it must compile as C99 (no C++ types, no torch types) and link from a C program."""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include "mdc_b200.h"\n#include <stdio.h>\nint main(void) { mdc_fov* f = 0; printf("%s %d\\n", mdc_version(), mdc_fov_is_valid(f)); return 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "mono_dataset_code_b200", "lib")
    exe = tmp_path / "c_abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"), str(src), "-o", str(exe),
                        "-L" + libdir, "-lmdc_b200", "-Wl,-rpath," + libdir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("mdc_b200")
