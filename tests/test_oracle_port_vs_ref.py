"""Pins the plain-C restatement (oracle/port) against the reference's own code
compiled unmodified (oracle/_ref): tables, getters and per-frame outputs must be
bit-identical on every fixture.  CPU only."""
import numpy as np
import pytest

from conftest import CALIBS, BIG_CALIBS, assert_bits_equal
from mono_dataset_code_b200 import synthetic as S


@pytest.mark.parametrize("name", list(CALIBS) + list(BIG_CALIBS))
@pytest.mark.parametrize("float_math", [False, True])
def test_fov_tables_and_K(name, float_math, port, ref, ref_f, dataset_dir):
    files = dataset_dir(name)
    r = (ref_f if float_math else ref).fov(files["camera"])
    p = port.fov_from_file(files["camera"], float_math)
    assert r.valid and p is not None
    assert (r.in_w, r.in_h, r.out_w, r.out_h) == (p.in_w, p.in_h, p.out_w, p.out_h)
    rx, ry = r.tables()
    px, py = p.tables()
    assert_bits_equal(rx, px, "remapX")
    assert_bits_equal(ry, py, "remapY")
    kr, ko = r.K()
    pkr, pko = p.K()
    assert_bits_equal(kr, pkr, "Krect")
    assert_bits_equal(ko, pko, "Korg")
    # distortCoordinates on off-grid points
    rng = np.random.default_rng(5)
    x = rng.uniform(-20, r.out_w + 20, 4096).astype(np.float32)
    y = rng.uniform(-20, r.out_h + 20, 4096).astype(np.float32)
    x[0], y[0] = kr[0, 2], kr[1, 2]  # r == 0 branch if representable
    a = r.distort(x, y)
    b = p.distort(x, y)
    assert_bits_equal(a[0], b[0], "distort x")
    assert_bits_equal(a[1], b[1], "distort y")


@pytest.mark.parametrize("name", ["c1_crop_640", "full_blackpx", "odd_sizes", "upscale", "tum_explicit"])
def test_undistort_and_unmap(name, port, ref, dataset_dir):
    files = dataset_dir(name, vignette_zeros=(name == "odd_sizes"))
    r = ref.fov(files["camera"])
    ref.register_image(files["vignette"], files["vignette_pixels"])
    rp = ref.photo(files["pcalib"], files["vignette"], r.in_w, r.in_h)
    assert rp.valid_gamma and rp.valid_vignette
    rx, ry = r.tables()
    ginv, g = port.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
    assert_bits_equal(rp.ginv(), ginv, "GInv")
    assert_bits_equal(rp.g()[[0, 255]], g[[0, 255]], "G ends")
    # forward G: compare only entries the reference's search defines (it leaves the others uninitialised)
    assert_bits_equal(rp.g()[1:255], g[1:255], "G")
    vm, vmi = port.vignette_maps(files["vignette_pixels"])
    rvm, rvmi = rp.vignette_maps()
    assert_bits_equal(rvm, vm.reshape(-1), "vignetteMap")
    assert_bits_equal(rvmi, vmi.reshape(-1), "vignetteMapInv")
    for kind in ["uniform", "speckle", "white", "black", "gradient"]:
        img = S.frame(3, r.in_w, r.in_h, kind)
        for gamma in (0, 1):
            for vig in (0, 1):
                for kill in (0, 1):
                    a = rp.unmap(img, gamma, vig, kill)
                    b = port.unmap(ginv, vmi, img, gamma, vig, kill)
                    assert_bits_equal(a, b, f"unmap {kind} {gamma}{vig}{kill}")
        fl = rp.unmap(img, 1, 1, 1)
        assert_bits_equal(r.undistort(fl), port.undistort(rx, ry, r.in_w, fl), f"undistort<float> {kind}")
        assert_bits_equal(r.undistort(img), port.undistort(rx, ry, r.in_w, img), f"undistort<uchar> {kind}")


def test_vignette_8bit_and_invalid_objects(port, ref, dataset_dir, tmp_path):
    files = dataset_dir("c1_crop_640", vignette_depth=8)
    ref.register_image(files["vignette"], files["vignette_pixels"])
    rp = ref.photo(files["pcalib"], files["vignette"], 640, 480)
    vm, vmi = port.vignette_maps(files["vignette_pixels"])
    assert_bits_equal(rp.vignette_maps()[1], vmi.reshape(-1), "vinv 8-bit")
    # wrong vignette size -> gamma valid, vignette invalid; unMapImage drops to gamma only
    rp2 = ref.photo(files["pcalib"], files["vignette"], 320, 240)
    assert rp2.valid_gamma and not rp2.valid_vignette
    img = S.frame(0, 320, 240)
    ginv, _ = port.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
    assert_bits_equal(rp2.unmap(img, 1, 1, 0), port.unmap(ginv, None, img, 1, 1, 0), "gamma-only fallback")
    # empty file name -> everything invalid, unMapImage is a plain u8->float cast
    rp3 = ref.photo("", "", 320, 240)
    assert not rp3.valid_gamma and rp3.ginv() is None
    assert_bits_equal(rp3.unmap(img, 1, 1, 1), port.unmap(None, None, img, 1, 1, 1), "invalid photo")
    # non-monotone pcalib -> invalid
    bad = tmp_path / "bad.txt"
    v = S.ginv_raw()
    v[100] = v[99]
    bad.write_text(S.pcalib_txt(v))
    assert not ref.photo(str(bad), files["vignette"], 640, 480).valid_gamma
    assert port.photo_tables(v.astype(np.float32))[0] is None
    # camera.txt with 'none' / garbage -> invalid undistorter, undistort leaves the output untouched
    cam = tmp_path / "cam.txt"
    cam.write_text(S.camera_txt(640, 480, 640, 480, "none"))
    rf = ref.fov(str(cam))
    assert not rf.valid and port.fov_from_file(str(cam)) is None
    cam.write_text("1 2 3\n")
    assert not ref.fov(str(cam)).valid and port.fov_from_file(str(cam)) is None
    assert not ref.fov(str(tmp_path / "missing.txt")).valid


def test_undistort_dim_mismatch_leaves_output(ref, dataset_dir):
    files = dataset_dir("c1_crop_640")
    r = ref.fov(files["camera"])
    img = S.frame(0, 640, 480)
    out = np.full(640 * 480, 7.0, np.float32)
    r.undistort(img, out, n_in=640 * 480 - 1)
    assert (out == 7.0).all()
    r.undistort(img, out, n_out=640 * 480 + 5)
    assert (out == 7.0).all()
