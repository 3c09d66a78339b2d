"""CPU: the C restatement (oracle/port) and the product's HOST table builders against the
committed golden fixtures, which were produced by the reference's own code
(tests/golden/make_golden.py).  Runs without /root/reference and without a GPU."""
import glob
import os

import numpy as np
import pytest

from conftest import assert_bits_equal
from mono_dataset_code_b200 import api

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_case(path, tmp_path):
    g = np.load(path)
    cam, pc = tmp_path / "camera.txt", tmp_path / "pcalib.txt"
    cam.write_bytes(g["camera_txt"].tobytes())
    pc.write_bytes(g["pcalib_txt"].tobytes())
    return g, str(cam), str(pc)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_matches_golden(path, port, tmp_path):
    g, cam, pc = load_case(path, tmp_path)
    iw, ih, ow, oh = [int(v) for v in g["dims"]]
    f = port.fov_from_file(cam)
    rx, ry = f.tables()
    assert_bits_equal(rx, g["remap_x"], "remapX")
    assert_bits_equal(ry, g["remap_y"], "remapY")
    assert_bits_equal(f.K()[0], g["k_rect"], "Krect")
    assert_bits_equal(f.K()[1], g["k_org"], "Korg")
    ginv, gf = port.photo_tables(np.loadtxt(pc, dtype=np.float32))
    assert_bits_equal(ginv, g["ginv"], "GInv")
    _, vinv = port.vignette_maps(g["vignette_pixels"])
    assert_bits_equal(vinv.reshape(-1), g["vinv"], "vignetteMapInv")
    for flags in range(16):
        rectify, gm, v, k = flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, (flags >> 3) & 1
        for i, fr in enumerate(g["frames"]):
            out = port.get_image(rx, ry, iw, ih, ginv, vinv.reshape(-1), fr, rectify, gm, v, k)
            assert_bits_equal(out, g[f"out_{flags:02d}"][i], f"getImage flags={flags} frame={i}")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_product_host_models_match_golden(path, tmp_path):
    """Product host code (csrc/mdc_host_models.cpp, csrc/mdc_gray_image.cpp): tables bit-identical."""
    from mono_dataset_code_b200 import synthetic as S
    g, cam, pc = load_case(path, tmp_path)
    iw, ih, ow, oh = [int(v) for v in g["dims"]]
    u = api.UndistorterFOV(cam)
    assert u.isValid() and u.getInputDims() == (iw, ih) and u.getOutputDims() == (ow, oh)
    rx, ry = u.remap_tables()
    assert_bits_equal(rx, g["remap_x"], "remapX")
    assert_bits_equal(ry, g["remap_y"], "remapY")
    assert_bits_equal(u.getK_rect(), g["k_rect"], "Krect")
    assert_bits_equal(u.getK_org(), g["k_org"], "Korg")
    vig = tmp_path / "vignette.png"
    S.write_png_gray(str(vig), g["vignette_pixels"])
    p = api.PhotometricUndistorter(pc, str(vig), iw, ih)
    assert p.validGamma and p.validVignette
    assert_bits_equal(p.getGInv(), g["ginv"], "GInv")
    defined = g["g"] == g["g"]
    assert_bits_equal(p.getG()[[0, 255]], g["g"][[0, 255]], "G ends")
    assert_bits_equal(p.vignette_maps()[1], g["vinv"], "vignetteMapInv")


def test_restated_response_calib_loop_reproduces_the_programs_golden_output(port):
    """tests/golden/programs/response_calib.npz = output of the reference's responseCalib program (make_golden_programs.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "programs", "response_calib.npz"))
    w, h, nits, leak = int(g["w"]), int(g["h"]), int(g["nits"]), int(g["leak_padding"])
    data = np.stack([port.leak_padding(f, w, h, leak) for f in g["frames"]])
    t = g["exposures"]
    E = port.einit(data)
    nums, rmses = [], []
    for _ in range(nits):
        G = port.gstep(data, t, E)
        E = port.estep(data, t, G)
        port.rescale(E, G)
        r = port.rmse(data, t, G, E)
        rmses.append(r[0]); nums.append(r[1])
    fin = np.isfinite(g["G"])
    assert np.array_equal(np.isfinite(G), fin)
    assert np.max(np.abs(G[fin] - g["G"][fin]) / np.maximum(np.abs(g["G"][fin]), 1e-300)) < 5e-14
    assert np.array_equal(np.array(nums), g["log_num"])
    assert np.max(np.abs(np.array(rmses) - g["log_rmse"]) / g["log_rmse"]) < 5e-14


def test_pyramid_restatement_is_pinned_by_an_independent_formulation(port):
    """Row P has no reference source; oracle/port's pyramid is compared here, bit for bit, with a second formulation written from
    DSO's makeImages (tests/pyramid_spec.py): odd sizes, NaN / inf pixels, denormal-range sums."""
    import pyramid_spec as spec
    rng = np.random.default_rng(3)
    for (w, h, levels) in [(64, 48, 5), (67, 45, 5), (1, 1, 3), (3, 2, 2), (130, 70, 7)]:
        img = rng.uniform(-50, 300, w * h).astype(np.float32)
        img[rng.integers(0, w * h, 5)] = np.nan
        img[rng.integers(0, w * h, 3)] = np.inf
        img[rng.integers(0, w * h, 3)] = np.float32(1e-42)
        got = port.pyramid(img, w, h, levels)
        exp = spec.pyramid(img, w, h, levels)
        assert len(got) == len(exp) == levels
        for l in range(levels):
            assert got[l].shape == exp[l].shape == ((w >> l) * (h >> l),)
            assert np.array_equal(np.isnan(got[l]), np.isnan(exp[l]))
            m = ~np.isnan(exp[l])
            assert np.array_equal(got[l][m].view(np.uint32), exp[l][m].view(np.uint32)), (w, h, l)
    # tap order matters: ((a+b)+c)+d differs from (a+c)+(b+d) in the last bit on suitable data, and the spec takes the former
    a, b, c, d = np.float32(1e8), np.float32(1.0), np.float32(-1e8), np.float32(1.0)
    blk = np.array([[a, b], [c, d]], np.float32)
    assert spec.pyr_down(blk)[0, 0] == np.float32(0.25) * (((a + b) + c) + d) != np.float32(0.25) * ((a + c) + (b + d))
