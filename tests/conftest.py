"""pytest configuration: markers, shared fixtures, oracle/product loaders."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Calibration configurations used across the parity tests: name -> (in_w, in_h, out_w, out_h, mode, calib)
from mono_dataset_code_b200 import synthetic as S  # noqa: E402

CALIBS = {
    "c1_crop_640": (640, 480, 640, 480, "crop", S.TUM_CALIB),
    "tum_explicit": (1280, 1024, 640, 480, "0.4 0.53 0.5 0.5 0", S.TUM_CALIB),
    "full_blackpx": (640, 480, 640, 480, "full", (0.349153, 0.436593, 0.493140, 0.499021, 0.6)),
    "full_wrap": (640, 480, 640, 480, "full", S.TUM_CALIB),          # tan() past pi/2: negative focal, mostly black
    "omega0": (320, 240, 200, 100, "crop", (0.5, 0.6, 0.5, 0.5, 0.0)),
    "odd_sizes": (333, 217, 301, 199, "crop", (0.41, 0.52, 0.47, 0.52, 0.8)),
    "upscale": (160, 120, 400, 300, "crop", S.TUM_CALIB),
}
BIG_CALIBS = {
    "c2_crop_1280": (1280, 1024, 1280, 1024, "crop", S.TUM_CALIB),
    "c4_crop_1920": (1920, 1080, 1920, 1080, "crop", S.TUM_CALIB),
}


@pytest.fixture(scope="session")
def port():
    from oracle import loader
    loader.build("port")
    return loader.PortOracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import loader
    if os.path.isdir("/root/reference/src"):
        loader.build("ref")
    if not loader.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    return loader.RefOracle()


@pytest.fixture(scope="session")
def ref_f():
    from oracle import loader
    if not loader.ref_available(True):
        pytest.skip("oracle/_ref float variant not built")
    return loader.RefOracle(float_math=True)


@pytest.fixture()
def dataset_dir(tmp_path):
    def make(name_or_cfg, **kw):
        cfg = {**CALIBS, **BIG_CALIBS}[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
        iw, ih, ow, oh, mode, calib = cfg
        d = tmp_path / ("ds_%d" % len(list(tmp_path.iterdir())))
        return S.write_dataset_dir(str(d), iw, ih, ow, oh, mode, calib, **kw)
    return make


def bits(a):
    """View float32/float64 arrays as integers for bit-exact comparison (NaN payloads included)."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        # NaNs may differ in sign/payload between x86 and other producers: compare NaN masks separately
        na, nb = np.isnan(a), np.isnan(b)
        assert np.array_equal(na, nb), f"{what}: NaN masks differ ({na.sum()} vs {nb.sum()})"
        bad = (ba != bb) & ~na
        idx = np.flatnonzero(bad)
        assert idx.size == 0, f"{what}: {idx.size} of {a.size} values differ, first at {idx[:5]}: {a.reshape(-1)[idx[:5]]} vs {b.reshape(-1)[idx[:5]]}"
