"""CPU property test: for random calibrations, sizes and output modes the product's host table builder
(csrc/mdc_host_models.cpp) and the C restatement (oracle/port) produce bit-identical remap tables, in both
tan()/sqrt() overload variants.  (Both are separately pinned to the reference's compiled code on the fixed
fixtures; this extends the agreement to the parameter space.)"""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import assert_bits_equal
from mono_dataset_code_b200 import api, _lib
from oracle import loader

PORT = None


def port():
    global PORT
    if PORT is None:
        loader.build("port")
        PORT = loader.PortOracle()
    return PORT


def f32(lo, hi):
    return st.floats(float(np.float32(lo)), float(np.float32(hi)), width=32, allow_nan=False, allow_infinity=False)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(fx=f32(0.2, 1.2), fy=f32(0.2, 1.2), cx=f32(0.3, 0.7), cy=f32(0.3, 0.7),
       omega=st.one_of(st.just(0.0), f32(0.05, 1.3)),
       in_w=st.integers(16, 200), in_h=st.integers(16, 160), out_w=st.integers(1, 180), out_h=st.integers(1, 150),
       mode=st.sampled_from([_lib.FOV_CROP, _lib.FOV_FULL, _lib.FOV_EXPLICIT]),
       ofx=f32(0.2, 1.0), ofy=f32(0.2, 1.0), ocx=f32(0.4, 0.6), ocy=f32(0.4, 0.6), float_math=st.booleans())
def test_random_calibrations_bit_identical_tables(fx, fy, cx, cy, omega, in_w, in_h, out_w, out_h, mode, ofx, ofy, ocx, ocy, float_math):
    in_calib = np.array([fx, fy, cx, cy, omega], np.float32)
    out_calib = np.array([ofx, ofy, ocx, ocy, 0], np.float32)
    u = api.UndistorterFOV(params=(in_calib, in_w, in_h, mode, out_calib, out_w, out_h), float_math=float_math)
    p = port().fov(in_calib, in_w, in_h, mode, out_calib, out_w, out_h, float_math)
    assert u.isValid()
    rx, ry = u.remap_tables()
    px, py = p.tables()
    assert_bits_equal(rx, px, "remapX")
    assert_bits_equal(ry, py, "remapY")
    kr, ko = p.K()
    assert_bits_equal(u.getK_rect(), kr, "Krect")
    assert_bits_equal(u.getK_org(), ko, "Korg")
    # every surviving entry keeps its 4 taps inside the input image (the guarantee the kernels rely on)
    ok = rx >= 0
    if ok.any():
        assert rx[ok].min() > 0 and ry[ok].min() > 0 and rx[ok].max() < in_w - 1 and ry[ok].max() < in_h - 1
    assert ((rx < 0) == (ry < 0)).all()
