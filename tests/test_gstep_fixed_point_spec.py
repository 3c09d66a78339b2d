"""tests/gstep_fixed_point_spec.py is what the CUDA G-step is compared with bit for bit on the GPU; here the specification itself is
pinned on the CPU: against the oracle's restatement of main_responseCalib.cpp:286-304 (the reference's sequential fp64 chain), against
math.fsum (the correctly rounded exact sum) and against itself when the pixels are split into slices the way ranks split them."""
import math

import numpy as np
import pytest

import gstep_fixed_point_spec as fx
from oracle.loader import PortOracle


@pytest.fixture(scope="module")
def port():
    return PortOracle()


def stack(seed, n=31, npix=2000):
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, (n, npix), dtype=np.uint8)
    data[:, 3:30] = 255
    data[:, 100:140] = 9
    t = rng.uniform(0.05, 20.0, n)
    E = np.exp(rng.uniform(-10.0, 5.0, npix))
    return data, t, E


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_spec_matches_the_reference_chain_to_rounding_and_the_exact_sum_closely(port, seed):
    data, t, E = stack(seed)
    t[4] = -0.7
    G = fx.gstep(data, t, E)
    G_ref = port.gstep(data, t, E)
    m = np.isfinite(G_ref)
    assert np.array_equal(np.isfinite(G), m)
    assert np.max(np.abs(G[m] - G_ref[m]) / np.abs(G_ref[m])) < 1e-10
    prod = E[None, :] * t[:, None]
    pmax = np.abs(E).max() * np.abs(t).max()
    for b in range(255):
        sel = data == b
        if sel.any():
            exact = math.fsum(prod[sel].tolist()) / int(sel.sum())
            assert abs(G[b] - exact) <= 4e-15 * pmax + 1e-15 * abs(exact)


def test_slices_add_up_to_the_whole_bit_for_bit():
    data, t, E = stack(7, n=19, npix=1500)
    whole = fx.gstep(data, t, E)
    cuts = [0, 384, 385, 1100, 1500]                     # ragged slices, one of a single pixel
    words = np.zeros(4, np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        words = np.maximum(words, fx.scale_words(E[a:b], t))
    limbs, special, counts = np.zeros(768, np.int64), np.zeros(256), np.zeros(256, np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        l, s, c = fx.accumulate(data[:, a:b], t, E[a:b], words)
        limbs += l; special += s; counts += c
    assert np.array_equal(fx.finish(words, limbs, special, counts).view(np.int64), whole.view(np.int64))


@pytest.mark.parametrize("case", ["nan_E", "inf_t", "zero_E", "huge_range", "tiny"])
def test_special_values_follow_the_reference(port, case):
    data, t, E = stack(11, n=13, npix=600)
    if case == "nan_E":
        E[50] = np.nan                                   # poisons exactly the bins pixel 50 has unsaturated samples in
    elif case == "inf_t":
        t[2] = np.inf                                    # no usable scale: every sample is range-checked on its own
    elif case == "zero_E":
        E[:] = 0.0
    elif case == "huge_range":
        E[::7] *= 1e200
        E[1::7] *= 1e-200
    else:
        E *= 1e-250                                      # (the scale exponent is clamped to +-1000: products below 2^-950 would lose bits)
    with np.errstate(invalid="ignore", over="ignore"):
        G, G_ref = fx.gstep(data, t, E), port.gstep(data, t, E)
    assert np.array_equal(np.isnan(G), np.isnan(G_ref))
    assert np.array_equal(np.isinf(G), np.isinf(G_ref))
    m = np.isfinite(G_ref)
    if case == "inf_t":
        # A non-finite exposure time leaves no common scale (include/mdc_b200.h): the bins it does not touch are summed at unit
        # resolution — each product rounded to an integer — which is what the kernels do, too.  Garbage in, coarse out; never wrong
        # about which bins are infinite.
        direct = np.array([(data == b).any() and not (data[2] == b).any() for b in range(255)] + [False])      # bins with samples, none from image 2
        assert direct.sum() > 20 and np.max(np.abs(G[direct] - G_ref[direct])) <= 0.5      # (the other finite entries are extrapolations, :300-304)
        return
    # one scale for all bins: a sample is resolved to 2^-48 of the largest product, so that is the error bound of a bin's mean — bins
    # whose products are all 2^48 times smaller than the largest one (only in "huge_range") come out as 0
    fin = np.isfinite(E)
    pmax = np.abs(E[fin]).max() * np.abs(t).max()
    assert np.all(np.abs(G[m] - G_ref[m]) <= 4e-15 * pmax + 1e-10 * np.abs(G_ref[m]))
    if case == "zero_E":
        assert np.all(G[m] == 0.0)
