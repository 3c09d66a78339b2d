"""csrc/mdc_atanf.h restates glibc's atanf so that distortCoordinates can run on the GPU with the host's bits
(SURVEY.md §8f N3).  CPU part: the restatement, compiled for the host, against the platform's libm."""
import ctypes
import ctypes.util

import numpy as np
import pytest

from mono_dataset_code_b200 import api
from conftest import assert_bits_equal


def libm_atanf(x: np.ndarray) -> np.ndarray:
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.atanf.restype = ctypes.c_float
    libm.atanf.argtypes = [ctypes.c_float]
    return np.array([libm.atanf(float(v)) for v in x], dtype=np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def interesting_inputs():
    edges = [0x31000000, 0x3ee00000, 0x3f300000, 0x3f980000, 0x401c0000, 0x4c000000, 0x7f800000]   # branch points of the algorithm
    u = [0, 1, 0x007fffff, 0x00800000, 0x7f7fffff, 0x7f800000]
    for e in edges:
        u += [e - 2, e - 1, e, e + 1, e + 2]
    u = np.array(u, dtype=np.uint32)
    return np.concatenate([u, u | np.uint32(0x80000000)]).view(np.float32)


def test_restated_atanf_matches_libm_on_branch_points_and_samples():
    rng = np.random.default_rng(7)
    x = np.concatenate([
        interesting_inputs(),
        rng.integers(0, 0x7f800000, 60000, dtype=np.uint32).view(np.float32),                          # any magnitude
        (rng.integers(0, 0x7f800000, 20000, dtype=np.uint32) | np.uint32(0x80000000)).view(np.float32),
        rng.uniform(0, 4, 60000).astype(np.float32),                                                   # where distortCoordinates lives
    ])
    assert_bits_equal(api.atanf_host(x), libm_atanf(x), "restated atanf vs libm")


def test_restated_atanf_nan_and_signs():
    x = np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0], dtype=np.float32)
    got = api.atanf_host(x)
    assert np.isnan(got[0]) and np.isnan(got[1])
    assert got[2] == np.float32(np.pi / 2) and got[3] == -np.float32(np.pi / 2)
    assert bits(got[4:5])[0] == 0 and bits(got[5:6])[0] == 0x80000000


@pytest.mark.gpu
def test_device_atanf_matches_host_libm_bitwise():
    import torch
    rng = np.random.default_rng(8)
    x = np.concatenate([
        interesting_inputs(),
        np.arange(0, 0x7f800000, 4099, dtype=np.uint32).view(np.float32),                 # strided sweep over every binade
        (np.arange(0, 0x7f800000, 8191, dtype=np.uint32) | np.uint32(0x80000000)).view(np.float32),
        rng.uniform(0, 4, 200000).astype(np.float32),
    ])
    dev = api.atanf_device(torch.from_numpy(x).cuda()).cpu().numpy()
    host = api.atanf_host(x)
    assert_bits_equal(dev, host, "device vs host evaluation of the restatement")      # (NaN payloads aside)
    sample = np.concatenate([x[:80], x[80::97]])
    assert_bits_equal(api.atanf_host(sample), libm_atanf(sample), "restatement vs this box's libm")
