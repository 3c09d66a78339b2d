"""csrc/mdc_jpeg.cpp: baseline JPEG -> 8-bit grey, compared byte for byte with the decoder behind cv::imread (the reference reads
its frames with cv::imread(..., CV_LOAD_IMAGE_GRAYSCALE), BenchmarkDatasetReader.h:252, :274)."""
import os

import numpy as np
import pytest

from mono_dataset_code_b200 import api

cv2 = pytest.importorskip("cv2")


def scene(rng, h, w, kind):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    if kind == "smooth":
        img = 128 + 90 * np.sin(xx * 0.07) * np.cos(yy * 0.05)
    elif kind == "noise":
        img = rng.integers(0, 256, (h, w)).astype(np.float64)
    elif kind == "edges":
        img = ((xx // 13 + yy // 9) % 2) * 255.0
    else:
        img = 40 + 0.6 * xx + rng.normal(0, 12, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def sequence_of(tmp_path, blobs):
    os.makedirs(tmp_path / "images", exist_ok=True)
    with open(tmp_path / "times.txt", "w") as t:
        for i, b in enumerate(blobs):
            (tmp_path / "images" / f"{i:05d}.jpg").write_bytes(b)
            t.write(f"{i} {i * 0.05:.3f} 1.0\n")
    return api.Sequence(str(tmp_path))


CASES = []
for kind in ("smooth", "noise", "edges", "ramp"):
    for quality in (35, 90, 100):
        CASES.append((kind, quality))


@pytest.mark.parametrize("size", [(64, 48), (83, 61), (8, 8), (1, 1), (17, 129), (640, 480)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_grey_jpeg_matches_opencv(tmp_path, size):
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    blobs, expect = [], []
    for kind, quality in CASES:
        img = scene(rng, h, w, kind)
        for extra in ([], [cv2.IMWRITE_JPEG_OPTIMIZE, 1], [cv2.IMWRITE_JPEG_RST_INTERVAL, 3]):
            ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality] + extra)
            assert ok
            blobs.append(enc.tobytes())
            expect.append(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
    seq = sequence_of(tmp_path, blobs)
    assert seq.getNumImages() == len(blobs)
    for i, exp in enumerate(expect):
        got = seq.getImageRaw_internal(i)
        assert got is not None, f"file {i} rejected"
        assert got.shape == exp.shape and np.array_equal(got, exp), f"file {i}: {np.count_nonzero(got != exp)} pixels differ"


@pytest.mark.parametrize("sampling", ["444", "422", "420", "411", "440"])
def test_colour_jpeg_read_as_grey_matches_opencv(tmp_path, sampling):
    """Grey read of a colour file = its luminance component (libjpeg out_color_space = JCS_GRAYSCALE)."""
    factor = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling, None)
    if factor is None or not hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        pytest.skip("this OpenCV cannot choose JPEG sampling factors")
    rng = np.random.default_rng(5)
    blobs, expect = [], []
    for (w, h) in ((96, 64), (75, 53), (33, 17)):
        bgr = np.stack([scene(rng, h, w, k) for k in ("smooth", "noise", "ramp")], axis=-1)
        for quality in (50, 95):
            for extra in ([], [cv2.IMWRITE_JPEG_RST_INTERVAL, 2]):
                ok, enc = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, factor] + extra)
                assert ok
                blobs.append(enc.tobytes())
                expect.append(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
    seq = sequence_of(tmp_path, blobs)
    for i, exp in enumerate(expect):
        got = seq.getImageRaw_internal(i)
        assert got is not None and np.array_equal(got, exp), f"file {i}"


def test_unsupported_and_corrupt_files_are_rejected(tmp_path):
    from mono_dataset_code_b200 import _lib
    img = scene(np.random.default_rng(1), 40, 56, "smooth")
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    ok2, base = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 80])
    blob = base.tobytes()
    seq = sequence_of(tmp_path, [prog.tobytes(), blob[:len(blob) // 3], blob[:2] + b"\x00\x11" + blob[4:]])
    assert seq.getImageRaw_internal(0) is None and b"baseline only" in _lib.lib.mdc_last_error()
    assert seq.getImageRaw_internal(2) is None
    trunc = seq.getImageRaw_internal(1)                     # a truncated scan decodes to something (like libjpeg, with a warning) or fails
    assert trunc is None or trunc.shape == img.shape


def test_random_corruption_never_crashes(tmp_path):
    """Bit flips and truncations anywhere in the file: the decoder either returns an image of the declared size or an error."""
    rng = np.random.default_rng(11)
    img = scene(rng, 48, 64, "ramp")
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 75, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    base = bytearray(enc.tobytes())
    blobs = []
    for k in range(300):
        b = bytearray(base)
        for _ in range(rng.integers(1, 4)):
            b[rng.integers(2, len(b))] ^= 1 << rng.integers(0, 8)
        if k % 5 == 0:
            b = b[: rng.integers(4, len(b))]
        blobs.append(bytes(b))
    seq = sequence_of(tmp_path, blobs)
    shapes = set()
    for i in range(len(blobs)):
        got = seq.getImageRaw_internal(i)
        if got is not None:
            shapes.add(got.shape)
    assert (48, 64) in shapes            # most single-bit flips in the entropy data still decode


def test_oversubscribed_huffman_table_is_rejected(tmp_path):
    """A DHT segment whose code lengths over-subscribe the prefix code (255 codes of length 1) passes the `total <= 256` check but
    must be refused before the 9-bit prefix table is filled (libjpeg: JERR_BAD_HUFF_TABLE); it used to write past the table."""
    from mono_dataset_code_b200 import _lib
    img = scene(np.random.default_rng(2), 24, 32, "smooth")
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 80])
    blob = enc.tobytes()
    at = blob.find(b"\xff\xc4")
    assert at > 0
    seg_len = (blob[at + 2] << 8) | blob[at + 3]
    bad_counts = bytes([255] + [0] * 15)
    for counts in (bad_counts, bytes([1, 3] + [0] * 14), bytes([0] * 8 + [255] + [0] * 7 )):
        total = sum(counts)
        dht = b"\xff\xc4" + (2 + 1 + 16 + total).to_bytes(2, "big") + b"\x00" + counts + bytes([i & 0xff for i in range(total)])
        evil = blob[:at] + dht + blob[at + 2 + seg_len:]
        seq = sequence_of(tmp_path / f"c{counts[0]}_{counts[1]}_{counts[8]}", [evil])
        got = seq.getImageRaw_internal(0)
        if counts[0] == 255 or counts[:2] == bytes([1, 3]):
            assert got is None and b"Huffman" in _lib.lib.mdc_last_error()
        else:       # 255 codes of length 9 fit the code space: a legal (if useless) table, decoded or rejected later without a crash
            assert got is None or got.shape == img.shape


def test_scalar_and_avx2_transforms_agree_with_opencv(tmp_path):
    """The inverse DCT exists twice (AVX2, chosen at run time, and the portable 64-bit scalar code).  Whatever this machine picks is
    what the tests above exercise; here a child process forces the scalar one (MDC_JPEG_SCALAR=1) and both must return OpenCV's bytes."""
    import subprocess
    import sys
    rng = np.random.default_rng(4)
    blobs, expect = [], []
    for kind, quality in CASES:
        img = scene(rng, 120, 200, kind)
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
        assert ok
        blobs.append(enc.tobytes())
        expect.append(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
    seq = sequence_of(tmp_path, blobs)
    for i, exp in enumerate(expect):
        assert np.array_equal(seq.getImageRaw_internal(i), exp)
    np.save(tmp_path / "expect.npy", np.stack(expect))
    child = ("import sys, numpy as np; sys.path.insert(0, %r); from mono_dataset_code_b200 import api; "
             "s = api.Sequence(%r); e = np.load(%r); "
             "bad = [i for i in range(len(e)) if not np.array_equal(s.getImageRaw_internal(i), e[i])]; print('BAD', bad); sys.exit(1 if bad else 0)"
             % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path), str(tmp_path / "expect.npy")))
    r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, MDC_JPEG_SCALAR="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_random_sizes_qualities_and_modes_match_opencv(tmp_path):
    """Property-style sweep (seeded): random sizes incl. non-multiples of the MCU, qualities 1..100, grey / colour with every sampling mode,
    optimised Huffman tables, restart intervals, progressive rejected — every accepted file decodes to OpenCV's bytes."""
    rng = np.random.default_rng(2024)
    sampling = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
                "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440} if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR") else {}
    blobs, expect = [], []
    for _ in range(60):
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 160))
        kind = ("smooth", "noise", "edges", "ramp")[int(rng.integers(0, 4))]
        img = scene(rng, h, w, kind)
        params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(1, 101))]
        if rng.random() < 0.3:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        if rng.random() < 0.3:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 9))]
        if sampling and rng.random() < 0.5:
            img = np.stack([img, np.roll(img, 3, 1), 255 - img], axis=2)      # a colour file, read as grey
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, list(sampling.values())[int(rng.integers(0, len(sampling)))]]
        ok, enc = cv2.imencode(".jpg", img, params)
        assert ok
        blobs.append(enc.tobytes())
        expect.append(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
    seq = sequence_of(tmp_path, blobs)
    for i, exp in enumerate(expect):
        got = seq.getImageRaw_internal(i)
        assert got is not None, f"file {i} rejected"
        assert got.shape == exp.shape and np.array_equal(got, exp), f"file {i}: {np.count_nonzero(got != exp)} pixels differ"
