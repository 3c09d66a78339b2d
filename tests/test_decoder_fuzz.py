"""The host decoders read files a user hands them (a crafted Huffman table once smashed the stack): tests/cpp/decoder_fuzz.cpp
damages valid JPEG / PNG / PGM files in seeded random ways and decodes them under AddressSanitizer + UndefinedBehaviorSanitizer, with the
AVX2 and with the scalar inverse DCT."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")
CSRC = os.path.join(ROOT, "mono_dataset_code_b200", "csrc")


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    d = tmp_path_factory.mktemp("fuzz")
    exe = str(d / "decoder_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I" + os.path.join(ROOT, "include"),
           "-I" + CSRC, os.path.join(ROOT, "tests", "cpp", "decoder_fuzz.cpp"), os.path.join(CSRC, "mdc_gray_image.cpp"),
           os.path.join(CSRC, "mdc_jpeg.cpp"), "-lz", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer-enabled g++ here: " + r.stdout[-300:])
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:120, 0:200]
    img = np.clip(40 + 0.6 * xx + rng.normal(0, 12, (120, 200)), 0, 255).astype(np.uint8)
    files = {}
    for name, im, params in (("grey.jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 85]),
                             ("colour_rst.jpg", np.stack([img, 255 - img, img], 2), [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_RST_INTERVAL, 4]),
                             ("grey8.png", img, []), ("grey16.png", img.astype(np.uint16) * 257, []), ("grey.pgm", img, [])):
        path = str(d / name)
        assert cv2.imwrite(path, im, params)
        files[name] = path
    return exe, files


@pytest.mark.parametrize("name", ["grey.jpg", "colour_rst.jpg", "grey8.png", "grey16.png", "grey.pgm"])
@pytest.mark.parametrize("scalar", ["0", "1"], ids=["avx2", "scalar"])
def test_damaged_files_never_trip_the_sanitizers(fuzzer, name, scalar):
    exe, files = fuzzer
    if scalar == "1" and not name.endswith(".jpg"):
        pytest.skip("the inverse DCT only matters for JPEG")
    env = dict(os.environ, MDC_JPEG_SCALAR=scalar, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe, files[name], "1500"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "rounds" in r.stdout


def test_damaged_zip_archives_never_trip_the_sanitizers(tmp_path):
    """The same for the sequence reader's own zip parser (stored and deflated entries)."""
    import zipfile
    exe = str(tmp_path / "zip_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I" + os.path.join(ROOT, "include"),
           "-I" + CSRC, os.path.join(ROOT, "tests", "cpp", "zip_fuzz.cpp"), os.path.join(CSRC, "mdc_sequence.cpp"), os.path.join(CSRC, "mdc_gray_image.cpp"),
           os.path.join(CSRC, "mdc_jpeg.cpp"), "-lz", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer-enabled g++ here: " + r.stdout[-300:])
    rng = np.random.default_rng(8)
    seq = tmp_path / "seq"
    os.makedirs(seq)
    with zipfile.ZipFile(seq / "images.zip", "w") as z:
        for i in range(4):
            img = rng.integers(0, 256, (40, 56)).astype(np.uint8)
            ok, enc = cv2.imencode(".jpg" if i % 2 else ".png", img)
            assert ok
            z.writestr(f"{i:05d}" + (".jpg" if i % 2 else ".png"), enc.tobytes(), zipfile.ZIP_DEFLATED if i >= 2 else zipfile.ZIP_STORED)
    with open(seq / "times.txt", "w") as f:
        for i in range(4):
            f.write(f"{i} {i * 0.05:.3f} 1.0\n")
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe, str(seq), "1200"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2500:]
    assert "archives opened" in r.stdout
