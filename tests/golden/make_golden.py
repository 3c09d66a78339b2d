"""Generates tests/golden/*.npz from the REFERENCE'S OWN CODE (oracle/_ref, i.e.
/root/reference/src/{FOVUndistorter,PhotometricUndistorter}.cpp compiled unmodified).

Run in the authoring container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist; tests compare the
C restatement (CPU) and the CUDA path (GPU) against them.
Contents per case: calibration text, vignette pixels, frames, the reference's remap tables,
GInv/G, vignetteMapInv, and getImage-equivalent outputs (unMapImage -> undistort<float>,
undistort<uchar>, unMapImage alone) for all 16 flag combinations.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mono_dataset_code_b200 import synthetic as S  # noqa: E402
from oracle import loader  # noqa: E402

CASES = {
    # name: (in_w, in_h, out_w, out_h, mode, calib, vignette_depth, zeros)
    "tiny_crop": (96, 80, 88, 72, "crop", S.TUM_CALIB, 16, False),
    "tiny_full": (80, 64, 96, 80, "full", (0.349153, 0.436593, 0.493140, 0.499021, 0.6), 8, True),
    "tiny_explicit": (128, 96, 64, 48, "0.4 0.53 0.5 0.5 0", S.TUM_CALIB, 16, False),
}


def ref_get_image(fov, photo, raw, rectify, g, v, k):
    """DatasetReader::getImage (BenchmarkDatasetReader.h:210-241) composed from the reference's two operators."""
    if g or v or k:
        t = photo.unmap(raw, g, v, k)
        return fov.undistort(t) if rectify else t
    return fov.undistort(raw) if rectify else raw.astype(np.float32)


def main():
    loader.build("ref")
    R = loader.RefOracle()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (iw, ih, ow, oh, mode, calib, depth, zeros) in CASES.items():
        d = tempfile.mkdtemp()
        files = S.write_dataset_dir(d, iw, ih, ow, oh, mode, calib, vignette_depth=depth, vignette_zeros=zeros)
        R.register_image(files["vignette"], files["vignette_pixels"])
        fov = R.fov(files["camera"])
        photo = R.photo(files["pcalib"], files["vignette"], iw, ih)
        assert fov.valid and photo.valid_gamma and photo.valid_vignette
        rx, ry = fov.tables()
        frames = np.stack([S.frame(0, iw, ih, "uniform"), S.frame(1, iw, ih, "speckle"), S.frame(2, iw, ih, "gradient")])
        data = {
            "camera_txt": np.frombuffer(open(files["camera"], "rb").read(), np.uint8),
            "pcalib_txt": np.frombuffer(open(files["pcalib"], "rb").read(), np.uint8),
            "vignette_pixels": files["vignette_pixels"], "frames": frames,
            "remap_x": rx, "remap_y": ry, "ginv": photo.ginv(), "g": photo.g(),
            "vinv": photo.vignette_maps()[1], "k_rect": fov.K()[0], "k_org": fov.K()[1],
            "dims": np.array([iw, ih, ow, oh], np.int32),
        }
        for flags in range(16):
            rectify, g, v, k = flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, (flags >> 3) & 1
            outs = np.stack([ref_get_image(fov, photo, f, rectify, g, v, k) for f in frames])
            data[f"out_{flags:02d}"] = outs
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **data)
        h = hashlib.sha256(b"".join(np.ascontiguousarray(data[k]).tobytes() for k in sorted(data))).hexdigest()[:16]
        print(name, os.path.getsize(path) // 1024, "KiB", h)


if __name__ == "__main__":
    main()
