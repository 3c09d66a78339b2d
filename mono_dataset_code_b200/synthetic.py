"""Synthetic calibration files and frames (SURVEY.md §8d "Synthetic calibration").

The reference ships no fixtures, so every test/bench input is
generated here, seeded and deterministic: ``camera.txt`` (4-line FOV-model file,
format of /root/reference/src/FOVUndistorter.cpp:63-123), ``pcalib.txt`` (256
increasing floats, PhotometricUndistorter.cpp:70-88), ``vignette.png`` (8- or
16-bit grey, PhotometricUndistorter.cpp:120-147) and mono8 frames.

No oracle or CUDA code is imported here; numpy + zlib only.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

# TUM-monoVO-style wide-angle calibration (synthetic).
TUM_CALIB = (0.349153, 0.436593, 0.493140, 0.499021, 0.933271)


def write_png_gray(path: str, img: np.ndarray) -> None:
    """Minimal lossless PNG writer for 2-D uint8 / uint16 arrays (colour type 0)."""
    assert img.ndim == 2 and img.dtype in (np.uint8, np.uint16)
    h, w = img.shape
    depth = 8 if img.dtype == np.uint8 else 16
    rows = img.astype(">u2").tobytes() if depth == 16 else img.tobytes()
    stride = w * depth // 8
    raw = b"".join(b"\x00" + rows[y * stride:(y + 1) * stride] for y in range(h))

    def chunk(tag: bytes, payload: bytes) -> bytes:
        return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, 0, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def write_pgm(path: str, img: np.ndarray) -> None:
    """Binary PGM (P5); 16-bit samples big-endian."""
    assert img.ndim == 2 and img.dtype in (np.uint8, np.uint16)
    h, w = img.shape
    maxv = 255 if img.dtype == np.uint8 else 65535
    with open(path, "wb") as f:
        f.write(f"P5\n{w} {h}\n{maxv}\n".encode())
        f.write(img.astype(">u2").tobytes() if img.dtype == np.uint16 else img.tobytes())


def camera_txt(in_w: int, in_h: int, out_w: int, out_h: int, mode: str = "crop",
               calib=TUM_CALIB) -> str:
    """mode: 'crop' | 'full' | 'none' | 'fx fy cx cy 0' (explicit output K, relative)."""
    l1 = " ".join(f"{c:.6f}" for c in calib)
    return f"{l1}\n{in_w} {in_h}\n{mode}\n{out_w} {out_h}\n"


def ginv_raw() -> np.ndarray:
    """GInv_raw[i] = 255*(i/255)^2.2 + 1e-3*i  (strictly increasing)."""
    i = np.arange(256, dtype=np.float64)
    return 255.0 * (i / 255.0) ** 2.2 + 1e-3 * i


def pcalib_txt(values: np.ndarray | None = None) -> str:
    v = ginv_raw() if values is None else values
    return " ".join(f"{x:.9g}" for x in v) + "\n"


def vignette_image(w: int, h: int, depth: int = 16, zeros: bool = False) -> np.ndarray:
    """V(x,y) = peak*(1-0.6 r^2), r^2 normalised so that corners have r^2 = 1."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    r2 = ((x - w / 2) ** 2 + (y - h / 2) ** 2) / ((w / 2) ** 2 + (h / 2) ** 2)
    peak = 65000.0 if depth == 16 else 250.0
    v = np.rint(peak * (1.0 - 0.6 * r2))
    img = v.astype(np.uint16 if depth == 16 else np.uint8)
    if zeros:  # exercise 1/0 = inf in vignetteMapInv
        img[0, 0] = 0
        img[h // 2, w // 2 - 1] = 0
        img[h - 1, w - 1] = 0
        img[h // 3, (2 * w) // 3] = 0
    return img


def frame(seed: int, w: int, h: int, kind: str = "uniform") -> np.ndarray:
    """mono8 frame, flat row-major (h*w,)."""
    n = w * h
    rng = np.random.default_rng(1000 + seed)
    if kind == "uniform":
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == "gradient":
        y, x = np.mgrid[0:h, 0:w]
        g = (x * 255.0 / max(w - 1, 1) * 0.6 + y * 255.0 / max(h - 1, 1) * 0.4)
        g = g + rng.normal(0, 6.0, (h, w))
        return np.clip(np.rint(g), 0, 255).astype(np.uint8).reshape(-1)
    if kind == "white":
        return np.full(n, 255, np.uint8)
    if kind == "black":
        return np.zeros(n, np.uint8)
    if kind == "speckle":  # mid-grey with isolated saturated pixels, incl. the 4 corners
        f = np.full((h, w), 128, np.uint8)
        ys = rng.integers(0, h, 64)
        xs = rng.integers(0, w, 64)
        f[ys, xs] = 255
        f[0, 0] = f[0, w - 1] = f[h - 1, 0] = f[h - 1, w - 1] = 255
        f[::32, ::32] = 255
        return f.reshape(-1)
    raise ValueError(kind)


def frames(n_frames: int, w: int, h: int, kind: str = "uniform", seed0: int = 0) -> np.ndarray:
    return np.stack([frame(seed0 + f, w, h, kind) for f in range(n_frames)])


def write_dataset_dir(path: str, in_w: int, in_h: int, out_w: int, out_h: int, mode: str = "crop",
                      calib=TUM_CALIB, vignette_depth: int = 16, vignette_zeros: bool = False,
                      pgm_sidecar: bool = True) -> dict:
    """Write camera.txt / pcalib.txt / vignette.png (+ vignette.pgm) into `path`."""
    os.makedirs(path, exist_ok=True)
    files = {"camera": os.path.join(path, "camera.txt"), "pcalib": os.path.join(path, "pcalib.txt"),
             "vignette": os.path.join(path, "vignette.png")}
    with open(files["camera"], "w") as f:
        f.write(camera_txt(in_w, in_h, out_w, out_h, mode, calib))
    with open(files["pcalib"], "w") as f:
        f.write(pcalib_txt())
    vig = vignette_image(in_w, in_h, vignette_depth, vignette_zeros)
    write_png_gray(files["vignette"], vig)
    if pgm_sidecar:
        files["vignette_pgm"] = os.path.join(path, "vignette.pgm")
        write_pgm(files["vignette_pgm"], vig)
    files["vignette_pixels"] = vig
    return files


def vignette_calib_problem(n: int, gw: int, gh: int, wI: int, hI: int, seed: int = 0, noise: float = 0.5):
    """A synthetic vignetteCalib state (what main_vignetteCalib.cpp holds at :395): a textured plane seen from n poses through
    a camera with a known vignette.  Returns dict(images [n, wI*hI], p2x, p2y [n, gw*gh] float32 with NaNs, true_vignette
    [wI*hI], true_plane [gw*gh]).  The maps are smooth projective warps; entries whose rounded position is not strictly inside
    (1, wI-2) x (1, hI-2) are NaN, as the reference makes them (:352-356); a few image pixels are NaN (:300-310)."""
    rng = np.random.default_rng(seed)
    gy, gx = np.mgrid[0:gh, 0:gw].astype(np.float64)
    plane = (110.0 + 60.0 * np.sin(gx * (9.0 / gw)) * np.cos(gy * (7.0 / gh)) + 25.0 * ((gx // max(gw // 8, 1) + gy // max(gh // 8, 1)) % 2)).ravel()
    yy, xx = np.mgrid[0:hI, 0:wI].astype(np.float64)
    r2 = ((xx - wI / 2) ** 2 + (yy - hI / 2) ** 2) / ((wI / 2) ** 2 + (hI / 2) ** 2)
    vig = (1.0 - 0.55 * r2).ravel()
    images = np.empty((n, wI * hI), np.float32)
    p2x = np.empty((n, gw * gh), np.float32)
    p2y = np.empty((n, gw * gh), np.float32)
    u, v = gx / gw - 0.5, gy / gh - 0.5
    for i in range(n):
        ang = rng.uniform(-0.5, 0.5)
        sc = rng.uniform(0.45, 0.8) * min(wI, hI)
        tx, ty = wI / 2 + rng.uniform(-0.2, 0.2) * wI, hI / 2 + rng.uniform(-0.2, 0.2) * hI
        px_, py_ = rng.uniform(-0.15, 0.15, 2)                      # mild perspective
        den = 1.0 + px_ * u + py_ * v
        X = (tx + sc * (np.cos(ang) * u - np.sin(ang) * v) / den).astype(np.float32).ravel()
        Y = (ty + sc * (np.sin(ang) * u + np.cos(ang) * v) / den).astype(np.float32).ravel()
        ui, vi = (X + np.float32(0.5)).astype(np.int64), (Y + np.float32(0.5)).astype(np.int64)
        bad = ~((ui > 1) & (vi > 1) & (ui < wI - 2) & (vi < hI - 2))
        X[bad] = np.nan
        Y[bad] = np.nan
        p2x[i], p2y[i] = X, Y
        # render: every image pixel takes the plane colour of the nearest plane point that projects onto it (good enough
        # for a parity fixture; uncovered pixels keep a flat background), times the vignette, plus noise
        img = np.full(wI * hI, 90.0)
        ok = ~bad
        img[vi[ok] * wI + ui[ok]] = plane[ok]
        img = img * vig + rng.normal(0.0, noise, wI * hI)
        img[rng.integers(0, wI * hI, max(wI * hI // 200, 1))] = np.nan
        images[i] = img.astype(np.float32)
    return dict(images=images, p2x=p2x, p2y=p2y, true_vignette=vig.astype(np.float32), true_plane=plane.astype(np.float32))
