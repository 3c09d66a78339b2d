"""ctypes front-ends for the two CPU oracles (TEST INFRASTRUCTURE).

* ``RefOracle``  – the reference's own FOVUndistorter.cpp / PhotometricUndistorter.cpp,
  compiled unmodified into ``oracle/_ref/libmdc_oracle_ref[_f].so`` (oracle/Makefile).
* ``PortOracle`` – this repo's plain-C restatement, ``oracle/_port/libmdc_oracle_port.so``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_DIR = os.path.join(HERE, "_port")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_ubyte)
_f64p = C.POINTER(C.c_double)


def _p(a, ty):
    return a.ctypes.data_as(ty)


def build(which: str = "all") -> None:
    """Run oracle/Makefile (`port`, `ref` or `all`).  `ref` is a no-op without /root/reference."""
    subprocess.run(["make", "-s", "-C", HERE, which], check=True, stdout=subprocess.DEVNULL)


def ref_available(float_math: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libmdc_oracle_ref_f.so" if float_math else "libmdc_oracle_ref.so"))


class RefOracle:
    """The reference's classes behind a C wrapper (oracle/ref/ref_wrapper.cpp)."""

    def __init__(self, float_math: bool = False):
        name = "libmdc_oracle_ref_f.so" if float_math else "libmdc_oracle_ref.so"
        self.lib = L = C.CDLL(os.path.join(REF_DIR, name))
        L.oref_fov_create.restype = C.c_void_p
        L.oref_fov_create.argtypes = [C.c_char_p]
        L.oref_photo_create.restype = C.c_void_p
        L.oref_photo_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        for fn in ("oref_fov_remap_x", "oref_fov_remap_y", "oref_photo_ginv", "oref_photo_g",
                   "oref_photo_vignette_map", "oref_photo_vignette_map_inv"):
            getattr(L, fn).restype = _f32p
            getattr(L, fn).argtypes = [C.c_void_p]
        for fn in ("oref_fov_valid", "oref_photo_valid_vignette", "oref_photo_valid_gamma"):
            getattr(L, fn).restype = C.c_int
            getattr(L, fn).argtypes = [C.c_void_p]
        L.oref_fov_destroy.argtypes = [C.c_void_p]
        L.oref_photo_destroy.argtypes = [C.c_void_p]
        L.oref_fov_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.oref_fov_K.argtypes = [C.c_void_p, _f32p, _f32p]
        L.oref_fov_omega.restype = C.c_float
        L.oref_fov_omega.argtypes = [C.c_void_p]
        L.oref_fov_original_calibration.argtypes = [C.c_void_p, _f32p]
        L.oref_fov_distort.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
        L.oref_fov_undistort_f32.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
        L.oref_fov_undistort_u8.argtypes = [C.c_void_p, _u8p, _f32p, C.c_int, C.c_int]
        L.oref_photo_unmap.argtypes = [C.c_void_p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oref_time_frames.restype = C.c_double
        L.oref_time_frames.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, _f64p]
        L.mdc_shim_register_image.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oref_pool_create.restype = C.c_void_p
        L.oref_pool_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oref_pool_run.restype = C.c_double
        L.oref_pool_run.argtypes = [C.c_void_p, C.c_int, _f64p]
        L.oref_pool_threads.argtypes = [C.c_void_p]
        L.oref_pool_numa_nodes.argtypes = [C.c_void_p]
        L.oref_pool_destroy.argtypes = [C.c_void_p]

    # -- image registry of the shim's cv::imread (pixels decoded by the caller, e.g. with cv2)
    def register_image(self, path: str, pixels: np.ndarray) -> None:
        assert pixels.ndim == 2 and pixels.dtype in (np.uint8, np.uint16)
        px = np.ascontiguousarray(pixels)
        self.lib.mdc_shim_register_image(path.encode(), px.shape[0], px.shape[1], 0 if px.dtype == np.uint8 else 2,
                                         px.ctypes.data_as(C.c_void_p))

    def clear_images(self) -> None:
        self.lib.mdc_shim_clear_images()

    def fov(self, camera_txt: str) -> "RefFov":
        return RefFov(self, camera_txt)

    def photo(self, pcalib: str, vignette: str, w: int, h: int) -> "RefPhoto":
        return RefPhoto(self, pcalib, vignette, w, h)

    def time_frames(self, fov: "RefFov", photo: "RefPhoto", frames: np.ndarray, n_frames: int, threads: int,
                    flags=(1, 1, 0)):
        """Wall seconds for n_frames of unMapImage -> undistort<float>; (seconds, [unmap_s, undistort_s] of worker 0)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        st = np.zeros(2, np.float64)
        s = self.lib.oref_time_frames(fov.h, photo.h, _p(frames, _u8p), frames.shape[0], n_frames, threads,
                                      flags[0], flags[1], flags[2], _p(st, _f64p))
        return s, st


class RefPool:
    """All-core arm of the reference's per-frame path: persistent pinned workers with private buffers, one replica of the
    reference objects and of the input frames per NUMA node, timed inside the library (oracle/ref/ref_wrapper.cpp)."""

    def __init__(self, o: RefOracle, camera_txt: str, pcalib: str, vignette: str, w: int, h: int, frames: np.ndarray,
                 threads: int = 0, flags=(1, 1, 0)):
        self.L = o.lib
        frames = np.ascontiguousarray(frames, np.uint8)
        self.h = self.L.oref_pool_create(camera_txt.encode(), pcalib.encode(), vignette.encode(), w, h, _p(frames, _u8p),
                                         frames.shape[0], threads, flags[0], flags[1], flags[2])
        if not self.h:
            raise RuntimeError("oref_pool_create failed (invalid calibration?)")
        self.threads = int(self.L.oref_pool_threads(self.h))
        self.numa_nodes = int(self.L.oref_pool_numa_nodes(self.h))

    def run(self, frames_per_worker: int):
        """(seconds measured inside the library, (min, max) per-worker busy seconds) for threads*frames_per_worker frames."""
        mm = np.zeros(2, np.float64)
        s = float(self.L.oref_pool_run(self.h, frames_per_worker, _p(mm, _f64p)))
        return s, (float(mm[0]), float(mm[1]))

    def close(self):
        if getattr(self, "h", None):
            self.L.oref_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefFov:
    def __init__(self, o: RefOracle, camera_txt: str):
        self.o, self.L = o, o.lib
        self.h = self.L.oref_fov_create(camera_txt.encode())
        self.valid = bool(self.L.oref_fov_valid(self.h))
        self.in_w = self.in_h = self.out_w = self.out_h = 0
        if self.valid:
            d = (C.c_int * 4)()
            self.L.oref_fov_dims(self.h, d)
            self.in_w, self.in_h, self.out_w, self.out_h = list(d)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oref_fov_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.out_w * self.out_h
        return (np.ctypeslib.as_array(self.L.oref_fov_remap_x(self.h), (n,)).copy(),
                np.ctypeslib.as_array(self.L.oref_fov_remap_y(self.h), (n,)).copy())

    def K(self):
        a, b = np.zeros(9, np.float32), np.zeros(9, np.float32)
        self.L.oref_fov_K(self.h, _p(a, _f32p), _p(b, _f32p))
        return a.reshape(3, 3), b.reshape(3, 3)

    def omega(self) -> float:
        return float(self.L.oref_fov_omega(self.h))

    def original_calibration(self):
        v = np.zeros(5, np.float32)
        self.L.oref_fov_original_calibration(self.h, _p(v, _f32p))
        return v

    def distort(self, x: np.ndarray, y: np.ndarray):
        x = np.ascontiguousarray(x, np.float32).copy()
        y = np.ascontiguousarray(y, np.float32).copy()
        self.L.oref_fov_distort(self.h, _p(x, _f32p), _p(y, _f32p), x.size)
        return x, y

    def undistort(self, img: np.ndarray, out: np.ndarray | None = None, n_in=None, n_out=None):
        n_out_true = self.out_w * self.out_h
        if out is None:
            out = np.full(max(n_out_true, 1), -12345.0, np.float32)
        n_in = img.size if n_in is None else n_in
        n_out = n_out_true if n_out is None else n_out
        img = np.ascontiguousarray(img)
        if img.dtype == np.uint8:
            self.L.oref_fov_undistort_u8(self.h, _p(img, _u8p), _p(out, _f32p), n_in, n_out)
        else:
            assert img.dtype == np.float32
            self.L.oref_fov_undistort_f32(self.h, _p(img, _f32p), _p(out, _f32p), n_in, n_out)
        return out


class RefPhoto:
    def __init__(self, o: RefOracle, pcalib: str, vignette: str, w: int, h: int):
        self.o, self.L = o, o.lib
        self.w, self.hh = w, h
        self.h = self.L.oref_photo_create(pcalib.encode(), vignette.encode(), w, h)
        self.valid_gamma = bool(self.L.oref_photo_valid_gamma(self.h))
        self.valid_vignette = bool(self.L.oref_photo_valid_vignette(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oref_photo_destroy(self.h)
            self.h = None

    def ginv(self):
        p = self.L.oref_photo_ginv(self.h)
        return np.ctypeslib.as_array(p, (256,)).copy() if p else None

    def g(self):
        p = self.L.oref_photo_g(self.h)
        return np.ctypeslib.as_array(p, (256,)).copy() if p else None

    def vignette_maps(self):
        if not self.valid_vignette:
            return None, None
        n = self.w * self.hh
        return (np.ctypeslib.as_array(self.L.oref_photo_vignette_map(self.h), (n,)).copy(),
                np.ctypeslib.as_array(self.L.oref_photo_vignette_map_inv(self.h), (n,)).copy())

    def unmap(self, img: np.ndarray, gamma: bool, vignette: bool, kill: bool):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(img.size, np.float32)
        self.L.oref_photo_unmap(self.h, _p(img, _u8p), _p(out, _f32p), img.size, int(gamma), int(vignette), int(kill))
        return out


class _PortFov(C.Structure):
    _fields_ = [("in_calib", C.c_float * 5), ("out_calib", C.c_float * 5), ("in_w", C.c_int), ("in_h", C.c_int),
                ("out_w", C.c_int), ("out_h", C.c_int), ("float_math", C.c_int)]


MODE_CROP, MODE_FULL, MODE_EXPLICIT = -1, -2, 0


class PortOracle:
    """Plain-C restatement (oracle/port/mdc_oracle_port.c)."""

    def __init__(self):
        path = os.path.join(PORT_DIR, "libmdc_oracle_port.so")
        if not os.path.exists(path):
            build("port")
        self.lib = L = C.CDLL(path)
        L.oport_fov_init.argtypes = [C.POINTER(_PortFov), _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int]
        L.oport_fov_distort.argtypes = [C.POINTER(_PortFov), _f32p, _f32p, C.c_int]
        L.oport_fov_build_tables.argtypes = [C.POINTER(_PortFov), _f32p, _f32p]
        L.oport_fov_build_tables.restype = C.c_int
        L.oport_fov_K.argtypes = [C.POINTER(_PortFov), _f32p, _f32p]
        L.oport_undistort_f32.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p]
        L.oport_undistort_u8.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _u8p, _f32p]
        L.oport_photo_tables.argtypes = [_f32p, _f32p, _f32p]
        L.oport_photo_tables.restype = C.c_int
        L.oport_vignette_maps.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]
        L.oport_unmap.argtypes = [_f32p, _f32p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oport_get_image.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _u8p, _f32p, _f32p,
                                      C.c_int, C.c_int, C.c_int, C.c_int]
        L.oport_pyr_down.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.oport_estep.argtypes = [_u8p, C.c_int, C.c_int, _f64p, _f64p, _f64p]
        L.oport_gstep.argtypes = [_u8p, C.c_int, C.c_int, _f64p, _f64p, _f64p]
        L.oport_rmse.argtypes = [_u8p, C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.oport_einit.argtypes = [_u8p, C.c_int, C.c_int, _f64p]
        L.oport_rescale.argtypes = [C.c_int, _f64p, _f64p]
        L.oport_rescale.restype = C.c_double
        L.oport_leak_padding.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int]
        L.oport_vc_plane_step.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _f64p]
        L.oport_vc_vignette_step.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _f64p]
        L.oport_vc_smooth.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.oport_vc_plane_maps.argtypes = [_f64p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p, _f32p]
        L.oport_vc_mask_maps.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.oport_vc_prepare_image.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _f32p]
        L.oport_time_frames.restype = C.c_double
        L.oport_time_frames.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _u8p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]

    # ---- FOV
    def fov(self, in_calib, in_w, in_h, mode, out_calib, out_w, out_h, float_math=False) -> "PortFov":
        return PortFov(self, in_calib, in_w, in_h, mode, out_calib, out_w, out_h, float_math)

    def fov_from_file(self, path: str, float_math=False):
        """Parse camera.txt like FOVUndistorter.cpp:63-123; None if the object would be invalid."""
        try:
            with open(path) as f:
                lines = f.read().split("\n")
        except OSError:
            return None
        lines += [""] * 4
        try:
            c = [float(v) for v in lines[0].split()[:5]]
            d = [int(v) for v in lines[1].split()[:2]]
            if len(c) != 5 or len(d) != 2:
                return None
        except ValueError:
            return None
        l3 = lines[2]
        out_calib = [0, 0, 0, 0, 0]
        if l3 == "crop":
            mode = MODE_CROP
        elif l3 == "full":
            mode = MODE_FULL
        elif l3 == "none":
            return None
        else:
            try:
                out_calib = [float(v) for v in l3.split()[:5]]
                if len(out_calib) != 5:
                    return None
            except ValueError:
                return None
            mode = MODE_EXPLICIT
        try:
            o = [int(v) for v in lines[3].split()[:2]]
            if len(o) != 2:
                return None
        except ValueError:
            return None
        return self.fov(c, d[0], d[1], mode, out_calib, o[0], o[1], float_math)

    def undistort(self, rx, ry, in_w, img):
        out = np.zeros(rx.size, np.float32)
        img = np.ascontiguousarray(img)
        if img.dtype == np.uint8:
            self.lib.oport_undistort_u8(_p(rx, _f32p), _p(ry, _f32p), in_w, rx.size, _p(img, _u8p), _p(out, _f32p))
        else:
            self.lib.oport_undistort_f32(_p(rx, _f32p), _p(ry, _f32p), in_w, rx.size, _p(img, _f32p), _p(out, _f32p))
        return out

    # ---- photometric
    def photo_tables(self, raw256):
        raw = np.ascontiguousarray(raw256, np.float32)
        ginv, g = np.zeros(256, np.float32), np.zeros(256, np.float32)
        ok = self.lib.oport_photo_tables(_p(raw, _f32p), _p(ginv, _f32p), _p(g, _f32p))
        return (ginv, g) if ok else (None, None)

    def vignette_maps(self, pixels: np.ndarray):
        px = np.ascontiguousarray(pixels)
        depth = 8 if px.dtype == np.uint8 else 16
        m, mi = np.zeros(px.size, np.float32), np.zeros(px.size, np.float32)
        self.lib.oport_vignette_maps(px.ctypes.data_as(C.c_void_p), depth, px.size, _p(m, _f32p), _p(mi, _f32p))
        return m, mi

    def unmap(self, ginv, vinv, img, gamma, vignette, kill):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(img.size, np.float32)
        self.lib.oport_unmap(_p(ginv, _f32p) if ginv is not None else None, _p(vinv, _f32p) if vinv is not None else None,
                             _p(img, _u8p), _p(out, _f32p), img.size, int(gamma), int(vignette), int(kill))
        return out

    def get_image(self, rx, ry, in_w, in_h, ginv, vinv, raw, rectify, gamma, vignette, kill):
        raw = np.ascontiguousarray(raw, np.uint8)
        n_out = rx.size if rectify else in_w * in_h
        out = np.zeros(n_out, np.float32)
        tmp = np.zeros(in_w * in_h, np.float32)
        self.lib.oport_get_image(_p(rx, _f32p), _p(ry, _f32p), in_w, in_h, rx.size,
                                 _p(ginv, _f32p) if ginv is not None else None,
                                 _p(vinv, _f32p) if vinv is not None else None,
                                 _p(raw, _u8p), _p(tmp, _f32p), _p(out, _f32p),
                                 int(rectify), int(gamma), int(vignette), int(kill))
        return out

    # ---- pyramid (not in the reference; parity unpinned)
    def pyramid(self, lvl0: np.ndarray, w: int, h: int, levels: int):
        out = [np.ascontiguousarray(lvl0, np.float32).reshape(-1)]
        for _ in range(1, levels):
            dw, dh = w >> 1, h >> 1
            dst = np.zeros(dw * dh, np.float32)
            self.lib.oport_pyr_down(_p(out[-1], _f32p), w, h, _p(dst, _f32p))
            out.append(dst)
            w, h = dw, dh
        return out

    # ---- responseCalib
    def estep(self, data: np.ndarray, t: np.ndarray, G: np.ndarray):
        data = np.ascontiguousarray(data, np.uint8)
        n, npix = data.shape
        t = np.ascontiguousarray(t, np.float64)
        G = np.ascontiguousarray(G, np.float64)
        E = np.zeros(npix, np.float64)
        self.lib.oport_estep(_p(data, _u8p), n, npix, _p(t, _f64p), _p(G, _f64p), _p(E, _f64p))
        return E

    def gstep(self, data, t, E):
        data = np.ascontiguousarray(data, np.uint8)
        n, npix = data.shape
        t = np.ascontiguousarray(t, np.float64)
        E = np.ascontiguousarray(E, np.float64)
        G = np.zeros(256, np.float64)
        self.lib.oport_gstep(_p(data, _u8p), n, npix, _p(t, _f64p), _p(E, _f64p), _p(G, _f64p))
        return G

    def rmse(self, data, t, G, E):
        data = np.ascontiguousarray(data, np.uint8)
        n, npix = data.shape
        out = np.zeros(2, np.float64)
        self.lib.oport_rmse(_p(data, _u8p), n, npix, _p(np.ascontiguousarray(t, np.float64), _f64p),
                            _p(np.ascontiguousarray(G, np.float64), _f64p),
                            _p(np.ascontiguousarray(E, np.float64), _f64p), _p(out, _f64p))
        return out

    def einit(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        n, npix = data.shape
        E = np.zeros(npix, np.float64)
        self.lib.oport_einit(_p(data, _u8p), n, npix, _p(E, _f64p))
        return E

    def rescale(self, E, G):
        return self.lib.oport_rescale(E.size, _p(E, _f64p), _p(G, _f64p))

    def leak_padding(self, img: np.ndarray, w: int, h: int, iters: int):
        img = np.ascontiguousarray(img, np.uint8).copy()
        tmp = np.zeros_like(img)
        self.lib.oport_leak_padding(_p(img, _u8p), _p(tmp, _u8p), w, h, iters)
        return img

    # ---- vignetteCalib optimiser (main_vignetteCalib.cpp:395-585)
    def vc_plane_step(self, images, p2x, p2y, wI, hI, vignette, plane_color, oth2, int_abs=True):
        """Returns (new plane colour, FF, FC, (E, R)); images [n, wI*hI], p2x/p2y [n, gw*gh] float32."""
        n, gwgh = p2x.shape
        pc = np.ascontiguousarray(plane_color, np.float32).copy()
        ff, fc, st = np.zeros(gwgh, np.float32), np.zeros(gwgh, np.float32), np.zeros(2, np.float64)
        self.lib.oport_vc_plane_step(_p(images, _f32p), _p(p2x, _f32p), _p(p2y, _f32p), n, gwgh, wI, hI, _p(vignette, _f32p), _p(pc, _f32p),
                                     int(oth2), int(int_abs), _p(ff, _f32p), _p(fc, _f32p), _p(st, _f64p))
        return pc, ff, fc, st

    def vc_vignette_step(self, images, p2x, p2y, wI, hI, plane_color, vignette, oth2, int_abs=True):
        """Returns (new normalised vignette, TT, CT, (E, R))."""
        n, gwgh = p2x.shape
        v = np.ascontiguousarray(vignette, np.float32).copy()
        tt, ct, st = np.zeros(wI * hI, np.float32), np.zeros(wI * hI, np.float32), np.zeros(2, np.float64)
        self.lib.oport_vc_vignette_step(_p(images, _f32p), _p(p2x, _f32p), _p(p2y, _f32p), n, gwgh, wI, hI, _p(plane_color, _f32p), _p(v, _f32p),
                                        int(oth2), int(int_abs), _p(tt, _f32p), _p(ct, _f32p), _p(st, _f64p))
        return v, tt, ct, st

    def vc_plane_maps(self, H, gw, gh, facw=5.0, fach=5.0):
        X, Y = np.zeros(gw * gh, np.float32), np.zeros(gw * gh, np.float32)
        self.lib.oport_vc_plane_maps(_p(np.ascontiguousarray(H, np.float64).ravel(), _f64p), gw, gh, facw, fach, _p(X, _f32p), _p(Y, _f32p))
        return X, Y

    def vc_mask_maps(self, X, Y, wI, hI):
        self.lib.oport_vc_mask_maps(_p(X, _f32p), _p(Y, _f32p), X.size, wI, hI)

    def vc_prepare_image(self, raw, wI, hI, mean_exposure, exposure, max_abs_grad=255):
        out = np.zeros(wI * hI, np.float32)
        self.lib.oport_vc_prepare_image(_p(np.ascontiguousarray(raw, np.float32), _f32p), wI, hI, float(mean_exposure), float(exposure), max_abs_grad,
                                        _p(out, _f32p))
        return out

    def vc_smooth(self, vignette, wI, hI, iters=4):
        out, tmp = np.zeros(wI * hI, np.float32), np.zeros(wI * hI, np.float32)
        self.lib.oport_vc_smooth(_p(np.ascontiguousarray(vignette, np.float32), _f32p), wI, hI, iters, _p(out, _f32p), _p(tmp, _f32p))
        return out

    def time_frames(self, rx, ry, in_w, in_h, out_w, out_h, ginv, vinv, frames, n_frames, threads, flags=3, levels=1):
        frames = np.ascontiguousarray(frames, np.uint8)
        return self.lib.oport_time_frames(_p(rx, _f32p), _p(ry, _f32p), in_w, in_h, out_w, out_h,
                                          _p(ginv, _f32p), _p(vinv, _f32p), _p(frames, _u8p), frames.shape[0],
                                          n_frames, threads, flags, levels)


class PortFov:
    def __init__(self, o: PortOracle, in_calib, in_w, in_h, mode, out_calib, out_w, out_h, float_math):
        self.o, self.L = o, o.lib
        self.s = _PortFov()
        ic = np.asarray(in_calib, np.float32)
        oc = np.asarray(out_calib, np.float32)
        self.L.oport_fov_init(C.byref(self.s), _p(ic, _f32p), in_w, in_h, mode, _p(oc, _f32p), out_w, out_h, int(float_math))
        self.in_w, self.in_h, self.out_w, self.out_h = in_w, in_h, out_w, out_h
        self.valid = True

    def tables(self):
        n = self.out_w * self.out_h
        rx, ry = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.L.oport_fov_build_tables(C.byref(self.s), _p(rx, _f32p), _p(ry, _f32p))
        return rx, ry

    def K(self):
        a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self.L.oport_fov_K(C.byref(self.s), _p(a, _f32p), _p(b, _f32p))

        def mat(v):
            return np.array([[v[0], 0, v[2]], [0, v[1], v[3]], [0, 0, 1]], np.float32)
        return mat(a), mat(b)

    def distort(self, x, y):
        x = np.ascontiguousarray(x, np.float32).copy()
        y = np.ascontiguousarray(y, np.float32).copy()
        self.L.oport_fov_distort(C.byref(self.s), _p(x, _f32p), _p(y, _f32p), x.size)
        return x, y
