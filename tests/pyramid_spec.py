"""Independent statement of the consumer-side pyramid — test infrastructure.

The reference repository has no pyramid; BASELINE.json asks for the one its consumer (DSO) builds.  This file states that convention
directly from DSO's public source (JakobEngel/dso, src/FullSystem/HessianBlocks.cpp, FrameHessian::makeImages):

    wG[l] = w >> l, hG[l] = h >> l
    dI_l[x + y*wl] = 0.25f * (dI_lm[2*x   + 2*y*wlm1]        + dI_lm[2*x+1 + 2*y*wlm1] +
                              dI_lm[2*x   + 2*y*wlm1 + wlm1] + dI_lm[2*x+1 + 2*y*wlm1 + wlm1]);

i.e. float32, the four taps added left to right, ((a + b) + c) + d with a = (2x, 2y), b = (2x+1, 2y), c = (2x, 2y+1),
d = (2x+1, 2y+1), then an exact multiplication by 0.25; a trailing odd row / column is dropped.  It is written with strided
numpy slices, not by looking at oracle/port/mdc_oracle_port.c:oport_pyr_down, so that the C restatement and the CUDA kernels are
checked against a second, independently written formulation (a wrong tap order in both of them would not pass here)."""
import numpy as np


def pyr_down(img: np.ndarray) -> np.ndarray:
    """One level: float32 [h, w] -> float32 [h >> 1, w >> 1]."""
    assert img.dtype == np.float32 and img.ndim == 2
    h, w = img.shape
    h2, w2 = h >> 1, w >> 1
    a = img[0:2 * h2:2, 0:2 * w2:2]
    b = img[0:2 * h2:2, 1:2 * w2:2]
    c = img[1:2 * h2:2, 0:2 * w2:2]
    d = img[1:2 * h2:2, 1:2 * w2:2]
    with np.errstate(invalid="ignore", over="ignore"):
        s = np.add(np.add(np.add(a, b, dtype=np.float32), c, dtype=np.float32), d, dtype=np.float32)
        return np.multiply(np.float32(0.25), s, dtype=np.float32)


def pyramid(level0: np.ndarray, w: int, h: int, levels: int):
    """level0 flat float32 [w*h] -> list of flat float32 arrays, level l has (w >> l) * (h >> l) pixels."""
    out = [np.ascontiguousarray(level0, np.float32).reshape(-1)]
    cur = out[0].reshape(h, w)
    for _ in range(1, levels):
        cur = pyr_down(cur)
        out.append(cur.reshape(-1).copy())
    return out
