"""Time the host JPEG decoder (csrc/mdc_jpeg.cpp through mdc_seq_read_gray8) on frames like bench.py's c4_sequence (1920x1080, quality 90),
one thread, and OpenCV's decoder (libjpeg-turbo) on the same bytes for scale.  Checks that both return the same pixels.  No GPU needed."""
import ctypes as C, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from mono_dataset_code_b200 import _lib

def main():
    W, H, N = 1920, 1080, 6
    root = tempfile.mkdtemp(prefix="mdc_jpegtime_")
    os.makedirs(os.path.join(root, "images"))
    yy, xx = np.mgrid[0:H, 0:W]
    rng = np.random.default_rng(11)
    blobs = []
    for i in range(N):
        img = (xx * (150.0 / W) + yy * (60.0 / H) + 40 * np.sin((xx + 13 * i) * 0.05) * np.cos((yy - 7 * i) * 0.04)
               + ((xx // 64 + yy // 64 + i) % 2) * 30 + rng.normal(0, 1.5, (H, W)))
        img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
        blobs.append(enc.tobytes())
        with open(os.path.join(root, "images", f"{i:05d}.jpg"), "wb") as f:
            f.write(blobs[-1])
    with open(os.path.join(root, "times.txt"), "w") as f:
        for i in range(N):
            f.write(f"{i} {i * 0.05:.3f} 1.0\n")
    h = C.c_void_p()
    _lib.check(_lib.lib.mdc_seq_open(root.encode(), C.byref(h)), "mdc_seq_open")
    out = np.empty(W * H, np.uint8)
    w_, h_ = C.c_int(), C.c_int()
    reps = int(os.environ.get("REPS", "5"))
    same = True
    for i in range(N):
        _lib.check(_lib.lib.mdc_seq_read_gray8(h, i, out.ctypes.data_as(C.c_void_p), out.size, C.byref(w_), C.byref(h_)), "read")
        ref = cv2.imdecode(np.frombuffer(blobs[i], np.uint8), cv2.IMREAD_GRAYSCALE)
        same &= bool(np.array_equal(out.reshape(H, W), ref))
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(N):
            _lib.lib.mdc_seq_read_gray8(h, i, out.ctypes.data_as(C.c_void_p), out.size, C.byref(w_), C.byref(h_))
    ours = (time.perf_counter() - t0) / (reps * N)
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(N):
            cv2.imdecode(np.frombuffer(blobs[i], np.uint8), cv2.IMREAD_GRAYSCALE)
    theirs = (time.perf_counter() - t0) / (reps * N)
    print({"ours_ms_per_frame": round(ours * 1e3, 3), "opencv_ms_per_frame": round(theirs * 1e3, 3), "identical_pixels": same,
           "jpeg_kb": round(np.mean([len(b) for b in blobs]) / 1024, 1)})

if __name__ == "__main__":
    main()
