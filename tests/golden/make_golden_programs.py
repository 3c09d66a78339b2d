"""Generates tests/golden/programs/response_calib.npz from a RUN OF THE REFERENCE PROGRAM (oracle/_ref/responseCalib_ref =
/root/reference/src/main_responseCalib.cpp compiled unmodified, `make -C oracle ref`).

Run in the authoring container (needs /root/reference):   python tests/golden/make_golden_programs.py
Contents: the raw frames and exposure times of a small sequence, the program's arguments, and what it wrote: the inverse response
of photoCalibResult/pcalib.txt and the per-iteration {sample count, rmse} of photoCalibResult/log.txt.  The GPU test runs
mdc_rc_leak_padding + mdc_response_calib on the same frames and compares with the program's output directly.
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import loader  # noqa: E402
import test_response_calib_reference_program as T  # noqa: E402


def main():
    loader.build("ref")
    out_dir = os.path.join(ROOT, "tests", "golden", "programs")
    os.makedirs(out_dir, exist_ok=True)
    nits, leak = 5, 2
    with tempfile.TemporaryDirectory() as tmp:
        seq, work = Path(tmp) / "seq", Path(tmp) / "work"
        seq.mkdir(); work.mkdir()
        frames, t = T.write_sequence(seq, np.random.default_rng(2024))
        r = subprocess.run([T.EXE, str(seq) + "/", f"iterations={nits}", f"leakPadding={leak}"], cwd=work, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        G = np.array(open(work / "photoCalibResult" / "pcalib.txt").read().split(), dtype=np.float64)
        log = np.loadtxt(work / "photoCalibResult" / "log.txt", ndmin=2)
    np.savez_compressed(os.path.join(out_dir, "response_calib.npz"), frames=frames, exposures=t, w=T.W, h=T.H, nits=nits, leak_padding=leak,
                        G=G, log_num=log[:, 2], log_rmse=log[:, 3])
    print("wrote response_calib.npz:", frames.shape, "G[255] =", G[255], "rmse", log[:, 3])


if __name__ == "__main__":
    main()
