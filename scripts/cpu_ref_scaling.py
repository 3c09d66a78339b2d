#!/usr/bin/env python
"""How the reference's per-frame path scales over the host threads of this box: frames/s at 1, 8, 16, 32, 64, all threads, pinned and
unpinned.  One JSON line each.  (This is the measurement that showed the peak at 16-32 workers — the boxes' 16-CPU cgroup quota —
after which bench.py's CPU arm started to sweep the worker count instead of using every hardware thread.)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

files = bench.write_calibration(tempfile.mkdtemp(prefix="mdc_cpu_"))
cores = os.cpu_count() or 1
for pin in ("1", "0"):
    os.environ["MDC_REF_PIN"] = pin
    for threads in sorted({1, 8, 16, 32, 64, cores}):
        if threads > cores:
            continue
        ref = bench.CpuReference(files, threads)
        ref.all_cores(2)
        n, s = ref.all_cores(12)
        print(json.dumps({"pinned": pin == "1", "threads": threads, "numa_nodes": ref.numa_nodes, "frames_per_s": n / s,
                          "per_thread_ms_per_frame": 1e3 * s / 12}), flush=True)
        ref.close()
