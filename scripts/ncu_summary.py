#!/usr/bin/env python
"""Turn an `ncu --set full` report into the evidence files under profiles/ (runs here, no GPU needed).

  python scripts/ncu_summary.py <report.ncu-rep> <out_summary.csv> [--traffic <out.json> --frames N --bytes-per-frame B --label TEXT]

The summary keeps the metrics the design discussion uses (time, DRAM bytes, LSU / TEX / shared-memory wavefronts, bank conflicts,
issue and pipe utilisation, stall reasons, registers, occupancy).  With --traffic it also writes the per-launch DRAM traffic that
bench.py reports as `roofline.traffic`, stamped with the git commit and the SHA-256 of the kernel sources it was captured from;
bench.py ignores the file (prints traffic: null) when the sources have changed since."""
import argparse
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["mono_dataset_code_b200/csrc/mdc_kernels.cu", "mono_dataset_code_b200/csrc/mdc_kernels.cuh"]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts", "l1tex__data_pipe_tex_wavefronts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__data_bank_reads.avg.pct", "l1tex__data_bank_writes.avg.pct", "l1tex__t_sector_hit_rate", "l1tex__t_sectors_pipe_tex_mem_texture",
        "l1tex__tex_writeback_active.avg.pct", "lts__t_sector_hit_rate", "lts__throughput.avg.pct",
        "sm__issue_active.avg.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_tma_cycles_active.avg.pct",
        "sm__inst_executed_pipe_tex.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "smsp__average_warps_issue_stalled", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "sm__maximum_warps_per_active_cycle_pct", "launch__shared_mem_per_block")


def sources_sha256():
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("summary")
    ap.add_argument("--traffic")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--bytes-per-frame", type=float, default=6553600.0)
    ap.add_argument("--label", default="")
    ap.add_argument("--launch", type=int, default=0, help="which captured launch (row) to summarise")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + a.launch]
    m = dict(zip(hdr, zip(units, vals)))
    with open(a.summary, "w") as f:
        f.write(f'metric,unit,"value (`ncu --set full --clock-control none`, one launch; {a.label})"\n')
        for k in ("Kernel Name", "Block Size", "Grid Size"):
            if k in m:
                f.write(f'{k},,"{m[k][1]}"\n')
        for h in hdr:
            if any(h.startswith(p) for p in KEEP) and "peer" not in h and "sysmem" not in h:
                f.write(f'{h},{m[h][0]},{m[h][1]}\n')
    print("wrote", a.summary)
    if a.traffic:
        def gb(name):
            u, v = m[name]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        doc = {"kernel": m["Kernel Name"][1], "frames_per_launch": a.frames, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
               "dram_bytes_per_frame": (rd + wr) / a.frames, "algorithmic_bytes_per_frame": a.bytes_per_frame,
               "traffic_over_algorithmic": (rd + wr) / a.frames / a.bytes_per_frame,
               "captured_at_commit": commit, "kernel_sources_sha256": sources_sha256(), "kernel_sources": KERNEL_SOURCES,
               "note": f"dram__bytes_read.sum + dram__bytes_write.sum of {m['Kernel Name'][1]}, one launch of {a.frames} frames, "
                       f"`ncu --set full --clock-control none` ({a.label}); algorithmic = {a.bytes_per_frame:.0f} B/frame"}
        with open(a.traffic, "w") as f:
            json.dump(doc, f, indent=1)
        print("wrote", a.traffic)


if __name__ == "__main__":
    sys.exit(main())
