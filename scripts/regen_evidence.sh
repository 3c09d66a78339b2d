#!/bin/bash
# Runs HERE (no GPU) after `gpurun -- bash scripts/gpu_prof.sh` has merged its reports into gpurun_out/: rewrites the ncu summaries and the
# DRAM-traffic files under profiles/ (stamped with the commit and the SHA-256 of the kernel sources — do this after the LAST edit of
# csrc/mdc_kernels.cu/.cuh, or bench.py reports `traffic: null`).
set -eu
cd "$(dirname "$0")/.."
S="python scripts/ncu_summary.py"
$S gpurun_out/prof_k1.ncu-rep profiles/r02_k1_ncu_summary.csv --traffic profiles/k1_traffic.json --frames 256 --bytes-per-frame 6553600 \
   --label "K1 fused_prepare_kernel<tma,vig,nopyr,3>, 256 C2 frames, scripts/gpu_prof.sh round 2"
$S gpurun_out/prof_k1_pyr.ncu-rep profiles/r02_k1_pyramid_ncu_summary.csv --launch 1 --traffic gpurun_out/c3_a.json --frames 256 --bytes-per-frame 8294400 \
   --label "K1 pyramid variant fused_prepare_kernel<tma,vig,pyr,2> (levels 0-2), 256 C2 frames"
$S gpurun_out/prof_k1_pyr.ncu-rep profiles/r02_k2_pyr_down2_ncu_summary.csv --launch 0 --traffic gpurun_out/c3_b.json --frames 256 --bytes-per-frame 8294400 \
   --label "K2 pyr_down2_kernel (levels 3-4 from level 2), 256 C2 frames"
python - <<'PY'
import json
a, b = json.load(open("gpurun_out/c3_a.json")), json.load(open("gpurun_out/c3_b.json"))
rd, wr = a["dram_bytes_read_per_launch"] + b["dram_bytes_read_per_launch"], a["dram_bytes_write_per_launch"] + b["dram_bytes_write_per_launch"]
doc = dict(a)
doc.update({"kernel": a["kernel"] + " + " + b["kernel"].replace("void ", ""), "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
            "dram_bytes_per_frame": (rd + wr) / 256, "traffic_over_algorithmic": (rd + wr) / 256 / 8294400.0,
            "note": "dram__bytes_read.sum + dram__bytes_write.sum of the two launches of one 5-level step (K1 pyramid variant writing levels 0-2, then "
                    "pyr_down2_kernel re-reading level 2 for levels 3-4), 256 C2 frames, `ncu --set full --clock-control none` (scripts/gpu_prof.sh round 2); "
                    "algorithmic = 8 294 400 B/frame; levels 3-4 (26 MB) are still partly in L2 when the second kernel ends, so its DRAM writes are under-counted"})
json.dump(doc, open("profiles/c3_pyramid_traffic.json", "w"), indent=1)
print("wrote profiles/c3_pyramid_traffic.json")
PY
# (the two captured launches of the pyramid step arrive as pyr_down2 of one step, then the fused kernel of the next)
# rc_stream_kernel launches captured from scripts/estep_time.py: row 0 = E-step, rows 1-4 = G-step (count plane), rows 5-6 = rmse
$S gpurun_out/prof_calib.ncu-rep profiles/r02_k3_estep_ncu_summary.csv --launch 0 --traffic profiles/c5_estep_traffic.json --frames 1 --bytes-per-frame 1008000000 \
   --label "rc_stream_kernel<EstepOp>, n=1000 x 1 MP, scripts/gpu_prof.sh round 2"
$S gpurun_out/prof_calib.ncu-rep profiles/r02_k3_gstep_ncu_summary.csv --launch 2 --label "rc_stream_kernel<GstepOp<count>>, n=1000 x 1 MP (fixed-point limbs on native shared-memory atomics)"
$S gpurun_out/prof_calib.ncu-rep profiles/r02_k3_rmse_ncu_summary.csv --launch 5 --label "rc_stream_kernel<RmseOp>, n=1000 x 1 MP"
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/launches.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
with open("profiles/r02_launches.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[hdr:]:
        w.writerow(r)
print("wrote profiles/r02_launches.csv", len(rows) - hdr, "rows")
PY
