#!/bin/bash
# ncu captures for profiles/ (round 2): launch list of the default bench command shape; full captures of K1 (plain and pyramid
# variants + the K2 tail), of the three streaming passes of the calibrator (E-step, G-step, rmse) and of the texture-gather K1.
# Summaries are produced afterwards, on the authoring box, with scripts/ncu_summary.py.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-batch 16 --no-cpu --no-seq > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_prepare -s 3 -c 1 -o gpurun_out/prof_k1 -f \
    python bench.py --steps 3 --warmup 3 --only-kernel > gpurun_out/ncu_full_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fused_prepare|pyr_down2" -s 6 -c 2 -o gpurun_out/prof_k1_pyr -f \
    python bench.py --steps 3 --warmup 3 --only-kernel --levels 5 > gpurun_out/ncu_full_k1_pyr.log 2>&1; echo "ncu k1 pyramid rc=$?"
# rc_stream_kernel launches of scripts/estep_time.py: #1-13 E-step, #14-17 G-step, #18-21 rmse -> capture #13 .. #19
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rc_stream_kernel -s 12 -c 7 -o gpurun_out/prof_calib -f \
    python scripts/estep_time.py > gpurun_out/ncu_full_calib.log 2>&1; echo "ncu calibrator rc=$?"
ls -la gpurun_out/*.ncu-rep
