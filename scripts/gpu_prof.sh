#!/bin/bash
# ncu captures for profiles/: launch list of the default bench command shape, full capture of K1 at the bench batch (256), E-step.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-batch 16 --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_prepare -s 3 -c 1 -o gpurun_out/prof_k1_b256 -f \
    python bench.py --steps 3 --warmup 3 --only-kernel > gpurun_out/ncu_full_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:estep -s 1 -c 1 -o gpurun_out/prof_estep -f \
    python bench.py --steps 3 --warmup 3 --e2e-batch 16 --no-cpu > gpurun_out/ncu_full_estep.log 2>&1; echo "ncu estep rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-200
ls -la gpurun_out | head -20
