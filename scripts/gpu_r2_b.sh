#!/bin/bash
# Round 2, second GPU call: hybrid loader sweep (staged + texture kernels side by side), texture-kernel floor study, CPU reference arm check.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/k1_study.py --hybrid > gpurun_out/k1_hybrid.jsonl 2> gpurun_out/k1_hybrid.err; echo "hybrid rc=$?"; cut -c1-260 gpurun_out/k1_hybrid.jsonl; tail -3 gpurun_out/k1_hybrid.err
timeout 600 python scripts/k1_study.py > gpurun_out/k1_study2.jsonl 2> gpurun_out/k1_study2.err; echo "study rc=$?"; grep -E '"tex"' gpurun_out/k1_study2.jsonl | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_b.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref arm rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "ref arm 2 rc=$?"; tail -1 gpurun_out/bench_ref2.json | cut -c1-400
timeout 400 python scripts/cpu_ref_scaling.py > gpurun_out/cpu_ref_scaling.jsonl 2> gpurun_out/cpu_ref_scaling.err; echo "cpu scaling rc=$?"; cat gpurun_out/cpu_ref_scaling.jsonl
