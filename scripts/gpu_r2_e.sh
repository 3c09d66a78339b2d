#!/bin/bash
# Full single-GPU check: smoke, GPU tests, bench (both arms).
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-6000; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-2500
