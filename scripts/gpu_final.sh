#!/bin/bash
# Round-end verification in one gpurun call: smoke, full GPU test suite, bench (TMA + LDG loaders, C4 geometry), timing scripts,
# ncu launch list of the default bench command shape.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --tma 0 --no-cpu --no-estep > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err; echo "bench ldg rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --geom 1920x1080 --only-kernel > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; tail -1 gpurun_out/bench_c4.json | cut -c1-300
for b in 1 0; do MDC_ESTEP_BULK=$b timeout 120 python scripts/estep_time.py 2>&1 | grep -v RMSE | tail -1; done > gpurun_out/estep_ab.jsonl; cat gpurun_out/estep_ab.jsonl | cut -c1-400
timeout 200 python scripts/vc_time.py 2>&1 | tail -1 > gpurun_out/vc_time.json; cat gpurun_out/vc_time.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-batch 16 --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu list rc=$?"
