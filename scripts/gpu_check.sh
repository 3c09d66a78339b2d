#!/bin/bash
# One gpurun call: smoke -> GPU parity tests -> bench -> ncu launch list -> ncu full capture of K1.
# Everything is wrapped in `timeout` so a hung kernel cannot eat the box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench ldg"; timeout 600 python bench.py --steps 10 --warmup 3 --tma 0 --no-cpu > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err; echo "bench rc=$?"; cat gpurun_out/bench_ldg.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --batch 64 --e2e-batch 16 --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_prepare -s 3 -c 2 -o gpurun_out/prof_k1 -f \
    python bench.py --steps 3 --warmup 3 --batch 64 --e2e-batch 16 --no-cpu > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out
