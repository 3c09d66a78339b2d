#!/bin/bash
# Round-end verification in one gpurun call (1 GPU): smoke, the whole GPU test suite, both bench arms, the K1 loader A/B and the
# calibrator timings.  Multi-GPU: scripts/gpu_multi.sh N under `gpurun --gpus N`.  ncu evidence: scripts/gpu_prof.sh.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
{ echo "nproc: $(nproc)"; echo "cgroup cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"; lscpu | grep -E "Model name|Socket|NUMA node|Thread|Core"; free -g | head -2; } > gpurun_out/host.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -v -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --tma 0 --no-cpu --no-estep --no-seq > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err; echo "bench ldg rc=$?"
timeout 300 python scripts/k1_study.py --quick 2>/dev/null | grep '^{' > gpurun_out/k1_loaders.jsonl; cut -c1-200 gpurun_out/k1_loaders.jsonl
timeout 200 python scripts/estep_time.py 2>&1 | grep -v RMSE | tail -1 > gpurun_out/estep_time.json; cut -c1-400 gpurun_out/estep_time.json
timeout 200 python scripts/vc_time.py 2>&1 | tail -1 > gpurun_out/vc_time.json; cat gpurun_out/vc_time.json
