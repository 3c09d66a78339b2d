#!/bin/bash
# Multi-GPU check (gpurun --gpus N): native NCCL table broadcast from C++, pixel-sharded calibrator on 2 GPUs, bench at N ranks.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt 2>&1; cat gpurun_out/gpus_n$N.txt
timeout 900 python -m pytest tests/test_nccl_cpp.py "tests/test_gpu_parity.py::test_pixel_sharded_calibrator_two_gpus_nccl" -m gpu -v -rs > gpurun_out/pytest_multigpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multigpu_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-7000; tail -5 gpurun_out/bench_n$N.err
