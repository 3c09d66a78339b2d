#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for npix in 1000000 909312 500096 250112 125056 2073600; do timeout 120 python scripts/estep_time.py --npix $npix 2>&1 | grep -v RMSE | tail -1 | cut -c1-330; done | tee gpurun_out/estep_sizes2.jsonl
