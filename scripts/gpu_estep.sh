#!/bin/bash
# responseCalib iteration: parity tests of the E-step / G-step / rmse kernels, timings, ncu capture of the E-step.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "estep or response" > gpurun_out/pytest_estep.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_estep.log
for b in 1 0; do MDC_ESTEP_BULK=$b timeout 120 python scripts/estep_time.py 2>&1 | tail -1; done | tee gpurun_out/estep_ab.jsonl
if [ "${NCU:-0}" = 1 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rc_stream|estep" -c 3 -o gpurun_out/prof_rc -f \
    python scripts/estep_time.py > gpurun_out/ncu_full_rc.log 2>&1; echo "ncu rc=$?"
fi
