#!/bin/bash
# Single GPU: full GPU test suite + calibrator timings + bench (after the pitch / rmse / E-step / decode-pool changes).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for b in 1 0; do MDC_ESTEP_BULK=$b timeout 120 python scripts/estep_time.py 2>&1 | grep -v RMSE | tail -1; done > gpurun_out/estep_ab.jsonl; cut -c1-600 gpurun_out/estep_ab.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','roofline','e2e','c3_pyramid','c4_sequence','c5_estep','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:1400])
"; tail -3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('REF', d['value'], d['run_config'], d['thread_sweep'])"
