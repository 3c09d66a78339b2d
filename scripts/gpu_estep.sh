#!/bin/bash
# E-step iteration: TMA origin probe, E-step parity tests, bench (no CPU leg), ncu capture of the E-step.
set -u
mkdir -p gpurun_out
for x in 0 16 2; do timeout 60 scripts/probes/bin/tma_origin_probe $x 32; done > gpurun_out/tma_probe.txt 2>&1
cat gpurun_out/tma_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "estep or response" > gpurun_out/pytest_estep.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_estep.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_nocpu.json 2> gpurun_out/bench_nocpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_nocpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'estep', d.get('c5_estep'))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep -s 1 -c 1 -o gpurun_out/prof_estep -f \
    python bench.py --steps 3 --warmup 3 --e2e-batch 16 --no-cpu > gpurun_out/ncu_full_estep.log 2>&1; echo "ncu estep rc=$?"
for b in 1 0; do MDC_ESTEP_BULK=$b timeout 120 python scripts/estep_time.py 2>&1 | tail -1; done | tee gpurun_out/estep_ab.jsonl
