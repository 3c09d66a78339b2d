#!/bin/bash
set -u
mkdir -p gpurun_out
# E-step: does the tile count per CTA (2.2 tiles of 1536 px on 296 CTAs at 1 MP) cost the last 20 %?  Sizes with exactly 2 and 3 tiles per CTA next to 1 MP.
for npix in 909312 1000000 1363968 454656; do timeout 120 python scripts/estep_time.py --npix $npix 2>&1 | grep -v RMSE | tail -1 | cut -c1-330; done | tee gpurun_out/estep_sizes.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_seq.json 2> gpurun_out/bench_seq.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_seq.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c4', json.dumps(d.get('c4_sequence'))[:2200])"; tail -3 gpurun_out/bench_seq.err
