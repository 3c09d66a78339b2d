#!/bin/bash
# tuning sweep: each line of $SWEEP is "ENV=VAL ENV=VAL ... [-- bench args]"
set -u
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
if [ "${TEST:-1}" = "1" ]; then timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
while IFS= read -r line; do
  [ -z "$line" ] && continue
  envs="${line%%--*}"; args=""; case "$line" in *--*) args="${line#*--}";; esac
  echo "## $line" >> gpurun_out/sweep.jsonl
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --only-kernel $args 2> gpurun_out/sweep.err | tail -1 | cut -c1-160 >> gpurun_out/sweep.jsonl; grep "\[mdc\]" gpurun_out/sweep.err | tail -3 >> gpurun_out/sweep.jsonl
done < "${SWEEP_FILE:-scripts/sweep.txt}"
cat gpurun_out/sweep.jsonl
