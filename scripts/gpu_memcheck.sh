#!/bin/bash
# memcheck alone (the python start-up is paged in first so that the sanitizer's attach timeout is not hit)
set -u
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
sed -n '/^cat > gpurun_out\/san.py/,/^PY$/p' scripts/gpu_sanitize.sh | sed '1d;$d' > gpurun_out/san.py
timeout 420 compute-sanitizer --tool memcheck --print-limit 10 python gpurun_out/san.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Error|error" gpurun_out/sanitize_memcheck.log | grep -v "^Input\|Failed to read" | head -8; tail -3 gpurun_out/sanitize_memcheck.log
