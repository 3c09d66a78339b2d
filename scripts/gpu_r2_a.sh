#!/bin/bash
# Round 2, first GPU call: probes (TMA box origins, texture gather on pitch-linear memory), the texture-gather loader's self-check,
# K1 floor study + loader A/B, GPU parity tests, one full ncu capture of the texture-gather K1.
set -u
mkdir -p gpurun_out
P=scripts/probes/bin
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
{
  for x in 0 1 2 4 8 16; do timeout 30 $P/tma_origin_probe $x 32 1; done
  for x in 1 2 4 8; do timeout 30 $P/tma_origin_probe $x 16 2; done
  for x in 1 2 4; do timeout 30 $P/tma_origin_probe $x 8 4; done
} > gpurun_out/tma_origin_probe.txt 2>&1
cat gpurun_out/tma_origin_probe.txt
timeout 120 $P/tex_gather_probe > gpurun_out/tex_gather_probe.txt 2>&1; echo "tex probe rc=$?"; cat gpurun_out/tex_gather_probe.txt
MDC_VERBOSE=1 timeout 300 python scripts/k1_study.py --quick > gpurun_out/k1_ab.jsonl 2> gpurun_out/k1_ab.err; echo "k1 ab rc=$?"; cat gpurun_out/k1_ab.jsonl; grep "\[mdc\]" gpurun_out/k1_ab.err | head
timeout 900 python scripts/k1_study.py > gpurun_out/k1_study.jsonl 2> gpurun_out/k1_study.err; echo "k1 study rc=$?"; cut -c1-330 gpurun_out/k1_study.jsonl
timeout 300 python scripts/k1_study.py --quick --geom 1920x1080 > gpurun_out/k1_ab_c4.jsonl 2>&1; tail -2 gpurun_out/k1_ab_c4.jsonl | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tex -s 3 -c 1 -o gpurun_out/prof_k1_tex -f \
    python bench.py --steps 3 --warmup 3 --only-kernel > gpurun_out/ncu_full_k1_tex.log 2>&1; echo "ncu k1 tex rc=$?"; tail -2 gpurun_out/ncu_full_k1_tex.log
ls -la gpurun_out | head -30
