#!/bin/bash
# racecheck over the calibrator's streaming kernels only: a short stack (the ring is filled once per tile, no stage is reused) and a
# long one (many ring wraps, the G-step's mid-run fold)
set -u
mkdir -p gpurun_out
cat > gpurun_out/san_calib.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from mono_dataset_code_b200 import api
ctx = api.Context(None, None, 0)
n, npix = int(sys.argv[1]), int(sys.argv[2])
data = torch.randint(0, 256, (n, npix), dtype=torch.uint8, device="cuda")
t = torch.linspace(0.1, 2, n, dtype=torch.float64, device="cuda")
G = torch.linspace(0, 255, 256, dtype=torch.float64, device="cuda")
E = torch.zeros(npix, dtype=torch.float64, device="cuda")
ctx.estep(data, t, G, E)
G2 = torch.zeros_like(G)
ctx.rc_gstep(data, t, E, G2)
ctx.rc_rmse(data, t, G, E)
torch.cuda.synchronize()
print("done")
PY
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for cfg in "21 4656" "200 7424"; do
  tag=$(echo $cfg | tr ' ' 'x')
  timeout 400 compute-sanitizer --tool racecheck --print-limit 4 python gpurun_out/san_calib.py $cfg > gpurun_out/sanitize_racecheck_calib_$tag.log 2>&1
  echo "racecheck $cfg rc=$?"; grep -E "RACECHECK SUMMARY|Race reported" gpurun_out/sanitize_racecheck_calib_$tag.log | cut -c1-200 | head -4
done
