#!/bin/bash
# compute-sanitizer passes over a small K1 workload (both loaders, pyramid on) + E-step: memcheck, racecheck, synccheck.
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, tempfile, numpy as np, torch
sys.path.insert(0, '.')
from mono_dataset_code_b200 import api, synthetic as S
iw, ih, ow, oh = 320, 240, 288, 200
files = S.write_dataset_dir(tempfile.mkdtemp(), iw, ih, ow, oh, "crop")
fov = api.UndistorterFOV(files["camera"]); photo = api.PhotometricUndistorter(files["pcalib"], files["vignette"], iw, ih)
prep = api.FramePreparer(fov, photo, 0)
fr = torch.from_numpy(S.frames(70, iw, ih)).cuda()
for tma in (1, 0):
    prep.ctx.configure(use_tma=tma)
    for lv in (1, 5):
        out = prep.prepare_device(fr, True, True, True, True, levels=lv)
torch.cuda.synchronize()
ctx = api.Context(None, None, 0)
data = torch.randint(0, 256, (9, 4096), dtype=torch.uint8, device="cuda")
E = torch.zeros(4096, dtype=torch.float64, device="cuda")
ctx.estep(data, torch.linspace(0.1, 2, 9, dtype=torch.float64, device="cuda"), torch.linspace(0, 255, 256, dtype=torch.float64, device="cuda"), E)
torch.cuda.synchronize()
print("done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/sanitize_$tool.log | grep -v "^Input\|Failed to read" | head -8
done
