#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sanitizer (TMA path, tiny)"
cat > /tmp/tiny.py <<'PY'
import sys, tempfile, numpy as np, torch
sys.path.insert(0, '.')
from mono_dataset_code_b200 import api, synthetic as S
iw, ih, ow, oh = 640, 480, 640, 480
files = S.write_dataset_dir(tempfile.mkdtemp(), iw, ih, ow, oh, "crop")
fov = api.UndistorterFOV(files["camera"]); photo = api.PhotometricUndistorter(files["pcalib"], files["vignette"], iw, ih)
prep = api.FramePreparer(fov, photo, 0)
prep.ctx.configure(use_tma=1)
fr = torch.from_numpy(S.frames(4, iw, ih)).cuda()
out = prep.prepare_device(fr, True, True, True, False, levels=1)
torch.cuda.synchronize()
print("ok", float(out[0][0, 1000]))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/tiny.py > gpurun_out/sanitizer.log 2>&1; echo "rc=$?"; grep -v "^Input\|^Out\|^new K\|^old K\|Reading\|Success" gpurun_out/sanitizer.log | head -60
echo "== ncu full (LDG variant)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_prepare -s 3 -c 1 -o gpurun_out/prof_k1_ldg -f \
    python bench.py --steps 3 --warmup 3 --batch 64 --only-kernel --tma 0 > gpurun_out/ncu_full_ldg.log 2>&1; echo "ncu full rc=$?"
