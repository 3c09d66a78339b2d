#!/bin/bash
# compute-sanitizer passes over a small K1 workload (both loaders, pyramid on), the responseCalib streaming kernels (E-step, G-step,
# rmse: bulk-copy ring and generic loaders), device distortCoordinates and the vignetteCalib kernels: memcheck, racecheck, synccheck.
set -u
mkdir -p gpurun_out
cat > gpurun_out/san.py <<'PY'
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
for n, npix in ((21, 1536 * 3 + 48), (200, 3584 * 2 + 256), (9, 1001)):      # bulk-copy ring (full tiles + a partial one; 200 exposures: the G-step folds its limb histograms mid-run), generic loader
    data = torch.randint(0, 256, (n, npix), dtype=torch.uint8, device="cuda")
    t = torch.linspace(0.1, 2, n, dtype=torch.float64, device="cuda")
    G = torch.linspace(0, 255, 256, dtype=torch.float64, device="cuda")
    E = torch.zeros(npix, dtype=torch.float64, device="cuda")
    ctx.estep(data, t, G, E)
    G2 = torch.zeros_like(G)
    ctx.rc_gstep(data, t, E, G2)
    ctx.rc_rmse(data, t, G, E)
    ctx.response_calib(data, t, 2, E, G2)
x = torch.rand(5000, device="cuda") * 300; y = torch.rand(5000, device="cuda") * 200
fov.distortCoordinatesDevice(x, y)
pr = S.vignette_calib_problem(5, 48, 40, 64, 56, seed=1)
d = {k: torch.from_numpy(pr[k]).cuda() for k in ("images", "p2x", "p2y")}
pc, v = torch.zeros(48 * 40, device="cuda"), torch.ones(64 * 56, device="cuda")
ctx.vignette_calib(d["images"], d["p2x"], d["p2y"], 48, 40, 64, 56, 2, 15, pc, v, True)
torch.cuda.synchronize()
print("done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python gpurun_out/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/sanitize_$tool.log | grep -v "^Input\|Failed to read" | head -8
done
