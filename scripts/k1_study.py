#!/usr/bin/env python
"""K1 floor study and loader A/B on one GPU (profiles/r02_k1_floor_study.md is written from its output).

For each loader (texture gather / TMA-staged) the shipped kernel <vignette, no pyramid, 3 CTAs/SM> is timed over a
device-resident batch of 256 C2 frames with parts of its frame loop switched off (MDC_K1_STUDY bit mask: 1 = tap
fetch, 2 = response-LUT look-ups, 4 = level-0 stores; 7 = the product), then a few launch-shape knobs are swept.
Prints one JSON line per configuration.  Study kernels write garbage: timing only.

  python scripts/k1_study.py [--batch 256] [--reps 10] [--geom 1280x1024] [--quick]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--geom", default="1280x1024")
    ap.add_argument("--quick", action="store_true", help="loader A/B only, no study / sweeps")
    ap.add_argument("--hybrid", action="store_true", help="sweep the hybrid loader's split and residency knobs")
    args = ap.parse_args()
    import torch
    from mono_dataset_code_b200 import api, synthetic as S

    w, h = [int(v) for v in args.geom.split("x")]
    files = S.write_dataset_dir(tempfile.mkdtemp(prefix="mdc_study_"), w, h, w, h, "crop")
    fov = api.UndistorterFOV(files["camera"])
    photo = api.PhotometricUndistorter(files["pcalib"], files["vignette"], w, h)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1000)
    B = args.batch
    frames = torch.randint(0, 256, (B, w * h), dtype=torch.uint8, device=dev, generator=g)
    outs = [torch.empty((B, (w >> l) * (h >> l)), dtype=torch.float32, device=dev) for l in range(5)]
    alg = B * (w * h + 4 * w * h)
    peak = 6572.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def run(name, loader, env, levels=1):
        for k in ("MDC_K1_STUDY", "MDC_TEX_TILES", "MDC_TEX_PREFETCH", "MDC_CTAS_PER_SM", "MDC_CHUNK_FRAMES", "MDC_TEX_MAX_ROWS", "MDC_HYB_TEX_PCT", "MDC_HYB_STG_CTAS", "MDC_K1_CARVEOUT"):
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in env.items()})
        ctx = api.Context(fov, photo, 0)
        if not ctx.loader_usable(loader):
            print(json.dumps({"config": name, "loader": loader, "skipped": "loader not usable"}), flush=True)
            return
        ctx.configure(use_tma=loader, ctas_per_sm=int(env.get("MDC_CTAS_PER_SM", 0)))
        try:
            for _ in range(3):
                ctx.prepare_batch(frames, 7, outs[:levels])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                ctx.prepare_batch(frames, 7, outs[:levels])
            b.record()
            torch.cuda.synchronize()
        except Exception as exc:
            print(json.dumps({"config": name, "loader": loader, "env": env, "error": repr(exc)}), flush=True)
            return
        ms = a.elapsed_time(b) / args.reps
        bytes_ = alg + (B * sum((w >> l) * (h >> l) * 4 for l in range(1, levels)) if levels > 1 else 0)
        print(json.dumps({"config": name, "loader": loader, "env": env, "levels": levels, "ms_per_launch": ms, "frames_per_s": B / (ms * 1e-3),
                          "alg_gbs": bytes_ / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / peak}), flush=True)
        ctx.close()

    for loader in ("tex", "tma"):
        run("product", loader, {})
    if args.hybrid:
        for stg in (2, 1):
            for pct in (25, 38, 50, 63):
                for tex_ctas in (3, 4):
                    run(f"hybrid tex {pct}% stg_ctas={stg} tex_ctas={tex_ctas}", "hybrid", {"MDC_HYB_TEX_PCT": pct, "MDC_HYB_STG_CTAS": stg, "MDC_CTAS_PER_SM": tex_ctas})
        run("hybrid tex 38% stg_ctas=2 prefetch", "hybrid", {"MDC_HYB_TEX_PCT": 38, "MDC_HYB_STG_CTAS": 2, "MDC_TEX_PREFETCH": 1})
        run("hybrid tex 38% stg_ctas=1 prefetch tex_ctas=2", "hybrid", {"MDC_HYB_TEX_PCT": 38, "MDC_HYB_STG_CTAS": 1, "MDC_TEX_PREFETCH": 1, "MDC_CTAS_PER_SM": 2})
        run("hybrid tex 38% pyramid", "hybrid", {"MDC_HYB_TEX_PCT": 38, "MDC_HYB_STG_CTAS": 2}, levels=5)
        return
    if args.quick:
        return
    names = {0: "none (loop skeleton)", 1: "taps only", 2: "LUT only", 3: "taps + LUT", 4: "stores only", 5: "taps + stores", 6: "LUT + stores"}
    for loader in ("tex", "tma"):
        for bits in (1, 2, 4, 3, 5, 6, 0):
            run("study: " + names[bits], loader, {"MDC_K1_STUDY": bits})
    for tiles in (1, 4, 8):
        run(f"tex tiles_per_cta={tiles}", "tex", {"MDC_TEX_TILES": tiles})
    for ctas in (2, 4):
        run(f"tex ctas_per_sm={ctas}", "tex", {"MDC_CTAS_PER_SM": ctas})
    run("tex prefetch", "tex", {"MDC_TEX_PREFETCH": 1})
    run("tex prefetch ctas=2", "tex", {"MDC_TEX_PREFETCH": 1, "MDC_CTAS_PER_SM": 2})
    for chunk in (16, 24):
        run(f"tex chunk={chunk}", "tex", {"MDC_CHUNK_FRAMES": chunk})
    run("tex rows<=65000 (chunk 48)", "tex", {"MDC_TEX_MAX_ROWS": 65000})
    run("tex pyramid 5 levels", "tex", {}, levels=5)
    run("tma pyramid 5 levels", "tma", {}, levels=5)
    run("tex pyramid 5 levels ctas=2", "tex", {"MDC_CTAS_PER_SM": 2}, levels=5)


if __name__ == "__main__":
    main()
