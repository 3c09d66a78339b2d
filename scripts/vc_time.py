"""Time the vignetteCalib optimiser kernels (plane step, vignette step, smoothing) on a synthetic device-resident problem:
n images of 1280x1024, 1000x1000 plane grid, smooth affine plane-to-image maps.  Prints one JSON line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_dataset_code_b200 import api


def main():
    n = int(os.environ.get("VC_N", "256"))
    gw = gh = 1000
    wI, hI = 1280, 1024
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1)
    gy, gx = torch.meshgrid(torch.arange(gh, device=dev, dtype=torch.float32), torch.arange(gw, device=dev, dtype=torch.float32), indexing="ij")
    u, v = (gx / gw - 0.5).ravel(), (gy / gh - 0.5).ravel()
    p2x = torch.empty((n, gw * gh), dtype=torch.float32, device=dev)
    p2y = torch.empty_like(p2x)
    for i in range(n):
        ang, sc = 0.4 * np.sin(i), 500.0 + 250.0 * np.cos(0.37 * i)
        tx, ty = 640 + 150 * np.sin(0.11 * i), 512 + 120 * np.cos(0.23 * i)
        X = tx + sc * (np.cos(ang) * u - np.sin(ang) * v)
        Y = ty + sc * (np.sin(ang) * u + np.cos(ang) * v)
        bad = ~(((X + 0.5).int() > 1) & ((Y + 0.5).int() > 1) & ((X + 0.5).int() < wI - 2) & ((Y + 0.5).int() < hI - 2))
        X[bad] = float("nan"); Y[bad] = float("nan")
        p2x[i], p2y[i] = X, Y
    images = torch.rand((n, wI * hI), dtype=torch.float32, device=dev, generator=g) * 200 + 20
    plane = torch.full((gw * gh,), 100.0, dtype=torch.float32, device=dev)
    vig = torch.ones(wI * hI, dtype=torch.float32, device=dev)
    ctx = api.Context(None, None, 0)
    visible = float(torch.isfinite(p2x).float().mean().item())

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps

    plane_ms = timed(lambda: ctx.vc_plane_step(images, p2x, p2y, gw, gh, wI, hI, vig, plane, 10000 * 10000))
    vig_ms = timed(lambda: ctx.vc_vignette_step(images, p2x, p2y, gw, gh, wI, hI, plane, vig, 10000 * 10000))
    smooth_ms = timed(lambda: ctx.vc_smooth(vig, wI, hI, 4))
    samples = n * gw * gh
    pc2, v2 = plane.clone(), vig.clone()
    ctx.vignette_calib(images, p2x, p2y, gw, gh, wI, hI, 1, 15, pc2, v2, True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.vignette_calib(images, p2x, p2y, gw, gh, wI, hI, 4, 15, pc2, v2, True)
    torch.cuda.synchronize()
    loop_ms = (time.perf_counter() - t0) * 1e3 / 4      # incl. the one-off bounds pass and the final smoothing
    print(json.dumps({"n": n, "plane_points": gw * gh, "visible_fraction": visible, "plane_step_ms": plane_ms, "vignette_step_ms": vig_ms,
                      "smooth4_ms": smooth_ms, "loop_ms_per_iteration": loop_ms, "plane_gsamples_per_s": samples / plane_ms / 1e6, "vignette_gsamples_per_s": samples / vig_ms / 1e6,
                      "map_bytes_gbs_plane": samples * 8 / plane_ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
