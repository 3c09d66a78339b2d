"""Time the E-step kernel (n=1000 x 1 MP, or --npix N) under the current MDC_* environment; prints one JSON line."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_dataset_code_b200 import api

def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(5)
    n, npix = 1000, 1000 * 1000
    if "--npix" in sys.argv:
        npix = int(sys.argv[sys.argv.index("--npix") + 1])
    data = torch.randint(0, 256, (n, npix), dtype=torch.uint8, device=dev, generator=g)
    t = torch.linspace(0.05, 20.0, n, dtype=torch.float64, device=dev)
    G = torch.linspace(0.0, 255.0, 256, dtype=torch.float64, device=dev)
    E = torch.empty(npix, dtype=torch.float64, device=dev)
    ctx = api.Context(None, None, 0)
    for _ in range(3):
        ctx.estep(data, t, G, E)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record(); ctx.estep(data, t, G, E); b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    alg = n * npix + 8 * npix

    def time_it(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    G2 = G.clone()
    gstep_ms = time_it(lambda: ctx.rc_gstep(data, t, E, G2))
    rmse_ms = time_it(lambda: ctx.rc_rmse(data, t, G, E))
    einit_ms = time_it(lambda: ctx.rc_einit(data, E.clone()))
    import time
    E3, G3 = torch.empty_like(E), torch.zeros_like(G)
    ctx.response_calib(data, t, 1, E3, G3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.response_calib(data, t, 5, E3, G3)
    torch.cuda.synchronize(); calib_ms = (time.perf_counter() - t0) * 1e3 / 5
    print(json.dumps({"npix": npix, "ms": float(np.mean(ms)), "min_ms": float(np.min(ms)), "gbs": alg / (np.mean(ms) * 1e-3) / 1e9,
                      "checksum": float(torch.nansum(E).item()), "gstep_ms": gstep_ms, "rmse_ms": rmse_ms, "einit_ms": einit_ms, "calib_ms_per_iteration": calib_ms,
                      "gstep_checksum": float(torch.nansum(G2).item()),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("MDC_")}}), flush=True)

if __name__ == "__main__":
    main()
