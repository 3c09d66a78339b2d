"""E-step / G-step / rmse time against the tile width (consumer warps per CTA, MDC_STREAM_WARPS) for several image sizes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_dataset_code_b200 import api

dev = torch.device("cuda:0")
ctx = api.Context(None, None, 0)
n = 1000
t = torch.linspace(0.05, 20.0, n, dtype=torch.float64, device=dev)
G = torch.linspace(0.0, 255.0, 256, dtype=torch.float64, device=dev)
for npix in (1000000, 500096, 250112, 125056):
    g = torch.Generator(device=dev); g.manual_seed(5)
    data = torch.randint(0, 256, (n, npix), dtype=torch.uint8, device=dev, generator=g)
    E = torch.empty(npix, dtype=torch.float64, device=dev)
    G2 = G.clone()
    for w in (0, 1, 2, 3, 4, 6, 8, 10, 12, 14):
        if w:
            os.environ["MDC_STREAM_WARPS"] = str(w)
        else:
            os.environ.pop("MDC_STREAM_WARPS", None)
        out = {"npix": npix, "warps": w or "auto"}
        for name, fn in (("estep_ms", lambda: ctx.estep(data, t, G, E)), ("gstep_ms", lambda: ctx.rc_gstep(data, t, E, G2)),
                         ("rmse_ms", lambda: ctx.rc_rmse(data, t, G, E))):
            fn(); fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record(); torch.cuda.synchronize()
            out[name] = round(a.elapsed_time(b) / 5, 4)
        print(json.dumps(out), flush=True)
    del data, E
