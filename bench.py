#!/usr/bin/env python
"""bench.py — frames/s of the per-frame image-preparation hot path (BASELINE.json metric).

Step   = one pass of the fused kernel K1 (photometric un-map + FOV rectification) over one batch
         of BATCH synthetic 1280x1024 mono8 frames per GPU (BASELINE.json configs[1]; the batch
         size and the 5-level pyramid of configs[2] are reported as the `c3_pyramid` extra).
value  = whole-job frames/s with the batch already resident in HBM (CUDA events, max over ranks).
e2e    = the same metric through the host-buffer C-ABI call (pinned host frames in, float images
         out; H2D + kernel + D2H inside the timed region).
roofline = algorithmic bytes of one K1 launch / its CUDA-event duration, vs MEASURED_PEAKS.json.
cpu_baseline = the reference's own CPU code (oracle/_ref, compiled from /root/reference) — or the
         C restatement when that is absent — timed on this box's host cores on a bounded sample.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched under torchrun)
  python bench.py --impl reference ...                            (CPU reference arm)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IN_W, IN_H, OUT_W, OUT_H = 1280, 1024, 1280, 1024
ALG_BYTES_PER_FRAME = IN_W * IN_H * 1 + OUT_W * OUT_H * 4                    # 6 553 600
PYR_EXTRA_BYTES = sum((OUT_W >> l) * (OUT_H >> l) * 4 for l in range(1, 5))    # 1 740 800


def set_geometry(w, h):
    """--geom WxH (tuning / other BASELINE configs, e.g. 1920x1080 = configs[3]); the default is configs[1]."""
    global IN_W, IN_H, OUT_W, OUT_H, ALG_BYTES_PER_FRAME, PYR_EXTRA_BYTES
    IN_W, IN_H, OUT_W, OUT_H = w, h, w, h
    ALG_BYTES_PER_FRAME = w * h + w * h * 4
    PYR_EXTRA_BYTES = sum((w >> l) * (h >> l) * 4 for l in range(1, 5))
FLAGS_ALL = 1 | 2 | 4        # rectify + removeGamma + removeVignette (the reference viewer's full correction)
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback
CPU_FRAMES_PER_THREAD = 16   # frames per worker and step of the CPU arm (also the length of its thread-count sweep runs)
ORIG_AFFINITY = None         # the process's CPU set before the GPU arm bound it to its GPU's NUMA node


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name):
    """Per-launch DRAM traffic of a kernel from a committed `ncu --set full` capture (profiles/<name>, written by
    scripts/ncu_summary.py) — only if it was captured from the kernel sources that are being benchmarked: the file carries their
    SHA-256, and a capture of older sources is reported as absent rather than passed off as a measurement of this binary."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            doc = json.load(f)
        h = hashlib.sha256()
        for rel in doc["kernel_sources"]:
            with open(os.path.join(ROOT, rel), "rb") as src:
                h.update(src.read())
        if h.hexdigest() != doc["kernel_sources_sha256"]:
            return None, f"profiles/{name} was captured at commit {doc.get('captured_at_commit')} from different kernel sources: stale, not reported"
        return doc, doc["note"] + f" [captured at commit {doc.get('captured_at_commit')}, kernel sources unchanged since]"
    except Exception as exc:
        return None, f"no usable ncu capture ({exc.__class__.__name__})"


def write_calibration(tmp):
    from mono_dataset_code_b200 import synthetic as S
    return S.write_dataset_dir(tmp, IN_W, IN_H, OUT_W, OUT_H, "crop")


class ClockSampler:
    """SM clock + throttle reasons polled through NVML (nvidia_ml_py) every few ms DURING the timed region;
    falls back to one nvidia-smi query if NVML is unavailable."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.gpu, self.samples, self.reason_bits, self._stop, self.t, self.h = gpu_index, [], 0, False, None, None
        self.max_mhz = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self._stop = False
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=20).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["nvml unavailable: single nvidia-smi sample after the run"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"], "samples": 0}
        self._stop = True
        self.t.join()
        reasons = [n for n, b in {**self.BAD, **self.NOTE}.items() if self.reason_bits & b]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU arms
def workload_name():
    """One string for both arms (the driver compares the two `config` objects)."""
    return (f"{IN_W}x{IN_H} mono8 -> {OUT_W}x{OUT_H} f32: GInv[I]*vignetteInv + FOV crop remap per frame "
            "(unMapImage + undistort<float>, BASELINE configs[1])")


def config_dict(batch=256):
    """`config` of the JSON line — the same object on both arms (what differs between them is in `run_config`)."""
    return {"workload": workload_name(), "flags": "rectify|removeGamma|removeVignette", "pyramid_levels": 1,
            "frames": "uniform random mono8, 16 distinct (CPU arm) / one resident batch per GPU (GPU arm)",
            "l2_policy": f"GPU arm: inputs larger than L2 ({batch * IN_W * IN_H >> 20} MiB in, {batch * OUT_W * OUT_H * 4 >> 20} MiB out per step of {batch} frames); "
                         "CPU arm: per-thread outputs + shared inputs exceed the last-level cache"}


def cpu_quota_cores():
    """CPU time the container may use per wall-clock second (cgroup CFS quota / period), or None if unlimited / unknown.  The GPU boxes of
    this pool give a 1-GPU job 16 CPUs' worth although 128 hardware threads are visible: more busy threads than that only get throttled."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                       # cgroup v2
            q, per = f.read().split()[:2]
        return None if q == "max" else float(q) / float(per)
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:      # cgroup v1
            q, per = float(f.read()), float(g.read())
        return None if q <= 0 else q / per
    except (OSError, ValueError):
        return None


class CpuReference:
    """The reference's own per-frame code on the host cores (oracle/_ref = /root/reference/src/*.cpp compiled unmodified), or the
    plain-C restatement when that build is absent.  Two figures:
      single_thread  the reference as shipped: one thread, unMapImage -> undistort<float> per frame (BenchmarkDatasetReader.h:222-223)
      all_cores      the same loop on every hardware thread: persistent pinned workers, private first-touched buffers, one replica
                     of the tables per NUMA node; timed inside the library between a start barrier and the last worker's finish."""

    def __init__(self, files, threads=0):
        from oracle import loader
        from mono_dataset_code_b200 import synthetic as S
        self.frames = S.frames(16, IN_W, IN_H)
        self.cores = os.cpu_count() or 1
        self.quota = cpu_quota_cores()
        self.pool = None
        if loader.ref_available():
            self.kind = "reference"
            self.R = R = loader.RefOracle()
            R.register_image(files["vignette"], files["vignette_pixels"])
            self.fov = R.fov(files["camera"])
            self.photo = R.photo(files["pcalib"], files["vignette"], IN_W, IN_H)
            assert self.fov.valid and self.photo.valid_vignette
            # The reference's loop does not scale to every hardware thread: beyond a few dozen workers the private
            # float images (2 x 5 MB per thread) evict each other and throughput FALLS (profiles/r02_cpu_ref_scaling.jsonl),
            # and the GPU boxes cap the container at 16 CPUs per GPU (cgroup quota) although 128 threads are visible.
            # The CPU arm therefore runs at the best worker count / placement of a short sweep — its fastest configuration.
            self.sweep = []
            counts = {8, 16, 32, 64, self.cores}
            if self.quota:                 # busy threads beyond the container's CPU quota only get throttled: sweep up to the quota
                q = max(1, int(self.quota))
                counts = {max(1, q // 4), max(1, q // 2), q}
            cands = [(threads, "0")] if threads else [(t, sp) for t in sorted(counts) if t <= self.cores for sp in ("0", "1")]
            best = None
            for t, spread in cands:
                os.environ["MDC_REF_SPREAD"] = spread
                pool = loader.RefPool(R, files["camera"], files["pcalib"], files["vignette"], IN_W, IN_H, self.frames, t, (1, 1, 0))
                pool.run(2)
                s = pool.run(CPU_FRAMES_PER_THREAD)[0]      # as long as a timed step: short bursts overstate the steady rate by up to 2x
                fps = t * CPU_FRAMES_PER_THREAD / s
                self.sweep.append({"threads": t, "spread_over_numa_nodes": spread == "1", "frames_per_s": round(fps, 1)})
                if best is None or fps > best[0]:
                    if best is not None:
                        best[1].close()
                    best = (fps, pool, spread)
                else:
                    pool.close()
            self.pool = best[1]
            self.placement = "spread over NUMA nodes" if best[2] == "1" else "compact"
            self.threads, self.numa_nodes = self.pool.threads, self.pool.numa_nodes
        else:
            self.kind = "port"
            self.P = P = loader.PortOracle()
            f = P.fov_from_file(files["camera"])
            self.rx, self.ry = f.tables()
            self.ginv, _ = P.photo_tables(np.loadtxt(files["pcalib"], dtype=np.float32))
            _, vinv = P.vignette_maps(files["vignette_pixels"])
            self.vinv = vinv.reshape(-1)
            self.threads, self.numa_nodes = (threads or self.cores), None

    def single_thread(self, n=150):
        if self.kind == "reference":
            self.R.time_frames(self.fov, self.photo, self.frames, 8, 1, (1, 1, 0))
            s, st = self.R.time_frames(self.fov, self.photo, self.frames, n, 1, (1, 1, 0))
            return {"value": n / s, "frames": n, "unmap_ms": 1e3 * st[0] / n, "undistort_ms": 1e3 * st[1] / n}
        self.P.time_frames(self.rx, self.ry, IN_W, IN_H, OUT_W, OUT_H, self.ginv, self.vinv, self.frames, 8, 1, 3, 1)
        s = self.P.time_frames(self.rx, self.ry, IN_W, IN_H, OUT_W, OUT_H, self.ginv, self.vinv, self.frames, n, 1, 3, 1)
        return {"value": n / s, "frames": n}

    def all_cores(self, frames_per_thread):
        """(frames, seconds) of one all-core pass."""
        if self.pool is not None:
            s, _ = self.pool.run(frames_per_thread)
            return self.threads * frames_per_thread, s
        n = self.threads * frames_per_thread
        return n, self.P.time_frames(self.rx, self.ry, IN_W, IN_H, OUT_W, OUT_H, self.ginv, self.vinv, self.frames, n, self.threads, 3, 1)

    def describe(self):
        return (f"{self.threads} pinned threads" + (f" ({self.placement}) on {self.numa_nodes} NUMA node(s), tables + inputs replicated per node" if self.numa_nodes else "")
                + f" = the fastest of a sweep over worker counts on this {self.cores}-thread host"
                + (f" (container CPU quota: {self.quota:g} CPUs)" if self.quota else "") + "; private buffers first-touched by their thread; "
                "timed inside the library; decode/alloc excluded")

    def close(self):
        if self.pool is not None:
            self.pool.close()


def cpu_baseline(files):
    ref = CpuReference(files)
    one = ref.single_thread()
    ref.all_cores(2)                                   # warm-up: page in, spin up
    fpt, n, s = CPU_FRAMES_PER_THREAD, 0, 0.0
    for _ in range(5):                                 # the same step the --impl reference arm times
        a, b = ref.all_cores(fpt)
        n += a
        s += b
    out = {"value": n / s, "unit": "frames/s", "cores": ref.threads, "kind": ref.kind,
           "sample": f"{n} frames of {IN_W}x{IN_H} (16 distinct, cycled), 5 steps of {fpt} per thread; " + ref.describe(),
           "single_thread": one, "thread_sweep": getattr(ref, "sweep", None), "cpu_quota_cores": ref.quota, "hardware_threads": ref.cores}
    ref.close()
    return out


def cpu_calibrator_sample():
    """The oracle's plain-C restatement of the responseCalib passes (main_responseCalib.cpp:283-346, one thread, like the reference)
    on a bounded sample: n = 1000 exposures x 20 000 pixels, scaled to the 1 MP of BASELINE configs[4]."""
    from oracle import loader
    port = loader.PortOracle()
    rng = np.random.default_rng(7)
    n, npix, full = 1000, 20000, 1000 * 1000
    data = rng.integers(0, 256, (n, npix), dtype=np.uint8)
    t = np.linspace(0.05, 20.0, n)
    G = np.linspace(0.0, 255.0, 256)
    t0 = time.perf_counter(); E = port.estep(data, t, G); t1 = time.perf_counter()
    port.gstep(data, t, E); t2 = time.perf_counter()
    port.rmse(data, t, G, E); t3 = time.perf_counter()
    k = full / npix
    return {"kind": "port", "cores": 1, "sample": f"n={n} x {npix} pixels, scaled x{k:.0f} to 1 MP",
            "estep_ms_per_pass": 1e3 * (t1 - t0) * k, "gstep_ms_per_pass": 1e3 * (t2 - t1) * k, "rmse_ms_per_pass": 1e3 * (t3 - t2) * k}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="mdc_bench_")
    files = write_calibration(tmp)
    ref = CpuReference(files)
    one = ref.single_thread()
    fpt = CPU_FRAMES_PER_THREAD
    for _ in range(max(args.warmup, 1)):
        ref.all_cores(2)
    frames = secs = 0.0
    for _ in range(args.steps):
        n, s = ref.all_cores(fpt)
        frames += n
        secs += s
    fps = frames / secs
    per_step = ref.threads * fpt
    line = {"impl": "reference", "metric": "frames_per_s_1280x1024_photometric_fov_undistort", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (uniform random mono8 frames, synthetic TUM-style calibration; SURVEY.md §8d)",
            "config": config_dict(),
            "run_config": {"frames_per_step": per_step, "threads": ref.threads, "numa_nodes": ref.numa_nodes},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ref.threads, "kind": ref.kind,
                             "sample": f"{per_step} frames/step x {args.steps} steps ({fpt} per thread and step); " + ref.describe(),
                             "single_thread": one, "cpu_quota_cores": ref.quota, "hardware_threads": ref.cores},
            "as_shipped_single_thread": one, "thread_sweep": getattr(ref, "sweep", None),
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    ref.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(local):
    """Run this rank (and every thread it starts later) on the CPUs of the NUMA node its GPU hangs off, so that pinned staging
    buffers, decode threads and the copy engine's host side are local to the GPU's PCIe root (8-GPU e2e 0.55)."""
    from mono_dataset_code_b200 import _lib
    info = {"gpu": local, "node": None, "cpus_bound": None}
    global ORIG_AFFINITY
    ORIG_AFFINITY = os.sched_getaffinity(0)
    if os.environ.get("MDC_NUMA_BIND", "1") == "0":
        info["disabled"] = True
        return info
    node = int(_lib.lib.mdc_device_numa_node(local))
    if node < 0:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(node=node, cpus_bound=len(cpus))
    except OSError:
        pass
    return info


def pcie_ceiling(dev, n_bytes=256 << 20, reps=5):
    """Plain pinned-memory copies of the e2e leg's size class: GB/s host->device and device->host (CUDA events)."""
    import torch
    host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    devb = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in {"h2d_gbs": (devb, host), "d2h_gbs": (host, devb)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        out[name] = reps * n_bytes / (a.elapsed_time(b) * 1e-3) / 1e9
    return out


def max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_estep_leg(args, ctx, comm, dev, rank, world, peak, barrier):
    """BASELINE configs[4] at N > 1: responseCalib over n = 1000 exposures x 1 MP with the image stack split by PIXEL RANGE over the
    ranks: E-step, G-step accumulation and rmse run on the local slice; per loop iteration one 256-double
    all-reduce (G-step) and three 2-double all-reduces (rmse).  Strong scaling: the 1 MP problem is fixed, each rank holds 1/N of it."""
    import torch
    import torch.distributed as dist
    from mono_dataset_code_b200 import sharding
    n_img, npix = 1000, 1000 * 1000
    lo, hi = sharding.shard_pixels(npix, rank, world)
    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    data = torch.randint(0, 256, (n_img, hi - lo), dtype=torch.uint8, device=dev, generator=g)
    t_exp = torch.linspace(0.05, 20.0, n_img, dtype=torch.float64, device=dev)
    G_tab = torch.linspace(0.0, 255.0, 256, dtype=torch.float64, device=dev)
    E = torch.empty(hi - lo, dtype=torch.float64, device=dev)

    def ms_of(fn, reps=5):
        fn(); barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return max_over_ranks(a.elapsed_time(b) / reps, dev, world)
    est_ms = ms_of(lambda: ctx.estep(data, t_exp, G_tab, E))
    scale4, limbs = torch.zeros(4, dtype=torch.int64, device=dev), torch.zeros(768, dtype=torch.int64, device=dev)
    special = torch.zeros(256, dtype=torch.float64, device=dev)
    gnum = torch.zeros(256, dtype=torch.int64, device=dev)
    G_new = torch.zeros_like(G_tab)

    def gstep():      # sums exact across ranks: scale (MAX), integer limbs + side sums + counts (SUM), finish
        ctx.rc_gstep_scale(E, t_exp, scale4)
        dist.all_reduce(scale4, op=dist.ReduceOp.MAX)
        ctx.rc_gstep_accumulate_exact(data, t_exp, E, scale4, limbs, special, gnum, False)
        dist.all_reduce(limbs); dist.all_reduce(special); dist.all_reduce(gnum)
        ctx.rc_gstep_finish_exact(scale4, limbs, special, gnum, G_new)
    g_ms = ms_of(gstep, 3)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)

    def rmse():
        ctx.rc_rmse_accumulate(data, t_exp, G_tab, E, acc)
        dist.all_reduce(acc)
    r_ms = ms_of(rmse, 3)
    # one whole iteration of the loop (G-step, E-step, rescale, 3 x rmse) in C++ on the native communicator
    loop_ms, how = None, "python host loop over torch.distributed"
    nits = 3
    E2, G2 = torch.zeros_like(E), torch.zeros_like(G_tab)
    if comm is not None:                  # first use of the communicator for all-reduces: NCCL connects its channels lazily (tens of ms)
        comm.response_calib_sharded(ctx, data, t_exp, 1, E2, G2)
    else:
        sharding.response_calib_sharded(ctx, data, t_exp, 1, E2, G2)
    barrier()
    t0 = time.perf_counter()
    if comm is not None:
        comm.response_calib_sharded(ctx, data, t_exp, nits, E2, G2)
        how = "mdc_response_calib_sharded (C++, ncclAllReduce in libmdc_b200_nccl.so)"
    else:
        sharding.response_calib_sharded(ctx, data, t_exp, nits, E2, G2)
    torch.cuda.synchronize()
    loop_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / nits, dev, world)
    # G must be identical on every rank
    Gs = [torch.zeros_like(G2) for _ in range(world)]
    dist.all_gather(Gs, G2)
    same = all(bool(torch.equal(Gs[0], x) or torch.equal(torch.nan_to_num(Gs[0]), torch.nan_to_num(x))) for x in Gs[1:])
    alg = n_img * npix + 8 * npix
    return {"ms_per_pass": est_ms, "algorithmic_bytes": alg, "achieved_gbs": alg / (est_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": alg / (est_ms * 1e-3) / 1e9 / (peak * world), "workload": "n=1000 x 1 MP u8 -> f64 E[1 MP]",
            "sharding": f"pixel-sharded over {world} ranks ({hi - lo} pixels on rank 0), strong scaling", "gstep_ms": g_ms, "rmse_ms": r_ms,
            "loop_iteration_ms": loop_ms, "loop": how, "collectives_per_iteration": "G-step: allreduce MAX(4 u64), SUM(768 i64), SUM(256 f64) [+ SUM(256 u64) once]; 3 x allreduce SUM(2 f64) for rmse",
            "G_identical_on_all_ranks": same}


def sequence_leg(args, dev, rank, world, local, numa, barrier):
    """BASELINE configs[3]: a 1920x1080 sequence of --seq-frames frames, frame-sharded over the ranks (sharding.shard_range).
    decode_inclusive: DatasetReader's path — JPEG entries of images.zip decoded on the host threads of the rank's NUMA node,
    H2D, K1, D2H into pinned float images (mdc_seq_prepare; BenchmarkDatasetReader.h:247-276 + :188-243).
    device_resident: the same number of frames through K1 alone at this geometry (batches of 256 already in HBM)."""
    import ctypes as C
    import zipfile
    import torch
    import torch.distributed as dist
    import cv2
    from mono_dataset_code_b200 import _lib, api, sharding, synthetic as S
    W, H, K, UNIQUE, CALL = 1920, 1080, 1024, 128, 512      # zip entries, distinct JPEG payloads among them, frames per mdc_seq_prepare call
    root = os.path.join(tempfile.gettempdir(), f"mdc_c4_{os.environ.get('MASTER_PORT', 'solo')}_{os.getppid() if world > 1 else os.getpid()}")
    if rank == 0:
        os.makedirs(root, exist_ok=True)
        files = S.write_dataset_dir(root, W, H, W, H, "crop")
        sizes = []
        with zipfile.ZipFile(os.path.join(root, "images.zip"), "w", zipfile.ZIP_STORED) as z:
            yy, xx = np.mgrid[0:H, 0:W]
            rng = np.random.default_rng(11)
            blobs = []
            for i in range(UNIQUE):     # a textured, slowly changing scene with mild sensor noise (JPEG size / decode cost of a real sequence, not of white noise)
                img = (xx * (150.0 / W) + yy * (60.0 / H) + 40 * np.sin((xx + 13 * i) * 0.05) * np.cos((yy - 7 * i) * 0.04)
                       + ((xx // 64 + yy // 64 + i) % 2) * 30 + rng.normal(0, 1.5, (H, W)))
                img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
                ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
                blobs.append(enc.tobytes())
                sizes.append(len(enc))
            for i in range(K):
                z.writestr(f"{i:05d}.jpg", blobs[i % UNIQUE])
        with open(os.path.join(root, "times.txt"), "w") as f:
            for i in range(K):
                f.write(f"{i} {i * 0.05:.3f} 1.0\n")
        meta = {"root": root, "jpeg_mean_bytes": float(np.mean(sizes))}
    else:
        meta = None
    if world > 1:
        box = [meta]
        dist.broadcast_object_list(box, src=0)
        meta = box[0]
    root = meta["root"]
    fov = api.UndistorterFOV(os.path.join(root, "camera.txt"))
    photo = api.PhotometricUndistorter(os.path.join(root, "pcalib.txt"), os.path.join(root, "vignette.png"), W, H)
    ctx = api.Context(fov, photo, local)
    seq = api.Sequence(root + "/")
    assert seq.getNumImages() == K
    begin, end = sharding.shard_range(args.seq_frames, rank, world)
    n_mine = end - begin
    # host threads: this rank's share of the CPUs it is bound to (ranks on the same NUMA node split them)
    nodes = [numa]
    if world > 1:
        nodes = [None] * world
        dist.all_gather_object(nodes, numa)
    same_node = sum(1 for x in nodes if x.get("node") == numa.get("node"))
    threads = max(1, len(os.sched_getaffinity(0)) // max(1, same_node))
    quota = cpu_quota_cores()
    if quota:                              # busy threads beyond the container's CPU quota only get throttled (and make every timing noisy)
        threads = max(1, min(threads, int(quota // world)))
    n_px = W * H
    h_out = C.c_void_p()
    _lib.check(_lib.lib.mdc_host_alloc(C.byref(h_out), CALL * n_px * 4), "mdc_host_alloc")
    ptrs = (C.c_void_p * 1)(h_out.value)

    def run_decode(n_frames):
        v, done = begin, 0
        while done < n_frames:
            first = v % K
            cnt = min(K - first, n_frames - done, CALL)
            _lib.check(_lib.lib.mdc_seq_prepare(ctx._h, seq._h, first, cnt, FLAGS_ALL, ptrs, 1, threads), "mdc_seq_prepare")
            v += cnt; done += cnt
    run_decode(min(n_mine, CALL))         # warm-up: page cache, thread pool, pinned staging
    # more decode threads are not always faster on these hosts (SMT siblings, shared memory bandwidth): short sweep, keep the best
    sweep, cap = [], threads
    for cand in sorted({max(1, cap // 4), max(1, cap // 2), cap}):
        threads = cand
        t0 = time.perf_counter()
        run_decode(min(n_mine, 2 * CALL))
        sweep.append({"threads": cand, "frames_per_s_this_rank": round(min(n_mine, 2 * CALL) / (time.perf_counter() - t0), 1)})
    threads = max(sweep, key=lambda r: r["frames_per_s_this_rank"])["threads"]
    # what the host side alone delivers at that thread count (zip read + JPEG decode into per-thread buffers, no GPU work)
    from concurrent.futures import ThreadPoolExecutor
    bufs = [np.empty(n_px, np.uint8) for _ in range(threads)]

    def decode_some(t):
        w_, h_ = C.c_int(), C.c_int()
        for i in range(t, 2 * CALL, threads):
            _lib.lib.mdc_seq_read_gray8(seq._h, i % K, bufs[t].ctypes.data_as(C.c_void_p), n_px, C.byref(w_), C.byref(h_))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(decode_some, range(threads)))
    decode_only = 2 * CALL / (time.perf_counter() - t0)
    if world > 1:                          # all ranks use the same count (the slowest rank sets the job's rate anyway)
        box = [threads]
        dist.broadcast_object_list(box, src=0)
        threads = box[0]
    barrier()
    t0 = time.perf_counter()
    run_decode(n_mine)
    barrier()
    dec_s = max_over_ranks(time.perf_counter() - t0, dev, world)
    _lib.lib.mdc_host_free(h_out)
    # device-resident: the same frame count through K1 at this geometry
    B = 256
    g = torch.Generator(device=dev)
    g.manual_seed(77 + rank)
    frames = torch.randint(0, 256, (B, n_px), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty((B, n_px), dtype=torch.float32, device=dev)
    launches = (n_mine + B - 1) // B
    for _ in range(3):
        ctx.prepare_batch(frames, FLAGS_ALL, [out])
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(launches):
        nb = min(B, n_mine - i * B)
        ctx.prepare_batch(frames[:nb], FLAGS_ALL, [out[:nb]])
    b.record()
    torch.cuda.synchronize()
    res_s = max_over_ranks(a.elapsed_time(b) * 1e-3, dev, world)
    ctx.close(); seq.close()
    if rank == 0:
        try:
            import shutil
            shutil.rmtree(root, ignore_errors=True)
        except Exception:
            pass
    total = args.seq_frames
    alg = total * (n_px + 4 * n_px)
    return {"workload": f"{W}x{H} mono8 sequence of {total} frames ({K} baseline-JPEG entries of images.zip, {UNIQUE} distinct, cycled; {CALL} frames per mdc_seq_prepare call), frame-sharded over {world} rank(s) "
                        "with shard_range; rectify|removeGamma|removeVignette",
            "decode_inclusive": {"value": total / dec_s, "unit": "frames/s", "seconds": dec_s, "decode_threads_per_rank": threads,
                                 "decode_thread_sweep": sweep, "host_decode_only_frames_per_s_rank0": decode_only,
                                 "d2h_bytes_per_frame": n_px * 4, "numa": nodes, "jpeg_mean_bytes": meta["jpeg_mean_bytes"],
                                 "cpu_quota_cores": cpu_quota_cores(),
                                 "path": "mdc_seq_prepare: zip read + JPEG decode (host) -> pinned -> H2D -> K1 -> D2H, chunks double-buffered"},
            "device_resident": {"value": total / res_s, "unit": "frames/s", "seconds": res_s, "alg_gbs": alg / res_s / 1e9,
                                "frames_per_launch": B, "launches_per_rank": launches}}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from mono_dataset_code_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)       # before any pinned buffer or worker thread exists
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- calibration: rank 0 parses/builds on the host, tables are broadcast over NCCL, every rank adopts them
    tmp = tempfile.mkdtemp(prefix="mdc_bench_")
    files = write_calibration(tmp) if rank == 0 else None
    n_in, n_out = IN_W * IN_H, OUT_W * OUT_H
    fov = photo = None
    if rank == 0:
        fov = api.UndistorterFOV(files["camera"])
        photo = api.PhotometricUndistorter(files["pcalib"], files["vignette"], IN_W, IN_H)
    comm, init_how = None, "single GPU: tables uploaded by mdc_ctx_create"
    if world == 1:
        ctx = api.Context(fov, photo, local)
    else:
        from mono_dataset_code_b200 import sharding
        try:
            # the C++ path (include/mdc_b200_nccl.h): the process group only carries the 128-byte NCCL id; the four tables travel with
            # ncclBroadcast inside libmdc_b200_nccl.so and every rank's context adopts its copy — no collective in steady state
            comm = sharding.NativeComm(local)
            ctx = comm.create_context(fov, photo)
            init_how = f"native: mdc_ctx_create_broadcast (libmdc_b200_nccl.so, ncclBroadcast x4, NCCL {comm.version})"
        except (ImportError, OSError) as exc:
            dims, tabs = sharding.broadcast_calibration(fov, photo, dev)     # same broadcast through torch.distributed
            torch.cuda.synchronize()
            ctx = api.Context.from_device_tables(local, *dims, *tabs)
            init_how = f"torch.distributed broadcast (native helper unavailable: {exc})"
    if args.tma is not None:
        ctx.configure(use_tma=args.tma)

    B = args.batch
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    frames = torch.randint(0, 256, (B, n_in), dtype=torch.uint8, device=dev, generator=g)
    lvl_px = [(OUT_W >> l) * (OUT_H >> l) for l in range(5)]
    outs = [torch.empty((B, px), dtype=torch.float32, device=dev) for px in lvl_px]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(levels, steps, warmup, sampler=None):
        for _ in range(warmup):
            ctx.prepare_batch(frames, FLAGS_ALL, outs[:levels])
        barrier()
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        t_all0.record()
        for a, b in ev:
            a.record()
            ctx.prepare_batch(frames, FLAGS_ALL, outs[:levels])
            b.record()
        t_all1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        total_ms = t_all0.elapsed_time(t_all1)
        per_launch_ms = [a.elapsed_time(b) for a, b in ev]
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, per_launch_ms, ctx.launch_count - l0, clocks

    TRAFFIC, TRAFFIC_NOTE = ncu_traffic("k1_traffic.json")
    sampler = ClockSampler(local) if rank == 0 else None
    total_ms, per_launch_ms, launches, clocks = timed(args.levels if args.only_kernel else 1, args.steps, args.warmup, sampler)
    value = world * B * args.steps / (total_ms * 1e-3)
    peak, peak_src = measured_peak()
    k1_ms = float(np.mean(per_launch_ms))
    achieved = B * ALG_BYTES_PER_FRAME / (k1_ms * 1e-3) / 1e9

    if args.only_kernel:
        if rank == 0:
            print(json.dumps({"value": value, "achieved_gbs": achieved, "frac": achieved / peak, "launch_ms": k1_ms, "clocks": clocks,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("MDC_")}, "tma": args.tma}), flush=True)
        return

    # configs[2]: + 5-level pyramid fused in the same kernel's epilogue
    p_total_ms, p_launch_ms, _, _ = timed(5, max(2, args.steps // 2), 2)
    p_steps = max(2, args.steps // 2)
    pyr_traffic, pyr_note = ncu_traffic("c3_pyramid_traffic.json")
    pyr = {"value": world * B * p_steps / (p_total_ms * 1e-3), "unit": "frames/s", "levels": 5,
           "traffic": (pyr_traffic["dram_bytes_per_frame"] * B / 1e9 if pyr_traffic else None), "traffic_note": pyr_note,
           "achieved_gbs": B * (ALG_BYTES_PER_FRAME + PYR_EXTRA_BYTES) / (float(np.mean(p_launch_ms)) * 1e-3) / 1e9}

    # ---- BASELINE configs[4]: responseCalib E-step, 1000 exposures x 1 MP, fp64, bit-exact kernel K3 (1 GPU only)
    estep = None
    if world > 1 and not args.no_estep:
        estep = sharded_estep_leg(args, ctx, comm, dev, rank, world, peak, barrier)
    if world == 1 and not args.no_estep:
        n_img, npix = 1000, 1000 * 1000
        data = torch.randint(0, 256, (n_img, npix), dtype=torch.uint8, device=dev, generator=g)
        t_exp = torch.linspace(0.05, 20.0, n_img, dtype=torch.float64, device=dev)
        G_tab = torch.linspace(0.0, 255.0, 256, dtype=torch.float64, device=dev)
        E_out = torch.empty(npix, dtype=torch.float64, device=dev)
        for _ in range(2):
            ctx.estep(data, t_exp, G_tab, E_out)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in evs:
            a.record(); ctx.estep(data, t_exp, G_tab, E_out); b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        alg = n_img * npix + 8 * npix
        es_traffic, es_note = ncu_traffic("c5_estep_traffic.json")
        estep = {"ms_per_pass": ms, "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                 "traffic": (es_traffic["dram_bytes_read_per_launch"] + es_traffic["dram_bytes_write_per_launch"]) / 1e9 if es_traffic else None,
                 "traffic_note": es_note,
                 "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "workload": "n=1000 x 1 MP u8 -> f64 E[1 MP]"}
        # the other passes of the calibrator over the same stack, and one whole iteration of its loop
        # (G-step, E-step, rescale, 3 x rmse; main_responseCalib.cpp:281-362)

        def ms_of(fn, reps=3):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        G_new = torch.zeros_like(G_tab)
        estep["gstep_ms"] = ms_of(lambda: ctx.rc_gstep(data, t_exp, E_out, G_new))
        estep["rmse_ms"] = ms_of(lambda: ctx.rc_rmse(data, t_exp, G_tab, E_out))
        E2, G2 = torch.zeros_like(E_out), torch.zeros_like(G_tab)
        ctx.response_calib(data, t_exp, 1, E2, G2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.response_calib(data, t_exp, 3, E2, G2)
        torch.cuda.synchronize()
        estep["loop_iteration_ms"] = 1e3 * (time.perf_counter() - t0) / 3
        estep["loop"] = "mdc_response_calib (G-step, E-step, rescale, 3 x rmse per iteration)"
        del E2
        del data, E_out

    # ---- e2e through the host-buffer C-ABI entry point (pinned host memory, copies inside the timed region)
    EB = args.e2e_batch
    import ctypes as C
    from mono_dataset_code_b200 import _lib
    h_in, h_out = C.c_void_p(), C.c_void_p()
    _lib.check(_lib.lib.mdc_host_alloc(C.byref(h_in), EB * n_in), "mdc_host_alloc")
    _lib.check(_lib.lib.mdc_host_alloc(C.byref(h_out), EB * n_out * 4), "mdc_host_alloc")
    np_in = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_ubyte)), (EB, n_in))
    np_in[:] = np.random.default_rng(1000 + rank).integers(0, 256, (EB, n_in), dtype=np.uint8)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.prepare_batch_host(np_in, FLAGS_ALL, [h_out.value])
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.prepare_batch_host(np_in, FLAGS_ALL, [h_out.value])      # synchronous: returns after the D2H copy
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * EB * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": EB * n_in,
           "d2h_bytes_per_step": EB * n_out * 4, "frames_per_step": EB, "steps": e2e_steps}
    # where each rank's staging memory lives, and what a plain pinned copy reaches on this rank's PCIe link (the e2e leg moves
    # 5 bytes per output pixel over it: 4 of them device->host)
    link = pcie_ceiling(dev)
    if world > 1:
        numas, links = [None] * world, [None] * world
        dist.all_gather_object(numas, numa)
        dist.all_gather_object(links, link)
    else:
        numas, links = [numa], [link]
    e2e["numa"] = numas
    e2e["pcie_copy_gbs_per_rank"] = links
    e2e["d2h_bound_frames_per_s"] = sum(l["d2h_gbs"] for l in links) * 1e9 / (n_out * 4)
    # latency of ONE frame through the same entry point (what DatasetReader::getImage does per call)
    for _ in range(3):
        ctx.prepare_batch_host(np_in[:1], FLAGS_ALL, [h_out.value])
    t0 = time.perf_counter()
    for i in range(50):
        ctx.prepare_batch_host(np_in[i % EB:i % EB + 1], FLAGS_ALL, [h_out.value])
    e2e["single_frame_ms"] = 1e3 * (time.perf_counter() - t0) / 50
    _lib.lib.mdc_host_free(h_in); _lib.lib.mdc_host_free(h_out)

    c4 = None
    if not args.no_seq:
        try:
            c4 = sequence_leg(args, dev, rank, world, local, numa, barrier)
        except Exception as exc:          # the headline line must not depend on this extra
            c4 = {"error": repr(exc)}
    if rank == 0:
        line = {"metric": "frames_per_s_1280x1024_photometric_fov_undistort", "value": value, "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (uniform random mono8 frames, synthetic TUM-style calibration; SURVEY.md §8d)",
                "config": config_dict(B),
                "run_config": {"frames_per_step_per_gpu": B,
                               "parallelism": f"frame-sharded dp{world}, tables NCCL-broadcast at init, no steady-state collective",
                               "init": init_how,
                               "loader": ({0: "ldg", 1: "tma", 2: "tex", 3: "hybrid"}[args.tma] if args.tma is not None else
                                          "auto(" + ctx.auto_loader() + ")")},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (TRAFFIC["dram_bytes_per_frame"] * B / 1e9 if TRAFFIC else None), "traffic_unit": "GB per launch",
                             "traffic_note": TRAFFIC_NOTE,
                             "peak_source": peak_src, "kernel": "fused_prepare_kernel",
                             "algorithmic_bytes_per_launch": B * ALG_BYTES_PER_FRAME, "launch_ms": k1_ms},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "c3_pyramid": pyr, "c4_sequence": c4, "c5_estep": estep}
        if world == 1 and not args.no_cpu:
            if ORIG_AFFINITY:
                os.sched_setaffinity(0, ORIG_AFFINITY)      # the CPU arm may use every core of the box, not just the GPU's node
            line["cpu_baseline"] = cpu_baseline(files)
            if estep is not None:
                try:
                    estep["cpu_baseline"] = cpu_calibrator_sample()
                except Exception as exc:      # the headline line must not depend on this extra
                    estep["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per step per GPU (device-resident)")
    ap.add_argument("--e2e-batch", type=int, default=256, help="frames per host-buffer call (16 pipeline chunks of 16 frames: fill / drain of the 3-deep pipeline < 2 %%)")
    ap.add_argument("--tma", "--loader", type=int, default=None, dest="tma",
                    help="force K1's input loader: 2 = texture gather, 1 = TMA, 0 = LDG (default: auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--levels", type=int, default=1, help="pyramid levels for --only-kernel sweeps")
    ap.add_argument("--no-estep", action="store_true", help="skip the configs[4] E-step leg")
    ap.add_argument("--no-seq", action="store_true", help="skip the configs[3] 1920x1080 sequence leg")
    ap.add_argument("--seq-frames", type=int, default=10000, help="frames of the configs[3] sequence (whole job)")
    ap.add_argument("--only-kernel", action="store_true", help="tuning sweeps: device-resident K1 timing only (no pyramid / e2e / cpu legs)")
    ap.add_argument("--geom", default=None, help="WxH for --only-kernel sweeps (default 1280x1024)")
    args = ap.parse_args()
    if args.geom:
        set_geometry(*[int(v) for v in args.geom.lower().split("x")])
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
