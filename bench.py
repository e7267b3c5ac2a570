#!/usr/bin/env python
"""bench.py - Mpixels/s of the quantise + filter-search hot path (BASELINE.json metric).

A "step" is one pass of the hot path (histogram kernel K1, quantise / filter-search kernel K2, batch
histogram kernel K3, plus the library's NCCL all-reduce of the 256-bin symbol histogram when N > 1) over
one batch of images.  --config picks the workload; every BASELINE.json config has one:

  (default) 3s  the metric's configuration, saturated: IMAGES x synthetic 3840x2160 RGBA per GPU
            (SURVEY 8d generator, seeds 4, 5, ...; BASELINE configs[2] replicated over seeds so that the
            machine is filled - one image is five dependent chains).  Weak scaling.
  1         suite/david.png -s19 -b2, one image (BASELINE configs[0]; pixels from tests/golden/fixtures.npz)
  2         suite/lena.png 512x512 -s20, one image (configs[1])
  3         one synthetic 3840x2160 RGBA image, --strength 0/20/40/85 (configs[2], as the reference's command
            line would run it: one image per call)
  4         1024 synthetic 1920x1080 RGBA images in total, sharded over the GPUs (configs[3]); strong scaling
  5         64 synthetic 8192x8192 RGBA images in total, sharded over the GPUs (configs[4]); strong scaling

  value     whole-job Mpx/s with the batch resident in HBM, CUDA events, max over ranks
  e2e       the same through the C-ABI host-buffer calls (pngloss_b200_submit / _wait, or the reference's own
            entry point optimize_with_rows for the one-image configs): pinned host buffers, H2D + kernels +
            D2H inside the timed region
  roofline  K2 (dominant kernel): algorithmic 8 B/px (4 read + 4 write) / K2's CUDA-event duration against
            the measured HBM copy bandwidth (MEASURED_PEAKS.json); roofline_k1 likewise (4 B/px)
  latency_bound  what actually bounds K2: dependent pixel steps per image, SM cycles per step, chains in flight
  cpu_baseline   the reference's own C code (oracle/_ref, compiled from the unmodified sources) on the host
            cores over a bounded sample of the same workload

Multi-GPU runs (torchrun launches one process per GPU) do not import torch: the ranks meet through the
library's own NCCL communicator (pngloss_b200_comm_*), the id travels through a file in /tmp.
--impl reference runs only the CPU arm, as the comparison line the driver asks for.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/s quantize+filter-search @4K RGBA"
UNIT = "Mpx/s"
K2_BYTES_PER_PX = 8.0    # DESIGN.md / SURVEY 8d: 4 B read + 4 B written per pixel
K1_BYTES_PER_PX = 4.0
L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="3s", choices=["1", "2", "3", "3s", "4", "5"],
                    help="BASELINE.json config (see the module docstring); default 3s = the metric's config, saturated")
    ap.add_argument("--images", type=int, default=0,
                    help="3s: images per GPU per step (default 2368 = 16 per SM: two whole steps fit the device, so the "
                         "end-to-end pipeline overlaps copies and kernels step against step; 3552 = 24 per SM "
                         "runs the large-batch kernel, see DESIGN.md); 4 / 5: total images of the job "
                         "(default 1024 / 64)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--strength", type=int, default=-1)
    ap.add_argument("--bleed", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per channel of K2 (0 = library default)")
    ap.add_argument("--bm", type=int, default=-1,
                    help="K2 candidate choice: 1 bucket maxima, 0 scan only, -1 library default")
    ap.add_argument("--lean", type=int, default=-1, help="K2 large-batch kernel: 1/-1 on where it applies, 0 off")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-k4", action="store_true", help="skip the K4 (scanlines) roofline measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per core of the CPU sample (0 = auto)")
    a = ap.parse_args()
    spec = {
        "3s": dict(w=3840, h=2160, strength=20, per_gpu=2368, total=0, scaling="weak", source="synth", seed0=4),
        "1": dict(w=180, h=215, strength=19, per_gpu=1, total=0, scaling="weak", source="david", seed0=0),
        "2": dict(w=512, h=512, strength=20, per_gpu=1, total=0, scaling="weak", source="lena", seed0=0),
        "3": dict(w=3840, h=2160, strength=20, per_gpu=1, total=0, scaling="weak", source="synth", seed0=4),
        "4": dict(w=1920, h=1080, strength=20, per_gpu=0, total=1024, scaling="strong", source="synth", seed0=100),
        "5": dict(w=8192, h=8192, strength=20, per_gpu=0, total=64, scaling="strong", source="synth", seed0=1000),
    }[a.config]
    a.width = a.width or spec["w"]
    a.height = a.height or spec["h"]
    a.strength = spec["strength"] if a.strength < 0 else a.strength
    a.scaling = spec["scaling"]
    a.source = spec["source"]
    a.seed0 = spec["seed0"]
    if a.scaling == "strong":
        a.total = a.images or spec["total"]
        a.images = 0
    else:
        a.total = 0
        a.images = a.images or spec["per_gpu"]
    return a


def workload_name(a, n_per_gpu):
    if a.source in ("david", "lena"):
        return (f"suite/{a.source}.png {a.width}x{a.height} (decoded RGBA from tests/golden/fixtures.npz), "
                f"one image, strength {a.strength}, bleed {a.bleed}")
    if a.scaling == "strong":
        return (f"{a.total} x synthetic {a.width}x{a.height} RGBA gradient+noise in total, sharded over the GPUs, "
                f"strength {a.strength}, bleed {a.bleed}")
    return (f"{n_per_gpu} x synthetic {a.width}x{a.height} RGBA gradient+noise per GPU, "
            f"strength {a.strength}, bleed {a.bleed}")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(a, n):
    """dram bytes per launch from the committed ncu capture (profiles/k2_traffic.json), reported only
    when that capture was taken on the workload being benchmarked."""
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            t = json.load(f)
        w = t["workload"]
        if (w["images_per_gpu"], w["width"], w["height"], w["strength"]) == (n, a.width, a.height, a.strength):
            return t
    except Exception:
        pass
    return None


def fixture_image(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    return np.ascontiguousarray(z[name])


# ---- reference CPU arm ---------------------------------------------------------------------------------
def cpu_reference_run(a, seconds_hint=12.0, rows=0, cores=None):
    """Times the reference C path on host cores: one strip (or one copy of the image, for the small suite
    images) per core, all at once (the reference is single-threaded and re-entrant; images are independent)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import Oracle, Reference, have_reference
    from checkers import row_pointers
    oracle = Oracle()
    cores = cores or os.cpu_count() or 1
    if have_reference():
        kind, ref = "reference", Reference()

        def one(img):
            h, w, _ = img.shape
            rf = np.zeros(h, np.uint8)
            ref.lib.optimize_with_rows(row_pointers(img), w, h, rf.ctypes.data, False, a.strength, a.bleed)
    else:
        kind = "port"

        def one(img):
            h, w, _ = img.shape
            rf = np.zeros(h, np.uint8)
            oracle.lib.oracle_optimize_with_rows(row_pointers(img), w, h, rf.ctypes.data, a.strength,
                                                 a.bleed, None)
    if a.source in ("david", "lena"):
        full = fixture_image(a.source)
        strips = [full.copy() for _ in range(cores)]
        what = f"{cores} copies of the whole image, one per core"
    else:
        # ~0.45 Mpx/s/core on this class of host: size the strip for about seconds_hint of work per core
        if rows <= 0:
            rows = int(max(16, min(a.height, 0.45e6 * seconds_hint / a.width)))
        gen_h = min(a.height, 2160)     # strips come from the top rows of the seed-`seed0` image
        full = oracle.synth(a.width, gen_h, a.seed0)
        strips = []
        for c in range(cores):
            y0 = (c * rows) % max(1, gen_h - rows + 1)
            strips.append(np.ascontiguousarray(full[y0:y0 + rows]).copy())
        what = (f"{cores} strips of {a.width}x{rows} (rows of the seed-{a.seed0} {a.width}x{a.height} image), "
                f"one per core")
    threads = [threading.Thread(target=one, args=(s,)) for s in strips]   # ctypes drops the GIL
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    px = sum(s.shape[0] * s.shape[1] for s in strips)
    return dict(value=px / dt / 1e6, unit=UNIT, cores=cores, kind=kind, seconds=dt,
                sample=f"{what}, strength {a.strength} bleed {a.bleed}, optimize_with_rows only")


def config_keys(a, n_per_gpu):
    return {"workload": workload_name(a, n_per_gpu), "baseline_config": a.config, "images_per_gpu": n_per_gpu,
            "width": a.width, "height": a.height, "strength": a.strength, "bleed": a.bleed}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    for _ in range(a.warmup):
        cpu_reference_run(a, rows=16)
    vals, secs, last = [], [], None
    for _ in range(a.steps):
        last = cpu_reference_run(a, seconds_hint=8.0, rows=a.cpu_rows)
        vals.append(last["value"])
        secs.append(last["seconds"])
    value = float(np.mean(vals))
    n_per_gpu = a.images if a.scaling == "weak" else a.total // max(1, world)
    cfg = config_keys(a, n_per_gpu)
    cfg["note"] = "CPU arm: every step times a bounded sample of this workload on all host cores"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic" if a.source == "synth" else "suite image", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks -------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- rendezvous of the ranks without torch ----------------------------------------------------------------------
def exchange_comm_id(rank, world):
    """Rank 0 creates the NCCL id and publishes it through a file in /tmp (one node, torchrun contract); the
    other ranks pick it up.  Files older than this process are leftovers of earlier runs and are ignored."""
    import pngloss_b200
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'none')}"
    path = f"/tmp/pngloss_b200_comm_{tag}.id"
    started = time.time() - 120.0        # every rank of one launch starts within two minutes of rank 0
    if rank == 0:
        uid = pngloss_b200.comm_unique_id()
        tmp = path + f".{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid, path
    deadline = time.time() + 300
    while time.time() < deadline:
        try:
            st = os.stat(path)
            if st.st_mtime >= started and st.st_size == pngloss_b200.COMM_ID_BYTES:
                with open(path, "rb") as f:
                    return f.read(), path
        except FileNotFoundError:
            pass
        time.sleep(0.05)
    raise RuntimeError("bench: no NCCL id from rank 0")


# ---- the B200 arm ---------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import pngloss_b200
    from pngloss_b200.shard import shard_seeds
    ctx = pngloss_b200.Context(local_rank)
    id_path = None
    if world > 1:
        # stale id files of an earlier launch on the same port are older than two minutes or get replaced here
        uid, id_path = exchange_comm_id(rank, world)
        ctx.comm_init_rank(world, rank, uid)
        ctx.barrier()
        if rank == 0:
            try:
                os.remove(id_path)
            except OSError:
                pass

    def allmax(v):
        return float(ctx.comm_allreduce([int(v * 1e3)], "max")[0]) / 1e3 if world > 1 else float(v)

    def allmin_int(v):
        return int(ctx.comm_allreduce([int(v)], "min")[0]) if world > 1 else int(v)

    if a.lanes:
        ctx.set_lanes(a.lanes)
    ctx.set_bucket_maxima(a.bm)
    ctx.set_lean(a.lean)
    w, h = a.width, a.height
    if a.scaling == "strong":
        # the job's images are dealt round-robin; every rank gets the same count (the totals divide by 8)
        n = a.total // world
        seeds = [a.seed0 + rank + world * i for i in range(n)]
        job_images = n * world
    else:
        n = a.images
        seeds = None
        job_images = None

    # Device-resident batch.  Large batches run in place (half the memory: 24 4K images per SM need 118 GB):
    # the step then re-creates its inputs on the device first (the synthetic generator kernel, ~0.5 % of the
    # step, inside the timed region and counted in gpu_launches).  Small batches keep inputs and outputs apart.
    img_bytes = w * h * 4
    in_place = a.source == "synth" and n * img_bytes * 2 > 165e9
    while True:
        try:
            batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n, in_place=in_place)
            break
        except pngloss_b200.PnglossError as e:
            if e.code != pngloss_b200.OUT_OF_MEMORY or n <= 8 or a.scaling == "strong":
                raise
            n = n * 2 // 3
    nmin = allmin_int(n)
    if nmin != n:                               # same batch on every rank
        batch.close()
        n = nmin
        batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n, in_place=in_place)
    if a.scaling == "weak":
        a.images = n
        seeds = shard_seeds(rank, world, n) if a.source == "synth" else None
        if seeds is not None and a.seed0 != 4:
            seeds = [s - 4 + a.seed0 for s in seeds]
    px_per_step_rank = n * w * h

    host_img = fixture_image(a.source) if a.source in ("david", "lena") else None

    def load_inputs():
        if host_img is not None:
            for i in range(n):
                batch.upload(i, host_img)
        else:
            for i in range(n):
                batch.synth(i, seeds[i])

    load_inputs()
    ctx.sync()
    small = n * img_bytes < 4 * L2_BYTES        # inputs that could stay in L2 between steps: flush it
    launches_per_step = [0]

    def step():
        extra = 0
        if in_place:
            load_inputs()                       # restore the inputs the previous step overwrote
            extra += n
        if small:
            ctx.flush_l2()                      # a memset, not one of our kernels: not counted
        batch.run(a.strength, a.bleed)
        if world > 1:
            batch.allreduce_histogram()         # the one collective of the path: 256 x u64 symbol counts
        launches_per_step[0] = extra + 3

    def fence():
        ctx.sync()
        ctx.barrier()

    for _ in range(a.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    k1_ms, k2_ms, k3_ms, run_ms = [], [], [], []
    fence()
    ctx.timer_start()
    t_wall = time.perf_counter()
    for _ in range(a.steps):
        step()
        batch.finish()
        t = batch.timings()
        k1_ms.append(t["k1_hist_ms"])
        k2_ms.append(t["k2_quantize_ms"])
        k3_ms.append(t["k3_batch_hist_ms"])
        run_ms.append(t["run_ms"])
    ms = ctx.timer_stop()
    fence()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if sampler else None
    st, bpp, retried = batch.finish()
    assert (st == 0).all(), "quantise kernel reported a failed image"
    info = batch.launch_info()
    global_hist_sum = int(batch.histogram().sum())

    if small:
        # the bracket contains the L2 flushes; the step is the sum of its kernels' own events
        ms_per_step = allmax(float(np.mean(run_ms)))
    else:
        ms_per_step = allmax(ms) / a.steps
    value = world * px_per_step_rank / (ms_per_step * 1e-3) / 1e6

    # ---- K4 (filtered scanlines: the repo's one HBM-bound hot kernel), outside the timed region, on a batch of its
    # own that is large enough to saturate HBM and small enough to sit next to nothing else ------------------------
    k4 = None
    if world == 1 and not a.no_k4:
        batch.close()
        batch = None
        k4 = measure_k4(ctx, pngloss_b200, min(n, 296), w, h, a.strength, a.bleed)

    # ---- end to end through the host-buffer C-ABI calls -------------------------------------------------
    e2e = None
    if not a.no_e2e:
        if batch is not None:
            batch.close()                       # give the HBM back; the host-buffer calls allocate their own
        batch = None
        if n == 1:
            e2e = e2e_single_image(a, ctx, pngloss_b200, host_img, seeds, w, h)
        else:
            e2e = e2e_job_api(a, ctx, pngloss_b200, seeds, n, w, h, world, allmax, allmin_int)

    if rank == 0:
        peak, peak_src = measured_peak()
        k2_s = float(np.mean(k2_ms)) * 1e-3
        k1_s = float(np.mean(k1_ms)) * 1e-3
        achieved = K2_BYTES_PER_PX * px_per_step_rank / k2_s / 1e9
        tr = ncu_traffic(a, n)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        images_per_cta = max(1, info["images_per_cta"])
        cfg = config_keys(a, n)
        cfg.update({
            "l2": ("flushed between steps (512 MB memset); the step time is the sum of its kernels' CUDA events"
                   if small else f"inputs {n * img_bytes / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed"),
            "device_batch": ("in place; the step re-creates its inputs with the generator kernel first"
                             if in_place else "separate input and output buffers"),
            "k2_kernel": "pl_k2_solo" if info.get("solo") else "pl_k2_lean" if info.get("lean") else "pl_k2_quantize",
            "k2_lanes_per_channel": 1 if info.get("solo") else 8 // images_per_cta,
            "k2_candidate_choice": "bucket maxima" if info["bucket_maxima"] else "scan",
            "k2_ctas": info["k2_ctas"], "k2_smem_bytes": info["k2_smem_bytes"],
            "collective": ("nccl all_reduce 256 x u64 per step, issued by the library "
                           "(pngloss_b200_batch_allreduce_histogram)") if world > 1 else "none (1 GPU)",
            "wall_ms_per_step": wall_ms / a.steps})
        if job_images:
            cfg["images_total"] = job_images
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic" if a.source == "synth" else "suite image",
            "config": cfg,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches_per_step[0]) * a.steps,
            "roofline": {"bound": "hbm", "kernel": cfg["k2_kernel"], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tr.get("k2_dram_bytes_per_launch") if tr else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_px": K2_BYTES_PER_PX,
                         "kernel_ms": k2_s * 1e3,
                         "note": "K2 is bound by its serial per-byte dependency chain, not by HBM "
                                 "(DESIGN.md); frac is reported against HBM as the spec asks"},
            "roofline_k1": {"bound": "hbm", "kernel": "pl_k1_orig_hist",
                            "achieved": K1_BYTES_PER_PX * px_per_step_rank / k1_s / 1e9, "peak": peak,
                            "unit": "GB/s", "frac": K1_BYTES_PER_PX * px_per_step_rank / k1_s / 1e9 / peak,
                            "traffic": tr.get("k1_dram_bytes_per_launch") if tr else None,
                            "kernel_ms": k1_s * 1e3},
            # What bounds K2: every image is w*h dependent pixel steps per filter candidate;
            # a step costs `cycles_per_pixel_step` SM cycles of latency, and images * 5 such chains run at once.
            "latency_bound": latency_bound(info, n, w, h, k2_s, sm_mhz),
            "roofline_k4": ({"bound": "hbm", "kernel": "pl_k4_scanlines", "achieved": k4["filter_gb_s"], "peak": peak,
                             "unit": "GB/s", "frac": k4["filter_gb_s"] / peak, "traffic": None,
                             "kernel_ms": k4["filter_ms"], "images": k4["images"],
                             "algorithmic_bytes_per_px": 4 + k4["bytes_per_pixel"],
                             "scan_kernel": {"kernel": "pl_k4_scan_output", "achieved": k4["scan_gb_s"],
                                             "frac": k4["scan_gb_s"] / peak, "kernel_ms": k4["scan_ms"]},
                             "note": "not part of the timed step: what the encoder's filtering pass costs on the "
                                     "device (pngloss_b200_batch_scanlines), timed alone with CUDA events"}
                            if k4 else None),
            "kernel_ms": {"k1_orig_hist": float(np.mean(k1_ms)), "k2_quantize": float(np.mean(k2_ms)),
                          "k3_batch_hist": float(np.mean(k3_ms))},
            "checks": {"symbols_counted": global_hist_sum,
                       "symbols_expected": world * px_per_step_rank * int(bpp[0]),
                       "retried_rows": int(retried.sum())},
        }
        if world == 1 and not a.no_cpu:
            cb = cpu_reference_run(a, rows=a.cpu_rows)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cb1 = cpu_reference_run(a, seconds_hint=4.0, rows=a.cpu_rows, cores=1)
            line["cpu_baseline"]["one_core"] = {"value": cb1["value"], "unit": UNIT, "sample": cb1["sample"]}
        print(json.dumps(line), flush=True)
    if batch is not None:
        batch.close()
    ctx.barrier()
    ctx.close()


def measure_k4(ctx, pngloss_b200, m, w, h, strength, bleed):
    """Device time of the two scanline kernels on m quantised images (best of three)."""
    batch = pngloss_b200.Batch(ctx, [w] * m, [h] * m, in_place=True)
    for i in range(m):
        batch.synth(i, 4 + i)
    batch.run(strength, bleed)
    batch.finish()
    best = None
    for _ in range(3):
        batch.scanlines()
        info = batch.scanline_info(0)
        t = (info["k4_scan_ms"], info["k4_filter_ms"])
        best = t if best is None else (min(best[0], t[0]), min(best[1], t[1]))
    bpp = info["bytes_per_pixel"]
    batch.close()
    px = m * w * h
    return {"images": m, "bytes_per_pixel": bpp, "scan_ms": best[0], "filter_ms": best[1],
            "scan_gb_s": px * 4 / best[0] / 1e6, "filter_gb_s": (px * (4 + bpp) + m * h) / best[1] / 1e6}


def latency_bound(info, n, w, h, k2_s, sm_mhz):
    """What bounds K2.  Every image is w*h dependent pixel steps per filter candidate; all
    images resident on the SMs advance together, so K2's time = waves x (w*h) x the wall-clock cost of one step,
    and throughput = images in flight / that cost."""
    ctas_per_sm = 2 if info.get("solo") else 3 if info.get("lean") else {8: 2, 4: 3}.get(info["images_per_cta"], 4)
    resident = 148 * ctas_per_sm
    waves = max(1, -(-info["k2_ctas"] // resident))
    in_flight = min(n, resident * info["images_per_cta"])
    cycles = k2_s * sm_mhz * 1e6 / (w * h * waves)
    return {"pixel_steps_per_image": w * h, "images_in_flight": in_flight, "chains_in_flight": in_flight * 5,
            "waves": waves, "cycles_per_pixel_step": cycles, "sm_mhz": sm_mhz,
            "mpx_s_per_image": sm_mhz / cycles,
            "bound_mpx_s": in_flight * sm_mhz / cycles,
            "note": "cycles_per_pixel_step = K2 time x SM clock / (pixels of one image x waves): the latency of one "
                    "step of an image's five dependent chains at this occupancy; bound_mpx_s = images in flight x "
                    "SM clock / cycles_per_pixel_step.  Warp instructions and issue-slot utilisation per step: "
                    "profiles/ (ncu)"}


def e2e_single_image(a, ctx, pngloss_b200, host_img, seeds, w, h):
    """One image per call through the reference's own entry point, optimize_with_rows (src/pngloss.c:266):
    host buffer in, quantised in place, row filters out; upload, kernels and download inside the timed call."""
    if host_img is None:
        b2 = pngloss_b200.Batch(ctx, [w], [h], in_place=True)
        b2.synth(0, seeds[0])
        src = ctx.pinned_empty((h, w, 4))
        b2.download_input(0, src)
        ctx.sync()
        b2.close()
    else:
        src = ctx.pinned_empty((h, w, 4))
        src[:] = host_img
    work = ctx.pinned_empty((h, w, 4))
    rf = np.zeros(h, np.uint8)
    times = []
    for it in range(a.warmup + a.steps):
        work[:] = src                           # the call quantises in place: restore the input (not timed)
        ctx.flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        rc = pngloss_b200.optimize_with_rows(work, rf, False, a.strength, a.bleed)
        dt = time.perf_counter() - t0
        assert rc == 0
        if it >= a.warmup:
            times.append(dt)
    e_step = float(np.mean(times)) * 1e3
    out = {"value": w * h / (e_step * 1e-3) / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": int(w * h * 4), "d2h_bytes_per_step": int(w * h * 4 + h),
           "ms_per_step": e_step, "images_per_gpu": 1,
           "api": "optimize_with_rows (the reference's own entry point, one image per call; host wall clock "
                  "around the blocking call, pinned host buffer)"}
    ctx.free_pinned(src)
    ctx.free_pinned(work)
    return out


def e2e_job_api(a, ctx, pngloss_b200, seeds, n, w, h, world, allmax, allmin_int):
    """The step through pngloss_b200_submit / _wait with host buffers.

    A job is a whole step and two steps are in flight: the upload of step k+1 and the download of step k-1 run
    under the kernels of step k.  When two steps do not fit the device (--images 3552: 118 GB each), the step is
    cut into jobs of 296 images (37 CTAs of the large-batch kernel) and the context runs as a pipeline
    (pngloss_b200_ctx_set_pipeline): every job computes on its own stream, about twelve jobs share the SMs at any
    time, and as one job's CTAs finish the next job's take their places.

    Host buffers: the output buffer holds a whole job (the pipeline: a ring of four jobs); the input buffer is a
    ring of at most 20 GB that shrinks to what the host can spare per rank, so that images_per_gpu stays the same
    at every N: image i is uploaded from input slot i % slots_in (inputs are never modified).  Every step uploads
    and downloads every image; the byte counts below are what crosses PCIe."""
    img_bytes = w * h * 4
    pipeline = a.config == "3s" and n * img_bytes * 2 > 165e9     # two whole steps do not fit the device
    if pipeline:
        per_job = 296
        jobs_per_step = -(-n // per_job)
        in_flight = min(jobs_per_step, 12) + 3
        ctx.set_lanes(1)
        ctx.set_lean(1)                         # every job is a small grid; together they fill three CTAs per SM
        ctx.set_pipeline(in_flight)
        out_segments = 4
    else:
        per_job = n                             # a job is a whole step; two steps in flight
        jobs_per_step = 1
        in_flight = 2
        out_segments = 1                        # step k+1's download starts long after step k's wait returned
    # host rings: the output ring holds whole jobs; the input ring shrinks to what the host can spare (inputs are
    # never modified, image i is uploaded from slot i % slots_in), at most 20 GB
    try:
        avail = int([ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 64 << 30
    spare = int(0.6 * avail / world) - out_segments * per_job * img_bytes
    slots_in = allmin_int(max(8, min(n, int(20e9 // img_bytes), spare // img_bytes)))
    src = ctx.pinned_empty((slots_in, h, w, 4))
    dst = ctx.pinned_empty((out_segments * per_job, h, w, 4))
    gen = min(slots_in, 64)
    b2 = pngloss_b200.Batch(ctx, [w] * gen, [h] * gen, in_place=True)
    for base in range(0, slots_in, gen):
        cnt = min(gen, slots_in - base)
        for i in range(cnt):
            b2.synth(i, seeds[(base + i) % len(seeds)])
            b2.download_input(i, src[base + i])
        ctx.sync()
    b2.close()
    filters = [np.zeros(h, np.uint8) for _ in range(out_segments * per_job)]
    state = {"job": 0}

    def job_args(j):
        idx = list(range(j * per_job, min(n, (j + 1) * per_job)))
        seg = (state["job"] % out_segments) * per_job
        state["job"] += 1
        return ([src[i % slots_in] for i in idx], [filters[seg + k] for k in range(len(idx))],
                [dst[seg + k] for k in range(len(idx))])

    trace = os.environ.get("PNGLOSS_BENCH_TRACE")
    rank = int(os.environ.get("RANK", "0"))

    def run_steps(count):
        # PNGLOSS_BENCH_TRACE=1: host-side time line of every job (seconds since the start of the call): when submit
        # was entered and returned, when wait was entered and returned.  A wait that returns long after the
        # kernels' share of the step is what PCIe / host-memory contention looks like.
        inflight, t0 = [], time.perf_counter()

        def finish(item):
            job, k, ts0, ts1 = item
            tw0 = time.perf_counter()
            res = job.wait()
            tw1 = time.perf_counter()
            assert all(r["status"] == 0 for r in res)
            if trace:
                print(f"[e2e-trace] rank {rank} job {k} submit {ts0 - t0:.3f}-{ts1 - t0:.3f} "
                      f"wait {tw0 - t0:.3f}-{tw1 - t0:.3f}", file=sys.stderr, flush=True)

        k = 0
        for _ in range(count):
            for j in range(jobs_per_step):
                ins, rfs, outs = job_args(j)
                ts0 = time.perf_counter()
                job = ctx.submit(ins, rfs, a.strength, a.bleed, outputs=outs)
                inflight.append((job, k, ts0, time.perf_counter()))
                k += 1
                if len(inflight) >= in_flight:
                    finish(inflight.pop(0))
        while inflight:
            finish(inflight.pop(0))

    run_steps(1)                                # creates the device batches of the pipeline
    ctx.barrier()
    ctx.timer_start()                           # CUDA events that cover the upload, compute and download streams
    tw = time.perf_counter()
    run_steps(a.steps)
    e_total = ctx.timer_stop()
    e_wall = (time.perf_counter() - tw) * 1e3
    e_step = allmax(e_total / a.steps)
    out = {"value": world * n * w * h / (e_step * 1e-3) / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": int(n * img_bytes), "d2h_bytes_per_step": int(n * (img_bytes + h)),
           "ms_per_step": e_step, "wall_ms_per_step": e_wall / a.steps, "images_per_gpu": n,
           "jobs_per_step": jobs_per_step, "images_per_job": per_job, "jobs_in_flight": in_flight,
           "host_input_ring_images": slots_in, "host_output_ring_images": out_segments * per_job,
           "api": ("pngloss_b200_submit / pngloss_b200_wait, pipeline of %d jobs in flight on their own streams "
                   "(pngloss_b200_ctx_set_pipeline)" % in_flight) if pipeline else
                  "pngloss_b200_submit / pngloss_b200_wait, two steps in flight (the copies of one step run under "
                  "the kernels of the other)",
           "note": "pinned host rings; every step uploads all its inputs and downloads all its results; the timed "
                   "region starts and ends with an idle device (pipeline fill and drain included)"}
    ctx.free_pinned(src)
    ctx.free_pinned(dst)
    return out


if __name__ == "__main__":
    main()
