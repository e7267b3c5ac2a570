#!/usr/bin/env python
"""bench.py - Mpixels/s of the quantise + filter-search hot path at 4K RGBA (BASELINE.json metric).

A "step" is one pass of the hot path (histogram kernel K1, quantise/filter-search kernel K2, batch
histogram kernel K3, plus the one NCCL all-reduce of the 256-bin symbol histogram when N > 1) over
one batch of synthetic 3840x2160 RGBA images (SURVEY 8d generator, seeds 4, 5, ...; BASELINE
configs[2] replicated over seeds so that the machine is filled - a single image is 5 busy warps).

  value     whole-job Mpx/s with the batch resident in HBM, CUDA events, max over ranks
  e2e       the same through the C-ABI host-buffer call pngloss_b200_optimize_batch: pinned host
            buffers, H2D + kernels + D2H inside the timed region
  roofline  K2 (dominant kernel): algorithmic 8 B/px (4 read + 4 write) / K2's CUDA-event duration
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own C code (oracle/_ref, compiled from the unmodified sources) on all
            host cores over a bounded sample of the same workload

--impl reference runs only that CPU arm, as the comparison line the driver asks for.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=2368,
                    help="images per GPU per step (default: 16 per SM, two CTAs of 8 images; falls back to "
                         "half if the device cannot hold them)")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--strength", type=int, default=20)
    ap.add_argument("--bleed", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per channel of K2 (0 = library default)")
    ap.add_argument("--bm", type=int, default=-1,
                    help="K2 candidate choice: 1 bucket maxima, 0 scan only, -1 library default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per core of the CPU sample (0 = auto)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.images} x synthetic {a.width}x{a.height} RGBA gradient+noise per GPU, "
            f"strength {a.strength}, bleed {a.bleed}")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(a):
    """dram bytes per launch from the committed ncu capture (profiles/k2_traffic.json), reported only
    when that capture was taken on the workload being benchmarked."""
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            t = json.load(f)
        w = t["workload"]
        if (w["images_per_gpu"], w["width"], w["height"], w["strength"]) == (a.images, a.width, a.height,
                                                                           a.strength):
            return t
    except Exception:
        pass
    return None


# ---- reference CPU arm ---------------------------------------------------------------------------------
def cpu_reference_run(a, seconds_hint=12.0, rows=0):
    """Times the reference C path on every host core at once: one strip of the 4K workload per core
    (the reference is single-threaded and re-entrant; images are independent)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes

    from checkers import Oracle, Reference, have_reference
    from checkers import row_pointers
    oracle = Oracle()
    cores = os.cpu_count() or 1
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
    # ~0.45 Mpx/s/core on this class of host: size the strip for about seconds_hint of work per core
    if rows <= 0:
        rows = int(max(16, min(a.height, 0.45e6 * seconds_hint / a.width)))
    full = oracle.synth(a.width, a.height, 4)
    strips = []
    for c in range(cores):
        y0 = (c * rows) % max(1, a.height - rows + 1)
        strips.append(np.ascontiguousarray(full[y0:y0 + rows]).copy())
    threads = [threading.Thread(target=one, args=(s,)) for s in strips]   # ctypes drops the GIL
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    px = sum(s.shape[0] * s.shape[1] for s in strips)
    return dict(value=px / dt / 1e6, unit=UNIT, cores=cores, kind=kind, seconds=dt,
                sample=f"{cores} strips of {a.width}x{rows} (rows of the seed-4 4K image), one per core, "
                       f"strength {a.strength} bleed {a.bleed}, optimize_with_rows only")


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": workload_name(a), "note":
                                        "CPU arm: bounded sample of the workload per step"},
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


# ---- the B200 arm ---------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import pngloss_b200
    from pngloss_b200.shard import shard_seeds
    ctx = pngloss_b200.Context(local_rank)
    if a.lanes:
        ctx.set_lanes(a.lanes)
    ctx.set_bucket_maxima(a.bm)
    n, w, h = a.images, a.width, a.height
    px_per_step_rank = n * w * h

    # the device-resident run keeps inputs and outputs apart (every step re-reads the same inputs):
    # 2 x 33 MB per 4K image; fall back to half the images where that does not fit
    while True:
        try:
            batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
            break
        except pngloss_b200.PnglossError as e:
            if e.code != pngloss_b200.OUT_OF_MEMORY or n <= 8:
                raise
            n //= 2
    if world > 1:                               # same batch on every rank
        nmin = torch.tensor([n], device=f"cuda:{local_rank}")
        dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
        if int(nmin.item()) != n:
            batch.close()
            n = int(nmin.item())
            batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    a.images = n
    px_per_step_rank = n * w * h
    seeds = shard_seeds(rank, world, n)        # distinct images on every rank (shards, not replicas)
    for i in range(n):
        batch.synth(i, seeds[i])
    ctx.sync()

    hist_alias = None
    if world > 1:
        # zero-copy torch view of the library's device-side batch histogram for the NCCL all-reduce
        class _Alias:
            __cuda_array_interface__ = {"shape": (256,), "typestr": "<i8", "version": 2,
                                        "data": (batch.histogram_device_ptr(), False)}
        hist_alias = torch.as_tensor(_Alias(), device=f"cuda:{local_rank}")

    def step():
        batch.run(a.strength, a.bleed)
        if world > 1:
            ctx.sync()                          # library stream -> NCCL stream
            dist.all_reduce(hist_alias)         # the one collective of the path: 256 x u64 symbol counts
            torch.cuda.synchronize()

    def fence():
        ctx.sync()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    for _ in range(a.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    k1_ms, k2_ms, k3_ms = [], [], []
    fence()
    ctx.timer_start()
    t_wall = time.perf_counter()
    for _ in range(a.steps):
        tw0 = time.perf_counter()
        step()
        tw1 = time.perf_counter()
        batch.finish()
        tw2 = time.perf_counter()
        t = batch.timings()
        if os.environ.get("PNGLOSS_BENCH_TRACE"):
            print(f"[trace] enqueue {tw1 - tw0:.3f}s finish {tw2 - tw1:.3f}s run_ms {t['run_ms']:.1f}",
                  file=sys.stderr, flush=True)
        k1_ms.append(t["k1_hist_ms"])
        k2_ms.append(t["k2_quantize_ms"])
        k3_ms.append(t["k3_batch_hist_ms"])
    ms = ctx.timer_stop()
    fence()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if sampler else None
    st, bpp, retried = batch.finish()
    assert (st == 0).all(), "quantise kernel reported a failed image"
    info = batch.launch_info()
    global_hist_sum = int(batch.histogram().sum()) if world == 1 else int(hist_alias.sum().item())

    if world > 1:
        tmax = torch.tensor([ms], device=f"cuda:{local_rank}")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / a.steps
    value = world * px_per_step_rank / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call --------------------------------------------------
    e2e = None
    if not a.no_e2e:
        batch.close()                           # give the HBM back; optimize_batch allocates its own
        # Host memory: one pinned input buffer that no step modifies and one pinned output buffer, within
        # 60 % of what the host has free.  When that cannot hold the step's images twice, the output
        # buffer stays complete and the input buffer holds fewer DISTINCT images: image i is uploaded from
        # input slot i % m.  Every step still uploads and downloads every image (the byte counts below are
        # what crosses PCIe); only the content of some uploads repeats.
        img_bytes = w * h * 4
        n2, m = n, n
        try:
            avail = int([ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0].split()[1]) * 1024
            budget = int(0.6 * avail / world / img_bytes)      # images the two buffers may hold together
            n2 = max(1, min(n, 2 * budget // 3))               # at least half of the inputs are distinct
            m = max(1, min(n2, budget - n2))
        except Exception:
            pass
        if world > 1:
            nmin = torch.tensor([n2, m], device=f"cuda:{local_rank}")
            dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
            n2, m = int(nmin[0].item()), int(nmin[1].item())
        # The steps go through the asynchronous call (pngloss_b200_submit / _wait), two in flight, so that
        # the upload of step k+1 and the download of step k-1 run under the kernels of step k.  (A step's
        # results are complete when its wait returns; the next step's download overwrites them afterwards.)
        while True:
            src = None
            try:
                src = ctx.pinned_empty((m, h, w, 4))
                dst = [ctx.pinned_empty((n2, h, w, 4))] * 2
                ok = 1
            except pngloss_b200.PnglossError:
                if src is not None:
                    ctx.free_pinned(src)
                ok = 0
            if world > 1:                           # all ranks shrink together
                flag = torch.tensor([ok], device=f"cuda:{local_rank}")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if ok and not int(flag.item()):
                    ctx.free_pinned(src)
                    ctx.free_pinned(dst[0])
                ok = int(flag.item())
            if ok:
                break
            if n2 <= 8:
                raise RuntimeError("bench: no pinned host memory for the end-to-end part")
            n2, m = n2 // 2, max(1, m // 2)
        b2 = pngloss_b200.Batch(ctx, [w] * m, [h] * m, in_place=True)
        for i in range(m):
            b2.synth(i, seeds[i])
            b2.download_input(i, src[i])
        ctx.sync()
        b2.close()
        filters = [[np.zeros(h, np.uint8) for _ in range(n2)] for _ in range(2)]
        imgs = [src[i % m] for i in range(n2)]
        outs = [[dst[k][i] for i in range(n2)] for k in range(2)]

        def run_steps(count):
            jobs = []
            for it in range(count):
                jobs.append(ctx.submit(imgs, filters[it % 2], a.strength, a.bleed, outputs=outs[it % 2]))
                if len(jobs) == 2:
                    res = jobs.pop(0).wait()
                    assert all(r["status"] == 0 for r in res)
            while jobs:
                res = jobs.pop(0).wait()
                assert all(r["status"] == 0 for r in res)

        run_steps(2)                            # the device is warm; this creates both device batches of the pipeline
        if world > 1:
            dist.barrier()
        ctx.timer_start()                       # CUDA events that cover the upload, compute and download streams
        tw = time.perf_counter()
        run_steps(a.steps)
        e_total = ctx.timer_stop()
        e_wall = (time.perf_counter() - tw) * 1e3
        e_step = e_total / a.steps
        if world > 1:
            tmax = torch.tensor([e_step], device=f"cuda:{local_rank}")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e_step = float(tmax.item())
        e2e = {"value": world * n2 * w * h / (e_step * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(n2 * img_bytes), "d2h_bytes_per_step": int(n2 * (img_bytes + h)),
               "ms_per_step": e_step, "wall_ms_per_step": e_wall / a.steps, "images_per_gpu": n2,
               "distinct_host_inputs_per_gpu": m,
               "api": "pngloss_b200_submit / pngloss_b200_wait, two steps in flight (pinned host input and "
                      "output buffers; every step uploads its input and downloads its result)"}
        ctx.free_pinned(src)
        ctx.free_pinned(dst[0])

    if rank == 0:
        peak, peak_src = measured_peak()
        k2_s = float(np.mean(k2_ms)) * 1e-3
        k1_s = float(np.mean(k1_ms)) * 1e-3
        achieved = K2_BYTES_PER_PX * px_per_step_rank / k2_s / 1e9
        tr = ncu_traffic(a)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(a), "images_per_gpu": n, "width": w, "height": h,
                       "strength": a.strength, "bleed": a.bleed,
                       "l2": f"inputs {n * w * h * 4 / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                       "k2_lanes_per_channel": 8 // info["images_per_cta"],
                       "k2_candidate_choice": "bucket maxima" if info["bucket_maxima"] else "scan",
                       "k2_ctas": info["k2_ctas"], "k2_smem_bytes": info["k2_smem_bytes"],
                       "collective": "nccl all_reduce 256 x u64 per step" if world > 1 else "none (1 GPU)",
                       "wall_ms_per_step": wall_ms / a.steps},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(info["launches"]) * a.steps,
            "roofline": {"bound": "hbm", "kernel": "pl_k2_quantize", "achieved": achieved, "peak": peak,
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
            "kernel_ms": {"k1_orig_hist": float(np.mean(k1_ms)), "k2_quantize": float(np.mean(k2_ms)),
                          "k3_batch_hist": float(np.mean(k3_ms))},
            "checks": {"symbols_counted": global_hist_sum,
                       "symbols_expected": world * px_per_step_rank * 4,
                       "retried_rows": int(retried.sum())},
        }
        if world == 1 and not a.no_cpu:
            cb = cpu_reference_run(a, rows=a.cpu_rows)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
