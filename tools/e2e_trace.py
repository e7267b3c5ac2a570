#!/usr/bin/env python
"""Timeline of the asynchronous host-buffer call: when does each submit / wait return?
Usage: python tools/e2e_trace.py [--images N --height H --steps K]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pngloss_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1184)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=270)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warm", type=int, default=2, help="jobs in flight during the warm-up")
    a = ap.parse_args()
    n, w, h = a.images, a.width, a.height
    ctx = pngloss_b200.Context(0)
    src = ctx.pinned_empty((n, h, w, 4))
    dst = [ctx.pinned_empty((n, h, w, 4)) for _ in range(2)]
    b = pngloss_b200.Batch(ctx, [w] * n, [h] * n, in_place=True)
    for i in range(n):
        b.synth(i, 4 + i)
        b.download_input(i, src[i])
    ctx.sync()
    b.close()
    imgs = [src[i] for i in range(n)]
    outs = [[dst[k][i] for i in range(n)] for k in range(2)]
    filters = [[np.zeros(h, np.uint8) for _ in range(n)] for _ in range(2)]

    def run(count, trace):
        t0 = time.perf_counter()
        jobs = []

        def stamp(what):
            if trace:
                print(f"  {time.perf_counter() - t0:8.3f} s  {what}", flush=True)
        for it in range(count):
            stamp(f"submit {it} ...")
            jobs.append(ctx.submit(imgs, filters[it % 2], 20, 2, outputs=outs[it % 2]))
            stamp(f"submit {it} returned")
            if len(jobs) == 2:
                jobs.pop(0).wait()
                stamp(f"wait {it - 1} returned")
        while jobs:
            jobs.pop(0).wait()
            stamp("wait (drain) returned")
        return time.perf_counter() - t0

    print("warm-up:", a.warm, "jobs")
    run(a.warm, True)
    ctx.timer_start()
    dt = run(a.steps, True)
    ms = ctx.timer_stop()
    print(f"{a.steps} steps: wall {dt:.3f} s, events {ms / 1e3:.3f} s, {n * w * h * a.steps / dt / 1e6:.1f} Mpx/s end to end")
    ctx.timer_start()
    ctx.optimize_batch(imgs, filters[0], 20, 2, outputs=outs[0])
    print(f"one blocking call: {ctx.timer_stop() / 1e3:.3f} s")
    ctx.close()


if __name__ == "__main__":
    main()
