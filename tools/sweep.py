#!/usr/bin/env python
"""Tuning sweep (not a bench line): K2 time against images-in-flight and lane mapping.
Usage: python tools/sweep.py [--width W --height H] --images 148,296 --lanes 8,4,2,1"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pngloss_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--images", default="148,296,592")
    ap.add_argument("--lanes", default="8,4,2,1")
    ap.add_argument("--bm", default="-1", help="bucket-maxima modes to sweep: -1 library default, 0 off, 1 on")
    ap.add_argument("--lean", default="-1", help="lean-kernel modes to sweep: -1 library default, 0 generic kernel, 1 lean")
    ap.add_argument("--solo", default="0", help="latency-kernel modes to sweep: 0 off, 1 one chain warp, 2 five chain warps, 3 four-warp CTA")
    ap.add_argument("--strength", type=int, default=20)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--profile", action="store_true", help="with a -DPL_K2_PROFILE build: per-filter busy cycles")
    a = ap.parse_args()
    ctx = pngloss_b200.Context(0)
    for n in [int(x) for x in a.images.split(",")]:
        batch = pngloss_b200.Batch(ctx, [a.width] * n, [a.height] * n)
        for i in range(n):
            batch.synth(i, 4 + i)
        ctx.sync()
        for lanes, bm, lean, solo in [(int(x), int(m), int(l), int(so)) for x in a.lanes.split(",")
                                      for m in a.bm.split(",") for l in a.lean.split(",") for so in a.solo.split(",")]:
            ctx.set_solo(solo)
            ctx.set_lanes(lanes)
            ctx.set_bucket_maxima(bm)
            ctx.set_lean(lean)
            best = None
            for _ in range(a.reps + 1):       # first run is the warm-up
                batch.run(a.strength, 2)
                st, _, _ = batch.finish()
                assert (st == 0).all()
                t = batch.timings()
                if best is None or t["k2_quantize_ms"] < best["k2_quantize_ms"]:
                    best = t
            if a.profile:
                h0 = batch.image_histogram(0)
                busy = [int(h0[2 * f]) for f in range(5)]
                print(json.dumps({"images": n, "lanes": lanes, "busy_kcycles_none_sub_up_avg_paeth": busy,
                                  "total_kcycles": int(h0[10]),
                                  "busy_frac": [round(b / max(1, int(h0[10])), 3) for b in busy]}), flush=True)
            px = n * a.width * a.height
            print(json.dumps({"images": n, "w": a.width, "h": a.height, "lanes": lanes, "bm": bm, "lean": lean, "solo_mode": solo,
                              "k1_ms": round(best["k1_hist_ms"], 3), "k2_ms": round(best["k2_quantize_ms"], 3),
                              "k2_mpx_s": round(px / best["k2_quantize_ms"] / 1e3, 1),
                              "k1_gpx_s": round(px / best["k1_hist_ms"] / 1e6, 2),
                              **batch.launch_info()}), flush=True)
        batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
