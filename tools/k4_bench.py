#!/usr/bin/env python
"""Device time and achieved HBM bandwidth of the scanline kernels (K4) on a batch of synthetic 4K images.
Usage: python tools/k4_bench.py [--images N --width W --height H]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pngloss_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1184)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--k2-rows", type=int, default=0,
                    help="unused; K4 is timed on K2's output of the full image")
    a = ap.parse_args()
    n, w, h = a.images, a.width, a.height
    ctx = pngloss_b200.Context(0)
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n, in_place=True)
    for i in range(n):
        batch.synth(i, 4 + i)
    batch.run(20, 2)
    batch.finish()
    best = None
    for _ in range(4):
        batch.scanlines()
        info = batch.scanline_info(0)
        t = (info["k4_scan_ms"], info["k4_filter_ms"])
        best = t if best is None else (min(best[0], t[0]), min(best[1], t[1]))
    px = n * w * h
    bpp = info["bytes_per_pixel"]
    # scan kernel reads 4 B/px; scanline kernel reads 4 B/px and writes bpp B/px (+ 1 B per row)
    scan_bytes, filt_bytes = px * 4, px * (4 + bpp) + n * h
    print(json.dumps({"images": n, "w": w, "h": h, "bytes_per_pixel": bpp,
                      "k4_scan_ms": round(best[0], 3), "k4_scan_gb_s": round(scan_bytes / best[0] / 1e6, 1),
                      "k4_filter_ms": round(best[1], 3), "k4_filter_gb_s": round(filt_bytes / best[1] / 1e6, 1),
                      "k4_gpx_s": round(px / (best[0] + best[1]) / 1e6, 2),
                      "k2_ms": round(batch.timings()["k2_quantize_ms"], 1)}))
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
