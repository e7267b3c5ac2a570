#!/usr/bin/env python
"""Tuning aid (not a bench line): K2 time of single real images (tests/golden/suite_fixtures.npz) and of a synthetic
strip through the library's default kernel choice.  Usage: PNGLOSS_B200_LIB=... python tools/solo_ab.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pngloss_b200  # noqa: E402


def rgba(a):
    h, w, c = a.shape
    out = np.zeros((h, w, 4), np.uint8)
    if c == 1:
        out[..., :3] = a
        out[..., 3] = 255
    elif c == 3:
        out[..., :3] = a
        out[..., 3] = 255
    else:
        out[:] = a
    return out


def main():
    ctx = pngloss_b200.Context(0)
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "suite_fixtures.npz"))
    res = {}
    for name in ["tux", "redbrush", "dice", "girl", "barbara", "tenko"]:
        img = rgba(z[name])
        h, w, _ = img.shape
        batch = pngloss_b200.Batch(ctx, [w], [h])
        batch.upload(0, img)
        best = 1e9
        for _ in range(3):
            batch.run(20, 2)
            st, _, _ = batch.finish()
            assert (st == 0).all()
            best = min(best, batch.timings()["k2_quantize_ms"])
        res[name] = round(w * h / best / 1e3, 3)
        batch.close()
    batch = pngloss_b200.Batch(ctx, [3840], [135])
    batch.synth(0, 4)
    best = 1e9
    for _ in range(3):
        batch.run(20, 2)
        batch.finish()
        best = min(best, batch.timings()["k2_quantize_ms"])
    res["synth3840x135"] = round(3840 * 135 / best / 1e3, 3)
    batch.close()
    print(json.dumps({"k2_mpx_s": res}))
    ctx.close()


if __name__ == "__main__":
    main()
