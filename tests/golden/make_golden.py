#!/usr/bin/env python
"""Generate tests/golden/golden.json and tests/golden/fixtures.npz.

TEST INFRASTRUCTURE.  Run in the dev container, where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

Every expected value is produced by the UNMODIFIED reference hot path
(oracle/_ref/libpngloss_ref.so, compiled from /root/reference/src by oracle/Makefile) through its
public entry point optimize_with_rows (reference src/pngloss_image.c:52).  The reference ships no
golden outputs of its own (its suite/run_suite.sh has no assertions), so these vectors are the pin.

Inputs are either
  * "synth": the stateless splitmix64 gradient+noise generator (oracle_synth_rgba), so
    only (w, h, seed) is stored, or
  * "fixture": decoded reference suite images (or crops of them) stored in fixtures.npz, because the
    reference tree does not exist on the GPU box.
For every case we store sha256[:16] of the input, of the output RGBA buffer and of row_filters[],
plus the per-filter row counts (none, sub, up, avg, paeth).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from checkers import (Oracle, Reference, filter_counts, load_suite_rgba, sha16, to_bpp)  # noqa: E402

LARGE = os.environ.get("GOLDEN_LARGE", "1") == "1"


def main():
    ref = Reference()
    orc = Oracle()
    fixtures = {}
    cases = []

    def add(name, src, img, s, b, filters=True, tier="small"):
        t0 = time.time()
        px, rf = ref.optimize(img, s, b, filters)
        dt = time.time() - t0
        case = dict(name=name, src=src, w=int(img.shape[1]), h=int(img.shape[0]), strength=s,
                    bleed=b, filters=filters, in_sha=sha16(img), px_sha=sha16(px),
                    filt_sha=sha16(rf) if filters else None,
                    nsuap=filter_counts(rf) if filters else None, tier=tier,
                    ref_seconds=round(dt, 3))
        cases.append(case)
        print(f"{name:44s} s={s:3d} b={b:5d} f={int(filters)} {case['px_sha']} {case['filt_sha']} "
              f"{case['nsuap']} {dt:.2f}s", flush=True)

    def synth(w, h, seed, bpp=4):
        return to_bpp(orc.synth(w, h, seed), bpp), dict(kind="synth", w=w, h=h, seed=seed, bpp=bpp)

    # --- SURVEY 8c synthetic vectors --------------------------------------------------------
    img, src = synth(64, 32, 7)
    for s, b, f in [(20, 2, True), (20, 2, False), (85, 1, True), (255, 2, True),
                    (5, 32767, True), (0, 2, True), (1, 1, True), (31, 2, True), (32, 2, True),
                    (33, 2, True), (63, 3, True), (64, 1, False), (127, 2, True), (128, 2, True)]:
        add("synth64x32", src, img, s, b, f)
    for bpp in (3, 2, 1):
        img, src = synth(64, 32, 7, bpp)
        add(f"synth64x32_bpp{bpp}", src, img, 20, 2, True)
        add(f"synth64x32_bpp{bpp}", src, img, 20, 2, False)
        add(f"synth64x32_bpp{bpp}", src, img, 85, 1, True)
    for (w, h) in [(1, 1), (1, 16), (16, 1), (2, 2), (3, 5), (4, 4), (5, 3)]:
        img, src = synth(w, h, 3)
        add(f"synth{w}x{h}", src, img, 20, 2, True)
        add(f"synth{w}x{h}", src, img, 20, 2, False)
    img, src = synth(48, 40, 11)
    add("synth48x40", src, img, 33, 1, True)
    add("synth48x40", src, img, 60, 3, False)
    # widths around the kernel's 32-pixel tiles
    for w in (27, 28, 29, 31, 32, 33, 35, 36, 37, 59, 60, 61, 63, 64, 65, 95, 96, 97, 100, 127, 129):
        for bpp in (4, 3, 2, 1):
            img, src = synth(w, 9, 1000 + w, bpp)
            add(f"synth{w}x9_bpp{bpp}", src, img, 20, 2, True)
    img, src = synth(256, 128, 2)
    add("synth256x128", src, img, 20, 2, True)
    add("synth256x128", src, img, 40, 2, False)
    img, src = synth(1024, 512, 1)
    add("synth1024x512", src, img, 20, 2, True, tier="medium")
    if LARGE:
        img, src = synth(1920, 1080, 100)
        add("synth1920x1080", src, img, 20, 2, True, tier="large")
        img, src = synth(3840, 2160, 4)
        for s in (0, 20, 40, 85):
            add("synth3840x2160", src, img, s, 2, True, tier="large")

    # --- reference suite images (decoded to RGBA8 like the reference reader does) -------------
    def fixture(key, arr):
        fixtures[key] = np.ascontiguousarray(arr)
        return dict(kind="fixture", key=key)

    david = load_suite_rgba("david.png")
    src = fixture("david", david)
    for s, b, f in [(19, 2, True), (0, 2, True), (85, 2, True), (19, 2, False), (20, 2, True),
                    (40, 2, True)]:
        add("david", src, david, s, b, f)
    rose = load_suite_rgba("rose.png")
    src = fixture("rose", rose)
    add("rose", src, rose, 19, 2, True)
    add("rose", src, rose, 19, 2, False)
    add("rose", src, rose, 85, 1, True)
    lena = load_suite_rgba("lena.png")
    src = fixture("lena", lena)
    add("lena", src, lena, 20, 2, True, tier="medium")
    add("lena", src, lena, 19, 2, True, tier="medium")
    add("lena", src, lena, 20, 2, False, tier="medium")
    crops = {
        "tux": (slice(40, 168), slice(60, 188)),
        "redbrush": (slice(150, 246), slice(180, 308)),
        "dice": (slice(100, 196), slice(300, 460)),
        "girl": (slice(200, 264), slice(300, 428)),
        "ssr": (slice(300, 364), slice(400, 560)),
        "barbara": (slice(100, 196), slice(100, 228)),
        "parrots": (slice(100, 164), slice(200, 328)),
        "tenko": (slice(100, 164), slice(200, 328)),
    }
    for name, (ys, xs) in crops.items():
        full = load_suite_rgba(name + ".png")
        crop = np.ascontiguousarray(full[ys, xs])
        src = fixture(name + "_crop", crop)
        add(name + "_crop", src, crop, 19, 2, True)
        add(name + "_crop", src, crop, 40, 1, True)
        add(name + "_crop", src, crop, 19, 2, False)
    # gray + alpha (bpp 2) from a real image with transparency
    tux = load_suite_rgba("tux.png")
    ga = to_bpp(np.ascontiguousarray(tux[40:168, 60:188]), 2)
    src = fixture("tux_crop_ga", ga)
    add("tux_crop_ga", src, ga, 19, 2, True)
    add("tux_crop_ga", src, ga, 30, 2, False)

    # full suite images: hashes only (inputs are read from /root/reference/suite when present)
    if LARGE:
        for name, s, b in [("tux", 19, 2), ("redbrush", 19, 2), ("dice", 40, 1), ("girl", 19, 2),
                           ("ssr", 19, 2), ("barbara", 19, 2), ("parrots", 19, 2),
                           ("tenko", 20, 2)]:
            full = load_suite_rgba(name + ".png")
            add(name, dict(kind="suite", file=name + ".png"), full, s, b, True, tier="suite")

    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **fixtures)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py",
                       reference="oracle/_ref/libpngloss_ref.so built from /root/reference/src "
                                 "(optimize_state.c color_delta.c pngloss_image.c), gcc -g -O2",
                       cases=cases), f, indent=1)
    print(f"{len(cases)} cases, {len(fixtures)} fixtures")


if __name__ == "__main__":
    main()
