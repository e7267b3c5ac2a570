#!/usr/bin/env python
"""Round-2 additions to the golden vectors (TEST INFRASTRUCTURE; run in the dev container, where
/root/reference exists, after `make -C oracle`):

  * tests/golden/suite_fixtures.npz - the eight full suite images of the tier "suite" goldens, decoded to
    RGBA8 the way the reference reader hands them to the hot path and stored with only the channels that carry
    information (gray images 1 channel, opaque colour 3), so that the GPU box - which has no /root/reference -
    can run those goldens too;
  * one full-size 8192 x 8192 golden (BASELINE configs[4]: synthetic seed 1000, strength 20, bleed 2), tier
    "huge", appended to golden.json: hashes of the UNMODIFIED reference's output (oracle/_ref).

    python tests/golden/make_golden_r2.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from checkers import Oracle, Reference, filter_counts, load_suite_rgba, sha16  # noqa: E402


def main():
    path = os.path.join(HERE, "golden.json")
    with open(path) as f:
        doc = json.load(f)
    cases = doc["cases"]
    packed = {}
    for c in cases:
        if c["tier"] != "suite":
            continue
        a = load_suite_rgba(c["src"]["file"])
        assert sha16(a) == c["in_sha"]
        gray = (a[..., 0] == a[..., 1]).all() and (a[..., 1] == a[..., 2]).all()
        opaque = (a[..., 3] == 255).all()
        chans = [1] if gray and opaque else [1, 3] if gray else [0, 1, 2] if opaque else [0, 1, 2, 3]
        packed[c["src"]["file"][:-4]] = np.ascontiguousarray(a[..., chans])
    np.savez_compressed(os.path.join(HERE, "suite_fixtures.npz"), **packed)
    print("suite_fixtures.npz:", {k: v.shape for k, v in packed.items()})

    if not any(c["tier"] == "huge" for c in cases):
        ref, orc = Reference(), Oracle()
        w = h = 8192
        img = orc.synth(w, h, 1000)
        t0 = time.time()
        px, rf = ref.optimize(img, 20, 2, True)
        dt = time.time() - t0
        cases.append(dict(name="synth8192x8192", src=dict(kind="synth", w=w, h=h, seed=1000, bpp=4), w=w, h=h,
                          strength=20, bleed=2, filters=True, in_sha=sha16(img), px_sha=sha16(px),
                          filt_sha=sha16(rf), nsuap=filter_counts(rf), tier="huge", ref_seconds=round(dt, 1)))
        print("8192x8192:", cases[-1]["px_sha"], cases[-1]["filt_sha"], cases[-1]["nsuap"], f"{dt:.0f}s")
        with open(path, "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
