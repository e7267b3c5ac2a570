"""Access to the committed golden vectors (tests/golden/golden.json + fixtures.npz)."""
import json
import os

import numpy as np

from checkers import SUITE_DIR, Oracle, load_suite_rgba, to_bpp

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLDEN = json.load(f)["cases"]
_FIX = None
_SUITE_FIX = None


def case_id(c):
    return f"{c['name']}-s{c['strength']}-b{c['bleed']}-{'rf' if c['filters'] else 'null'}"


def cases(*tiers):
    return [c for c in GOLDEN if c["tier"] in tiers]


def load_input(c, oracle: Oracle = None) -> np.ndarray:
    global _FIX
    src = c["src"]
    if src["kind"] == "synth":
        oracle = oracle or Oracle()
        return to_bpp(oracle.synth(src["w"], src["h"], src["seed"]), src["bpp"])
    if src["kind"] == "fixture":
        if _FIX is None:
            _FIX = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
        return np.ascontiguousarray(_FIX[src["key"]])
    if src["kind"] == "suite":
        if os.path.isdir(SUITE_DIR):
            return load_suite_rgba(src["file"])
        # no reference tree (GPU box): the packed copy made by tests/golden/make_golden_r2.py
        global _SUITE_FIX
        fix = os.path.join(HERE, "golden", "suite_fixtures.npz")
        if not os.path.exists(fix):
            return None
        if _SUITE_FIX is None:
            _SUITE_FIX = np.load(fix)
        p = _SUITE_FIX[src["file"][:-4]]
        h, w, nch = p.shape
        a = np.empty((h, w, 4), np.uint8)
        if nch <= 2:                     # gray (+ alpha): G into R, G, B
            a[..., 0] = a[..., 1] = a[..., 2] = p[..., 0]
            a[..., 3] = p[..., 1] if nch == 2 else 255
        else:
            a[..., :3] = p[..., :3]
            a[..., 3] = p[..., 3] if nch == 4 else 255
        return a
    raise ValueError(src)
