"""Access to the committed golden vectors (tests/golden/golden.json + fixtures.npz)."""
import json
import os

import numpy as np

from checkers import SUITE_DIR, Oracle, load_suite_rgba, to_bpp

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLDEN = json.load(f)["cases"]
_FIX = None


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
        if not os.path.isdir(SUITE_DIR):
            return None
        return load_suite_rgba(src["file"])
    raise ValueError(src)
