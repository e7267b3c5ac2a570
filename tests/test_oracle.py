"""CPU tests: the oracle restatement is pinned against (a) the golden vectors produced by the
unmodified reference and (b) the compiled reference itself on fresh random inputs."""
import numpy as np
import pytest

from checkers import (Oracle, Reference, filter_counts, have_reference, sha16, to_bpp)
from golden_cases import case_id, cases, load_input


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def check_case(oracle, c):
    img = load_input(c, oracle)
    if img is None:
        pytest.skip("reference suite not present")
    assert sha16(img) == c["in_sha"], "input generator / fixture drifted"
    px, rf = oracle.optimize(img, c["strength"], c["bleed"], c["filters"])
    assert sha16(px) == c["px_sha"]
    if c["filters"]:
        assert sha16(rf) == c["filt_sha"]
        assert filter_counts(rf) == c["nsuap"]


@pytest.mark.parametrize("c", cases("small"), ids=case_id)
def test_oracle_matches_golden_small(oracle, c):
    check_case(oracle, c)


@pytest.mark.parametrize("c", cases("medium", "suite"), ids=case_id)
def test_oracle_matches_golden_medium(oracle, c):
    check_case(oracle, c)


@pytest.mark.slow
@pytest.mark.parametrize("c", cases("large"), ids=case_id)
def test_oracle_matches_golden_large(oracle, c):
    check_case(oracle, c)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")
def test_oracle_matches_reference_random(oracle):
    """Fresh random shapes / strengths / bleeds / bpp paths, byte-for-byte against the reference."""
    ref = Reference()
    rng = np.random.default_rng(20260101)
    for i in range(60):
        w = int(rng.integers(1, 80))
        h = int(rng.integers(1, 40))
        bpp = int(rng.integers(1, 5))
        kind = i % 3
        if kind == 0:
            img = oracle.synth(w, h, 5000 + i)
        elif kind == 1:
            img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)          # white noise
        else:
            img = (rng.integers(0, 4, (h, w, 4)) * 85).astype(np.uint8)    # few levels, many ties
        if rng.random() < 0.5:
            img[rng.random((h, w)) < 0.2, 3] = 0                            # transparent holes
        img = to_bpp(img, bpp)
        s = int(rng.choice([0, 1, 2, 7, 19, 20, 40, 85, 200, 255]))
        b = int(rng.choice([1, 2, 3, 16, 32767]))
        nf = bool(rng.random() < 0.7)
        a = ref.optimize(img, s, b, nf)
        o = oracle.optimize(img, s, b, nf)
        assert np.array_equal(a[0], o[0]), (i, w, h, bpp, s, b, nf)
        if nf:
            assert np.array_equal(a[1], o[1]), (i, w, h, bpp, s, b, nf)


def test_strength_zero_is_identity(oracle):
    img = oracle.synth(40, 20, 9)
    px, _ = oracle.optimize(img, 0, 2, True)
    assert np.array_equal(px, img)


def test_trace_cost_identity(oracle):
    """Row cost restructuring used on the GPU: the bit cost of a row equals
    sum_s dcount[s] * (33 + clz32(freq_end[s])).  Checked through the trace."""
    img = oracle.synth(33, 12, 77)
    px, rf, tr = oracle.optimize(img, 20, 2, True, trace=True)
    assert tr["final_frequency"].sum() == 33 * 12 * 4
    assert (tr["row_costs"].min(axis=1) < np.iinfo(np.uint64).max).all()
    assert (tr["row_strength"] == 20).all()


def test_ulog2_identity():
    """The kernel's bit cost uses 33 + clz32(f) for the reference's ulog2(UINTMAX_MAX / f)
    (src/optimize_state.c:338,564-572; ulog2 returns the bit length).  Exhaustive near powers of two,
    random elsewhere."""
    rng = np.random.default_rng(5)
    fs = set(int(x) for x in rng.integers(1, 2**31 - 1, 20000))
    for k in range(31):
        for d in (-2, -1, 0, 1, 2):
            f = (1 << k) + d
            if 1 <= f < 2**31:
                fs.add(f)
    fs.update(range(1, 5000))
    for f in fs:
        clz32 = 32 - f.bit_length()
        assert ((2**64 - 1) // f).bit_length() == 33 + clz32, f


def test_int16_wrap_is_unreachable(oracle):
    """The error cells are int16 and the reference relies on wrap-on-store (src/color_delta.h:6); the kernels
    keep that narrowing explicitly.  No input can actually reach it: every error cell is a convex combination of
    earlier (here - back) differences (the ten Sierra taps of a pixel sum to the difference, a cell's incoming
    weights sum to 32/32), and |here - back| never exceeds max(|error|, 255 + strength), so cells stay within
    about +-511.  Adversarial inputs at the extreme settings confirm it: zero wraps, and the restatement still
    equals the reference."""
    import ctypes
    from checkers import Reference, have_reference
    oracle.lib.oracle_int16_wraps.restype = ctypes.c_uint64
    oracle.lib.oracle_int16_wraps.argtypes = [ctypes.c_int]
    rng = np.random.default_rng(1)
    h, w = 6, 400
    checker = np.zeros((h, w, 4), np.uint8)
    checker[(np.add.outer(np.arange(h), np.arange(w)) % 2) == 0] = 255
    checker[..., 3] = np.where(checker[..., 0] > 0, 255, 1)
    stripes = np.full((h, w, 4), 255, np.uint8)
    stripes[:, ::7] = 0
    stripes[..., 3] = 200
    noise = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    ref = Reference() if have_reference() else None
    for img in (checker, stripes, noise):
        for (s, b) in ((255, 1), (200, 1), (85, 1)):
            oracle.lib.oracle_int16_wraps(1)
            px, rf = oracle.optimize(img, s, b, True)
            assert oracle.lib.oracle_int16_wraps(1) == 0
            if ref is not None:
                px2, rf2 = ref.optimize(img, s, b, True)
                assert np.array_equal(px, px2) and np.array_equal(rf, rf2)


def test_original_frequency_equals_the_reference_init(oracle):
    """optimize_state_init's original_frequency[5][256] (reference src/optimize_state.c:66-83), called in the
    reference library itself, against the restatement - for every bytes-per-pixel mode.  (The emulator and GPU
    tests compare the histogram kernel K1 with the restatement / with this.)"""
    from checkers import Reference, have_reference, to_bpp
    if not have_reference():
        pytest.skip("oracle/_ref not built")
    ref = Reference()
    for bpp in (1, 2, 3, 4):
        img = to_bpp(oracle.synth(37, 13, 70 + bpp), bpp)
        chans = {1: [1], 2: [1, 3], 3: [0, 1, 2], 4: [0, 1, 2, 3]}[bpp]
        packed = np.ascontiguousarray(img[:, :, chans]).reshape(img.shape[0], -1)
        assert np.array_equal(ref.original_frequency(packed, bpp), oracle.original_frequency(packed, bpp)), bpp
