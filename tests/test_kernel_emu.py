"""CPU tests of the *kernel source* (pngloss_b200/csrc/pl_kernels.cuh) executed on the SIMT emulator
in tests/simt_emu, compared with the oracle.  These check the kernel logic - lane mapping, fix-up
of the channel order, streamed error windows, histogram-delta cost, winner commit, retry - before any
GPU time is spent; the parity tests proper are the -m gpu tests that go through the C-ABI."""
import numpy as np
import pytest

import io

from checkers import Oracle, png_from_scanlines, png_scanlines, to_bpp
from emu import Emu


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def compare(emu, oracle, imgs, s, b, null_filters, lpc):
    got = emu.optimize(imgs, s, b, null_filters, lpc)
    total = np.zeros(256, np.uint64)
    for i, img in enumerate(imgs):
        px, rf, tr = oracle.optimize(img, s, b, not null_filters, trace=True)
        assert got["status"][i][0] == 0
        assert np.array_equal(got["pixels"][i], px), f"image {i} pixels"
        if not null_filters:
            assert np.array_equal(got["filters"][i], rf), f"image {i} filters"
        assert np.array_equal(got["final_hist"][i], tr["final_frequency"]), f"image {i} histogram"
        assert got["status"][i][2] == int((tr["row_strength"] != s).sum() and
                                          (s - tr["row_strength"].astype(int)).sum())
        total += tr["final_frequency"]
    assert np.array_equal(got["batch_hist"], total)
    return got


CASES = [  # w, h, seed, bpp, strength, bleed, null_filters
    (16, 6, 3, 4, 20, 2, False),
    (37, 9, 5, 4, 20, 2, False),
    (64, 7, 7, 3, 20, 2, False),
    (33, 8, 9, 2, 20, 2, False),
    (40, 8, 11, 1, 20, 2, False),
    (35, 7, 13, 4, 85, 1, True),
    (20, 5, 15, 4, 255, 2, False),
    (9, 9, 17, 4, 0, 2, False),
    (1, 1, 3, 4, 20, 2, False),
    (1, 16, 3, 4, 20, 2, False),
    (16, 1, 3, 4, 20, 2, False),
    (31, 4, 19, 4, 5, 32767, False),
    (66, 3, 21, 2, 40, 3, True),
]


BM = 16        # added to lpc: the bucket-maxima variant of K2 (see emu.py)
IN_PLACE = 32  # added to lpc: the kernel writes its output over its input


@pytest.mark.parametrize("lpc", [8, 4, 2, 1, BM + 8, BM + 2, BM + 1])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "w%d-h%d-seed%d-bpp%d-s%d-b%d-null%d" % c)
def test_emu_single_image(emu, oracle, case, lpc):
    w, h, seed, bpp, s, b, nf = case
    img = to_bpp(oracle.synth(w, h, seed), bpp)
    got = compare(emu, oracle, [img], s, b, nf, lpc)
    assert got["status"][0][1] == bpp or (w * h == 1)


@pytest.mark.parametrize("lpc", [8, 4, 2, 1, BM + 4, BM + 1, IN_PLACE + 8, IN_PLACE + BM + 1])
def test_emu_batch_mixed_modes(emu, oracle, lpc):
    """Several images per CTA (CPW = 8/lpc), every bytes-per-pixel mode, partially filled last CTA."""
    imgs = [to_bpp(oracle.synth(29, 6, 100 + i), (i % 4) + 1) for i in range(11)]
    compare(emu, oracle, imgs, 20, 2, False, lpc)


@pytest.mark.parametrize("lpc", [8, 2, 1, BM + 2, BM + 1, IN_PLACE + BM + 1])
def test_emu_retry_path(emu, oracle, lpc):
    """row_filters == NULL on tiny noisy images makes the libpng-heuristic check reject all five
    candidates now and then, which exercises the lower-strength retry (reference
    src/pngloss_image.c:211,273-274) - in some images of a CTA only."""
    rng = np.random.default_rng(7)
    imgs, retried = [], 0
    while len(imgs) < 16:
        img = rng.integers(0, 256, (3, 5, 4), dtype=np.uint8)
        img = to_bpp(img, int(rng.integers(1, 5)))
        _, _, tr = oracle.optimize(img, 20, 2, False, trace=True)
        hit = bool((tr["row_strength"] != 20).any())
        if hit or len(imgs) % 2 == 1:
            imgs.append(img)
            retried += hit
    assert retried >= 6
    got = compare(emu, oracle, imgs, 20, 2, True, lpc)
    assert (got["status"][:, 2] > 0).sum() == retried


def test_emu_noise_and_ties(emu, oracle):
    rng = np.random.default_rng(11)
    few = (rng.integers(0, 4, (6, 21, 4)) * 85).astype(np.uint8)      # few levels: many frequency ties
    noise = rng.integers(0, 256, (6, 21, 4), dtype=np.uint8)
    holes = noise.copy()
    holes[rng.random((6, 21)) < 0.3, 3] = 0                            # fully transparent pixels
    for lpc in (8, 2, BM + 1, BM + 4):
        compare(emu, oracle, [few, noise, holes, to_bpp(holes, 2)], 19, 2, False, lpc)
        compare(emu, oracle, [few, noise, holes, to_bpp(holes, 2)], 200, 1, True, lpc)


def test_emu_both_variants_of_each_path_ran(emu, oracle):
    """The kernel has a table and a computed variant of the Sierra taps and a skipped and a replayed
    variant of the channel fix-up; all four must have run (and stayed bit-exact)."""
    before = emu.counters()
    smooth = oracle.synth(40, 24, 5)
    compare(emu, oracle, [smooth], 20, 2, False, 8)          # lanes 8: taps from the 64-bit table
    compare(emu, oracle, [smooth, smooth], 20, 2, False, 1)  # lanes 1: taps from the 6-bit table
    rng = np.random.default_rng(3)
    noisy = rng.integers(0, 256, (8, 24, 4), dtype=np.uint8)
    compare(emu, oracle, [noisy], 255, 1, False, 8)          # errors beyond the range of either table:
    compare(emu, oracle, [noisy], 255, 1, False, 1)          # taps computed
    after = emu.counters()
    for key in ("taps_table", "taps_computed", "fixup_replay", "fixup_skipped"):
        assert after[key] > before[key], key


def test_emu_bucket_maxima_paths(emu, oracle):
    """Bucket-maxima variant: the look-up and the scan fall-back, the "+1" and the 64-bit-max table update
    must all have run; strengths around the smallest one that has a table (15), the largest, a band that
    wraps past +-128 (noise), clamped bands (values near 0 and 255) and every colour mode, bit-exact."""
    before = emu.counters()
    rng = np.random.default_rng(5)
    smooth = oracle.synth(48, 16, 5)
    dark = (rng.integers(0, 30, (10, 33, 4))).astype(np.uint8)            # clamped at 0
    bright = (255 - rng.integers(0, 30, (10, 33, 4))).astype(np.uint8)    # clamped at 255
    noise = rng.integers(0, 256, (10, 33, 4), dtype=np.uint8)
    for s in (14, 15, 16, 20, 31, 63, 127, 128, 255):
        compare(emu, oracle, [smooth], s, 2, False, BM + 1)
    for bpp in (1, 2, 3, 4):
        imgs = [to_bpp(x, bpp) for x in (dark, bright, noise)]
        compare(emu, oracle, imgs, 20, 2, False, BM + 1)
        compare(emu, oracle, imgs, 42, 1, True, BM + 2)
    after = emu.counters()
    for key in ("bm_lookup", "bm_scan", "bm_fast_update", "bm_general_update"):
        assert after[key] > before[key], key


def test_emu_bucket_maxima_width_limit(emu, oracle):
    """Relative counts of a row fit the table's 17-bit field only for widths below 16384: at and beyond that
    the bucket-maxima kernel must fall back to the scan by itself."""
    before = emu.counters()
    img = oracle.synth(16384, 1, 77)
    compare(emu, oracle, [img], 20, 2, False, BM + 1)
    after = emu.counters()
    assert after["bm_lookup"] == before["bm_lookup"] and after["bm_scan"] > before["bm_scan"]
    narrow = oracle.synth(16380, 1, 78)
    compare(emu, oracle, [narrow], 20, 2, False, BM + 1)
    assert emu.counters()["bm_lookup"] > after["bm_lookup"]


def test_emu_k1_histograms(emu, oracle):
    """K1's per-channel histograms, folded by colour mode, equal optimize_state_init's table."""
    for bpp in (1, 2, 3, 4):
        img = to_bpp(oracle.synth(23, 11, 40 + bpp), bpp)
        got = emu.optimize([img], 10, 2, False, 8)
        chans = {1: [1], 2: [1, 3], 3: [0, 1, 2], 4: [0, 1, 2, 3]}[bpp]
        folded = got["chan_hist"][0][:, chans, :].sum(axis=1)
        packed = np.ascontiguousarray(img[:, :, chans]).reshape(img.shape[0], -1)
        assert np.array_equal(folded, oracle.original_frequency(packed, bpp))


def test_emu_synth_matches_oracle(emu, oracle):
    for (w, h, seed) in [(64, 32, 7), (17, 70, 12345), (1, 1, 3)]:
        assert np.array_equal(emu.synth(w, h, seed), oracle.synth(w, h, seed))


def test_emu_k4_scanlines(emu, oracle):
    """K4: colour-type detection on the output, row-0 heuristic, per-row filters, narrowed and filtered
    scanlines - against a numpy restatement of the PNG filters, and through a real PNG decoder."""
    from PIL import Image
    rng = np.random.default_rng(21)
    for (w, h) in [(37, 9), (1, 5), (300, 4), (256, 3), (5, 1), (2052, 3), (1028, 2)]:
        for bpp in (1, 2, 3, 4):
            src = to_bpp(oracle.synth(w, h, 60 + bpp), bpp) if w > 1 else \
                to_bpp(rng.integers(0, 256, (h, w, 4), dtype=np.uint8), bpp)
            px, rf = oracle.optimize(src, 20, 2, True)
            rf = rf.copy()
            rf[1:] = rng.choice([0x08, 0x10, 0x20, 0x40, 0x80], size=h - 1)   # every filter type, any row
            (got_bpp, got_f0, got), = emu.scanlines([px], [rf])
            want_bpp, want_f0, want = png_scanlines(px, rf)
            assert (got_bpp, got_f0) == (want_bpp, want_f0), (w, h, bpp)
            assert np.array_equal(got, want), (w, h, bpp)
            dec = np.asarray(Image.open(io.BytesIO(png_from_scanlines(w, h, got_bpp, got))).convert("RGBA"))
            assert np.array_equal(dec, px), (w, h, bpp)


# ---- the lean kernel (pl_k2_lean.cuh): bulk-copy ring, 16-bit count increments, direct bucket update ----------
LEAN = 64 + BM + 1   # emu flag: pl_k2_lean (one lane per channel, bucket maxima)

LEAN_CASES = [  # w, h, seed, bpp, strength, bleed, null_filters  (widths are multiples of 4)
    (16, 6, 3, 4, 20, 2, False),     # one full tile
    (36, 9, 5, 4, 20, 2, False),     # three tiles, ragged tail of 4
    (64, 7, 7, 3, 20, 2, False),
    (32, 8, 9, 2, 20, 2, False),
    (40, 8, 11, 1, 20, 2, False),
    (100, 5, 13, 4, 85, 1, True),
    (20, 5, 15, 4, 255, 2, False),   # no bucket table at this strength: every byte scans
    (12, 9, 17, 4, 0, 2, False),
    (4, 1, 3, 4, 20, 2, False),
    (4, 16, 3, 4, 20, 2, False),
    (48, 1, 3, 4, 20, 2, False),
    (28, 4, 19, 4, 15, 32767, False),
    (68, 3, 21, 2, 40, 3, True),
    (132, 4, 23, 4, 126, 2, False),  # the largest strength with a table
]


@pytest.mark.parametrize("case", LEAN_CASES, ids=lambda c: "w%d-h%d-seed%d-bpp%d-s%d-b%d-null%d" % c)
def test_emu_lean_single_image(emu, oracle, case):
    w, h, seed, bpp, s, b, nf = case
    img = to_bpp(oracle.synth(w, h, seed), bpp)
    compare(emu, oracle, [img], s, b, nf, LEAN)


@pytest.mark.parametrize("flags", [LEAN, IN_PLACE + LEAN])
def test_emu_lean_batch_mixed_modes(emu, oracle, flags):
    """11 images = one full CTA of 8 and a partially filled one, every bytes-per-pixel mode."""
    imgs = [to_bpp(oracle.synth(52, 6, 100 + i), (i % 4) + 1) for i in range(11)]
    compare(emu, oracle, imgs, 20, 2, False, flags)


def test_emu_lean_retry_path(emu, oracle):
    rng = np.random.default_rng(7)
    imgs, retried = [], 0
    while len(imgs) < 16:
        img = rng.integers(0, 256, (3, 8, 4), dtype=np.uint8)
        img = to_bpp(img, int(rng.integers(1, 5)))
        _, _, tr = oracle.optimize(img, 20, 2, False, trace=True)
        hit = bool((tr["row_strength"] != 20).any())
        if hit or len(imgs) % 2 == 1:
            imgs.append(img)
            retried += hit
    assert retried >= 4
    got = compare(emu, oracle, imgs, 20, 2, True, LEAN)
    assert (got["status"][:, 2] > 0).sum() == retried


def test_emu_lean_noise_ties_and_paths(emu, oracle):
    """Frequency ties, noise (bands that wrap past +-128: the seam), fully transparent pixels, clamped bands;
    the look-up, the scan, the direct and the general bucket update must all have run."""
    before = emu.counters()
    rng = np.random.default_rng(11)
    few = (rng.integers(0, 4, (6, 24, 4)) * 85).astype(np.uint8)
    noise = rng.integers(0, 256, (6, 24, 4), dtype=np.uint8)
    holes = noise.copy()
    holes[rng.random((6, 24)) < 0.3, 3] = 0
    dark = (rng.integers(0, 30, (6, 24, 4))).astype(np.uint8)
    bright = (255 - rng.integers(0, 30, (6, 24, 4))).astype(np.uint8)
    smooth = oracle.synth(24, 6, 5)
    imgs = [few, noise, holes, to_bpp(holes, 2), dark, bright, smooth, to_bpp(noise, 1), to_bpp(dark, 3)]
    for s, b, nf in ((19, 2, False), (15, 1, False), (63, 3, True), (126, 2, False), (200, 1, True)):
        compare(emu, oracle, imgs, s, b, nf, LEAN)
    after = emu.counters()
    for key in ("bm_lookup", "bm_scan", "bm_fast_update", "bm_general_update", "fixup_replay", "fixup_skipped",
                "taps_table", "taps_computed"):
        assert after[key] > before[key], key
    # a band that the byte-range clamp empties collapses onto a value OUTSIDE its bucket (few grey levels, wide
    # bands): the chosen symbol must then be filed under its own bucket, not under the band's
    wide = (np.random.default_rng(99).integers(0, 6, (2, 256, 4)) * 51).astype(np.uint8)
    for s in (126, 120, 100):
        compare(emu, oracle, [wide], s, 2, False, LEAN)


def test_emu_lean_early_bulk_copies(oracle):
    """The same kernel with bulk copies that land the moment they are issued (the default lands them at the
    wait): a tile that is refilled while a warp still reads it would show up here."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from checkers import Oracle, to_bpp
        from emu import Emu
        emu, oracle = Emu(), Oracle()
        imgs = [to_bpp(oracle.synth(84, 5, 200 + i), (i %% 4) + 1) for i in range(9)]
        got = emu.optimize(imgs, 20, 2, False, %d)
        for i, img in enumerate(imgs):
            px, rf = oracle.optimize(img, 20, 2, True)
            assert np.array_equal(got["pixels"][i], px) and np.array_equal(got["filters"][i], rf), i
        print("ok")
    """) % (os.path.dirname(os.path.abspath(__file__)), LEAN)
    env = dict(os.environ, PL_EMU_BULK_EARLY="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_emu_lean_full_rgba_ctas(emu, oracle):
    """Eight live RGBA images per CTA: the all-lanes-active specialisation of the lean row pass (two full CTAs),
    with transparent holes, noise and the retry path (null filters) in the mix."""
    rng = np.random.default_rng(23)
    imgs = [oracle.synth(44, 5, 300 + i) for i in range(12)]
    for k in range(4):
        n = rng.integers(0, 256, (5, 44, 4), dtype=np.uint8)
        n[..., 3] = np.where(rng.random((5, 44)) < 0.2, 0, np.maximum(n[..., 3], 1))
        n[0, 0, 3] = 7   # stays on the RGBA path
        imgs.append(n)
    compare(emu, oracle, imgs, 20, 2, False, LEAN)
    compare(emu, oracle, imgs, 31, 1, True, LEAN)


# ---- the latency kernel (pl_k2_solo.cuh): one image per CTA, chain / producer / post warps ----------------------
SOLO5 = 128   # emu flag: pl_k2_solo<5> (one chain warp carries the five candidates)
SOLO1 = 256   # emu flag: pl_k2_solo<1> (one chain warp per candidate)
SOLOC = 512   # emu flag: pl_k2_solo<5, compact> (four warps: chain, producer, two post warps)

SOLO_CASES = CASES + [
    (32, 3, 31, 4, 20, 2, False),    # exactly one tile: the error rows' four extra cells need a tile of their own
    (29, 4, 33, 4, 20, 2, False),    # ... or fit the last tile (28 < 29 + 4 = 33 > 32: they do not)
    (28, 4, 35, 4, 20, 2, False),    # 28 + 4 = 32: they just fit
    (161, 5, 37, 4, 20, 2, False),   # more tiles than ring stages
    (130, 3, 39, 3, 63, 1, True),
    (132, 4, 23, 4, 126, 2, False),  # the largest strength with a table
    (100, 5, 13, 4, 85, 1, True),
]


# every case on the default layout, the other two layouts on the cases that differ in structure (widths around the tile
# size, every colour mode, NULL filters, strengths with and without a table)
SOLO_RUNS = [(c, SOLO5) for c in SOLO_CASES] + \
            [(c, f) for f in (SOLO1, SOLOC) for c in SOLO_CASES if c[:3] in
             {(37, 9, 5), (64, 7, 7), (33, 8, 9), (40, 8, 11), (35, 7, 13), (20, 5, 15), (9, 9, 17), (1, 1, 3),
              (32, 3, 31), (28, 4, 35), (161, 5, 37), (130, 3, 39)}]


@pytest.mark.parametrize("case,flag", SOLO_RUNS,
                         ids=lambda v: ("w%d-h%d-seed%d-bpp%d-s%d-b%d-null%d" % v) if isinstance(v, tuple) else
                         {SOLO5: "one-chain-warp", SOLO1: "five-chain-warps", SOLOC: "four-warp-cta"}[v])
def test_emu_solo_single_image(emu, oracle, case, flag):
    w, h, seed, bpp, s, b, nf = case
    img = to_bpp(oracle.synth(w, h, seed), bpp)
    got = compare(emu, oracle, [img], s, b, nf, flag)
    assert got["status"][0][1] == bpp or (w * h == 1)


@pytest.mark.parametrize("flag", [SOLO5, SOLO1, SOLOC, IN_PLACE + SOLO5, IN_PLACE + SOLO1, IN_PLACE + SOLOC])
def test_emu_solo_batch_mixed_modes(emu, oracle, flag):
    imgs = [to_bpp(oracle.synth(70, 6, 100 + i), (i % 4) + 1) for i in range(6)]
    compare(emu, oracle, imgs, 20, 2, False, flag)


@pytest.mark.parametrize("flag", [SOLO5, SOLO1, SOLOC])
def test_emu_solo_retry_path(emu, oracle, flag):
    rng = np.random.default_rng(7)
    imgs, retried = [], 0
    while len(imgs) < 12:
        img = rng.integers(0, 256, (3, 8, 4), dtype=np.uint8)
        img = to_bpp(img, int(rng.integers(1, 5)))
        _, _, tr = oracle.optimize(img, 20, 2, False, trace=True)
        hit = bool((tr["row_strength"] != 20).any())
        if hit or len(imgs) % 2 == 1:
            imgs.append(img)
            retried += hit
    assert retried >= 4
    got = compare(emu, oracle, imgs, 20, 2, True, flag)
    assert (got["status"][:, 2] > 0).sum() == retried


@pytest.mark.parametrize("flag", [SOLO5, SOLO1, SOLOC])
def test_emu_solo_noise_ties_and_paths(emu, oracle, flag):
    """Frequency ties, noise (the seam), fully transparent pixels, clamped bands; the look-up, the scan, both bucket
    updates, the fix-up replay and both tap paths must all have run."""
    before = emu.counters()
    rng = np.random.default_rng(11)
    few = (rng.integers(0, 4, (5, 36, 4)) * 85).astype(np.uint8)
    noise = rng.integers(0, 256, (5, 36, 4), dtype=np.uint8)
    holes = noise.copy()
    holes[rng.random((5, 36)) < 0.3, 3] = 0
    dark = (rng.integers(0, 30, (5, 36, 4))).astype(np.uint8)
    bright = (255 - rng.integers(0, 30, (5, 36, 4))).astype(np.uint8)
    smooth = oracle.synth(36, 5, 5)
    imgs = [few, noise, holes, to_bpp(holes, 2), dark, bright, smooth, to_bpp(noise, 1), to_bpp(dark, 3)]
    runs = ((19, 2, False), (63, 3, True), (126, 2, False), (200, 1, True)) if flag == SOLO5 else \
        ((19, 2, False), (200, 1, True))   # (the other layouts only differ in who does the helper work)
    for s, b, nf in runs:
        compare(emu, oracle, imgs, s, b, nf, flag)
    after = emu.counters()
    for key in ("bm_lookup", "bm_scan", "fixup_replay", "fixup_skipped", "taps_table", "solo_fast", "solo_general"):
        assert after[key] > before[key], key
    wide = (np.random.default_rng(99).integers(0, 6, (2, 256, 4)) * 51).astype(np.uint8)
    for s in ((126, 120, 100) if flag == SOLO5 else (126,)):
        compare(emu, oracle, [wide], s, 2, False, flag)


@pytest.mark.parametrize("flag", [SOLO5, SOLOC])
def test_emu_solo_small_strengths(emu, oracle, flag):
    """Strengths below 15: the other kernels have no winner table there; the latency kernel's has room for up to 259
    buckets per candidate (one symbol each at strength 0), so its fast path covers them too."""
    before = emu.counters()
    rng = np.random.default_rng(31)
    noise = rng.integers(0, 256, (5, 40, 4), dtype=np.uint8)
    holes = noise.copy()
    holes[rng.random((5, 40)) < 0.3, 3] = 0
    few = (rng.integers(0, 4, (5, 40, 4)) * 85).astype(np.uint8)
    edge = np.concatenate([rng.integers(0, 12, (5, 20, 4)), 255 - rng.integers(0, 12, (5, 20, 4))], axis=1).astype(np.uint8)
    imgs = [noise, holes, few, edge, oracle.synth(40, 5, 9), to_bpp(noise, 1), to_bpp(holes, 2), to_bpp(edge, 3)]
    for s, b, nf in ((0, 2, False), (1, 1, False), (7, 3, True), (14, 2, False)):
        compare(emu, oracle, imgs, s, b, nf, flag)
    after = emu.counters()
    assert after["solo_fast"] > before["solo_fast"]
