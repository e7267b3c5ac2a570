"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI of
libpngloss_b200.so; the oracle / golden vectors are only the checker.  Integer path: the bar is
bit-exact pixels, row filters and histograms."""
import numpy as np
import pytest

import pngloss_b200
from checkers import PNG_MASKS, Oracle, filter_counts, sha16, to_bpp
from golden_cases import case_id, cases, load_input

pytestmark = pytest.mark.gpu

import os as _os
ROOT_DIR = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def ctx():
    c = pngloss_b200.Context(0)
    yield c
    c.close()


def run_dropin(img, s, b, want_filters=True):
    got = img.copy()
    rf = np.zeros(img.shape[0], np.uint8) if want_filters else None
    rc = pngloss_b200.optimize_with_rows(got, rf, False, s, b)
    assert rc == 0
    return got, rf


# ---- golden vectors produced by the unmodified reference ------------------------------------------
@pytest.mark.parametrize("c", cases("small", "medium", "suite"), ids=case_id)
def test_golden_through_dropin_entry(oracle, c):
    img = load_input(c, oracle)
    if img is None:
        pytest.skip("suite image not available (tests/golden/suite_fixtures.npz)")
    assert sha16(img) == c["in_sha"]
    px, rf = run_dropin(img, c["strength"], c["bleed"], c["filters"])
    assert sha16(px) == c["px_sha"]
    if c["filters"]:
        assert filter_counts(rf) == c["nsuap"]
        assert sha16(rf) == c["filt_sha"]


@pytest.mark.parametrize("bm", [0, 1], ids=["scan", "bucket-maxima"])
@pytest.mark.parametrize("lanes", [8, 4, 2, 1])
def test_golden_batched_all_lane_mappings(ctx, oracle, lanes, bm):
    """All small golden cases of one (strength, bleed) in a single batch: mixed sizes, mixed colour
    modes, mixed row_filters/NULL, every lane mapping and both candidate-choice variants of the
    quantise kernel."""
    ctx.set_lanes(lanes)
    ctx.set_bucket_maxima(bm)
    groups = {}
    for c in cases("small"):
        groups.setdefault((c["strength"], c["bleed"]), []).append(c)
    for (s, b), cs in groups.items():
        imgs = [load_input(c, oracle).copy() for c in cs]
        rfs = [np.zeros(im.shape[0], np.uint8) if c["filters"] else None for c, im in zip(cs, imgs)]
        res = ctx.optimize_batch(imgs, rfs, s, b)
        for c, im, rf, r in zip(cs, imgs, rfs, res):
            assert r["status"] == 0
            assert sha16(im) == c["px_sha"], case_id(c)
            if c["filters"]:
                assert sha16(rf) == c["filt_sha"], case_id(c)
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_golden_large_4k_and_1080p(ctx, oracle):
    """BASELINE configs[2]: 3840x2160 RGBA at strength 0/20/40/85 (+ the 1080p vector), full size,
    against hashes produced by the reference."""
    for c in cases("large"):
        src = c["src"]
        batch = pngloss_b200.Batch(ctx, [src["w"]], [src["h"]])
        batch.synth(0, src["seed"])          # device generator == oracle generator (checked below)
        inp = np.zeros((src["h"], src["w"], 4), np.uint8)
        batch.download_input(0, inp)
        batch.run(c["strength"], c["bleed"])
        st, bpp, retried = batch.finish()
        out = np.zeros_like(inp)
        rf = np.zeros(src["h"], np.uint8)
        batch.download(0, out, rf)
        ctx.sync()
        assert sha16(inp) == c["in_sha"], "device generator drifted"
        assert st[0] == 0 and bpp[0] == 4
        assert sha16(out) == c["px_sha"], case_id(c)
        assert sha16(rf) == c["filt_sha"], case_id(c)
        assert int(batch.image_histogram(0).sum()) == src["w"] * src["h"] * 4
        batch.close()


# ---- fresh inputs against the oracle -----------------------------------------------------------------
def test_random_inputs_match_oracle(ctx, oracle):
    rng = np.random.default_rng(424242)
    for rep in range(6):
        imgs, params = [], []
        s = int(rng.choice([0, 1, 7, 19, 20, 40, 85, 200, 255]))
        b = int(rng.choice([1, 2, 3, 16, 32767]))
        for i in range(24):
            w = int(rng.integers(1, 150))
            h = int(rng.integers(1, 40))
            kind = i % 3
            if kind == 0:
                img = oracle.synth(w, h, 9000 + 100 * rep + i)
            elif kind == 1:
                img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            else:
                img = (rng.integers(0, 4, (h, w, 4)) * 85).astype(np.uint8)
            if rng.random() < 0.5:
                img[rng.random((h, w)) < 0.2, 3] = 0
            imgs.append(to_bpp(img, int(rng.integers(1, 5))))
            params.append(bool(rng.random() < 0.7))
        work = [im.copy() for im in imgs]
        rfs = [np.zeros(im.shape[0], np.uint8) if nf else None for im, nf in zip(imgs, params)]
        ctx.set_lanes([8, 4, 2, 1][rep % 4])
        ctx.set_bucket_maxima([1, 0, -1][rep % 3])
        res = ctx.optimize_batch(work, rfs, s, b)
        for im, got, rf, nf, r in zip(imgs, work, rfs, params, res):
            want_px, want_rf, tr = oracle.optimize(im, s, b, nf, trace=True)
            assert r["status"] == 0
            assert np.array_equal(got, want_px), (rep, im.shape, s, b, nf)
            if nf:
                assert np.array_equal(rf, want_rf), (rep, im.shape, s, b, nf)
            assert r["retried_rows"] == int((s - tr["row_strength"].astype(int)).sum())
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_retry_path_on_device(ctx, oracle):
    """row_filters == NULL on tiny noisy images: some rows find no libpng-consistent candidate and are
    retried at lower strength (reference src/pngloss_image.c:211,273-274)."""
    rng = np.random.default_rng(7)
    imgs = [to_bpp(rng.integers(0, 256, (3, 5, 4), dtype=np.uint8), int(rng.integers(1, 5)))
            for _ in range(64)]
    for lanes, bm in ((8, 0), (1, 0), (1, 1), (2, 1)):
        ctx.set_lanes(lanes)
        ctx.set_bucket_maxima(bm)
        work = [im.copy() for im in imgs]
        res = ctx.optimize_batch(work, None, 20, 2)
        total_retries = 0
        for im, got, r in zip(imgs, work, res):
            want_px, _, tr = oracle.optimize(im, 20, 2, False, trace=True)
            assert np.array_equal(got, want_px)
            assert r["retried_rows"] == int((20 - tr["row_strength"].astype(int)).sum())
            total_retries += r["retried_rows"]
        assert total_retries > 0
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_secondary_entry_points(oracle):
    """optimize_with_stride / optimizeForAverageFilter (reference src/pngloss_image.c:29-50):
    row_filters == NULL semantics, padded stride."""
    img = oracle.synth(50, 21, 31)
    want, _ = oracle.optimize(img, 30, 3, False)
    padded = np.zeros((21, 64, 4), np.uint8)
    padded[:, :50] = img
    view = padded[:, :50]
    pngloss_b200.optimize_with_stride(view, False, 30, 3)
    assert np.array_equal(view, want)
    assert not padded[:, 50:].any()
    want2, _ = oracle.optimize(img, 25, 2, False)
    tight = img.copy()
    pngloss_b200.optimizeForAverageFilter(tight, 25)
    assert np.array_equal(tight, want2)


def test_optimize_image_packed_entry(oracle):
    """The reference's inner entry on packed pixels with explicit bytes_per_pixel
    (src/pngloss_image.c:159), all four modes, against the oracle's restatement of it."""
    for bpp in (1, 2, 3, 4):
        rgba = oracle.synth(45, 13, 60 + bpp)
        chans = {1: [1], 2: [1, 3], 3: [0, 1, 2], 4: [0, 1, 2, 3]}[bpp]
        packed = np.ascontiguousarray(rgba[:, :, chans]).reshape(13, 45 * bpp)
        want = packed.copy()
        want_rf = np.zeros(13, np.uint8)
        assert oracle.lib.oracle_optimize_image(want.ctypes.data, 45, 13, bpp, 45 * bpp,
                                                want_rf.ctypes.data, 25, 2, None) == 0
        got = packed.copy()
        rf = np.zeros(13, np.uint8)
        assert pngloss_b200.optimize_image(got, bpp, rf, False, 25, 2) == 0
        assert np.array_equal(got, want) and np.array_equal(rf, want_rf), bpp


def test_non_contiguous_row_pointers(oracle):
    """rows[] need not be equally spaced."""
    import ctypes
    img = oracle.synth(40, 10, 77)
    want_px, want_rf = oracle.optimize(img, 20, 2, True)
    rows_mem = [np.ascontiguousarray(img[y]).copy() for y in range(10)]
    order = [3, 7, 1, 9, 0, 5, 2, 8, 6, 4]
    keep = [rows_mem[i] for i in order]  # scrambles allocation order, not image order
    ptrs = (ctypes.c_void_p * 10)(*[r.ctypes.data for r in rows_mem])
    rf = np.zeros(10, np.uint8)
    lib = pngloss_b200.load_library()
    assert lib.optimize_with_rows(ptrs, 40, 10, rf.ctypes.data, False, 20, 2) == 0
    assert np.array_equal(np.stack(rows_mem), want_px) and np.array_equal(rf, want_rf)
    del keep


def test_forced_bytes_per_pixel(ctx, oracle):
    """The reference's optimize_image takes an explicit bytes_per_pixel (src/pngloss_image.c:159); an
    opaque gray image forced to gray+alpha must quantise the alpha plane too."""
    import ctypes
    img = to_bpp(oracle.synth(33, 9, 5), 1)
    packed = np.ascontiguousarray(img[:, :, [1, 3]]).reshape(9, 66).copy()
    rf_want = np.zeros(9, np.uint8)
    rc = oracle.lib.oracle_optimize_image(packed.ctypes.data, 33, 9, 2, 66, rf_want.ctypes.data, 20, 2,
                                          None)
    assert rc == 0
    got = img.copy()
    rf = np.zeros(9, np.uint8)
    res = ctx.optimize_batch([got], [rf], 20, 2, force_bpp=2)
    assert res[0]["bytes_per_pixel"] == 2
    assert np.array_equal(got[:, :, 1], packed[:, 0::2]) and np.array_equal(got[:, :, 3], packed[:, 1::2])
    assert np.array_equal(got[:, :, 0], got[:, :, 1]) and np.array_equal(got[:, :, 2], got[:, :, 1])
    assert np.array_equal(rf, rf_want)


# ---- size-independent properties at full size ------------------------------------------------------
def test_properties_full_size_batch(ctx, oracle):
    """8 x 1920x1080 (BASELINE configs[3] shape): strength 0 is the identity; replicas of one image
    give identical outputs; filters are valid masks; histogram sums; batch histogram = sum."""
    w, h, n = 1920, 1080, 8
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    for i in range(n):
        batch.synth(i, 100 + (i % 2))     # two distinct images, replicated
    inp = [np.zeros((h, w, 4), np.uint8) for _ in range(2)]
    batch.download_input(0, inp[0])
    batch.download_input(1, inp[1])
    batch.run(0, 2)
    st, bpp, _ = batch.finish()
    out = np.zeros((h, w, 4), np.uint8)
    for i in range(2):
        batch.download(i, out, None)
        ctx.sync()
        assert np.array_equal(out, inp[i]), "strength 0 must be lossless"
    ctx.set_lanes(4)
    batch.run(20, 2)
    st, bpp, _ = batch.finish()
    assert (st == 0).all() and (bpp == 4).all()
    outs, rfs = [], []
    for i in range(n):
        o = np.zeros((h, w, 4), np.uint8)
        rf = np.zeros(h, np.uint8)
        batch.download(i, o, rf)
        outs.append(o)
        rfs.append(rf)
    ctx.sync()
    golden = [c for c in cases("large") if c["name"] == "synth1920x1080"][0]
    assert sha16(outs[0]) == golden["px_sha"] and sha16(rfs[0]) == golden["filt_sha"]
    total = np.zeros(256, np.uint64)
    for i in range(n):
        assert np.array_equal(outs[i], outs[i % 2]) and np.array_equal(rfs[i], rfs[i % 2])
        assert np.isin(rfs[i], PNG_MASKS).all()
        hist = batch.image_histogram(i)
        assert int(hist.sum()) == w * h * 4
        total += hist
        # fully transparent input pixels stay fully transparent (reference optimize_state.c:158-164)
        assert (outs[i][..., 3][inp[i % 2][..., 3] == 0] == 0).all()
    assert np.array_equal(batch.histogram(), total)
    info = batch.launch_info()
    assert info["images_per_cta"] == 2 and info["k2_ctas"] == 4
    ctx.set_lanes(0)
    batch.close()


def test_large_host_batch_matches_device_resident_path(ctx, oracle):
    """A few hundred MB through the host-buffer call must equal the device-resident batch path and the
    golden vector (mixed colour modes inside one batch)."""
    n, w, h = 160, 1024, 512
    imgs = [oracle.synth(w, h, 1 + i) for i in range(n)]
    imgs[5] = to_bpp(imgs[5], 3)
    imgs[9] = to_bpp(imgs[9], 1)
    work = [im.copy() for im in imgs]
    rfs = [np.zeros(h, np.uint8) for _ in range(n)]
    res = ctx.optimize_batch(work, rfs, 20, 2)
    assert all(r["status"] == 0 for r in res)
    assert res[5]["bytes_per_pixel"] == 3 and res[9]["bytes_per_pixel"] == 1
    golden = [c for c in cases("medium") if c["name"] == "synth1024x512"][0]
    assert sha16(work[0]) == golden["px_sha"] and sha16(rfs[0]) == golden["filt_sha"]
    ctx.set_lanes(2)
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    for i in range(n):
        batch.upload(i, imgs[i])
    batch.run(20, 2)
    st, _, _ = batch.finish()
    assert (st == 0).all()
    out = np.zeros((h, w, 4), np.uint8)
    rf = np.zeros(h, np.uint8)
    for i in range(n):
        batch.download(i, out, rf)
        ctx.sync()
        assert np.array_equal(out, work[i]) and np.array_equal(rf, rfs[i]), i
    batch.close()
    ctx.set_lanes(0)


def test_batch_larger_than_memory_budget_is_grouped(oracle):
    """A host batch that does not fit the device at once runs as consecutive groups
    (PNGLOSS_B200_MEM_BUDGET_MB shrinks the budget so that small inputs exercise it)."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import pngloss_b200
from checkers import Oracle
o = Oracle()
imgs = [o.synth(200, 120, 300 + i) for i in range(9)]          # ~0.3 MB of device memory each
work = [im.copy() for im in imgs]
rfs = [np.zeros(120, np.uint8) for _ in imgs]
ctx = pngloss_b200.Context(0)
res = ctx.optimize_batch(work, rfs, 20, 2)
for im, got, rf, r in zip(imgs, work, rfs, res):
    px, want_rf = o.optimize(im, 20, 2, True)
    assert r["status"] == 0 and np.array_equal(got, px) and np.array_equal(rf, want_rf)
print("grouped ok")
""" % (ROOT_DIR, os.path.join(ROOT_DIR, "tests"))
    env = dict(os.environ, PNGLOSS_B200_MEM_BUDGET_MB="1")        # 1 MB: three images per group
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "grouped ok" in r.stdout, r.stderr


def test_full_width_8192_matches_oracle(ctx, oracle):
    """BASELINE configs[4] uses 8192-pixel rows: full width, a strip of rows, exact against the oracle."""
    img = oracle.synth(8192, 48, 1000)
    want_px, want_rf = oracle.optimize(img, 20, 2, True)
    for lanes, bm in ((8, 0), (1, 0), (1, 1), (2, 1), (8, 1)):
        ctx.set_lanes(lanes)
        ctx.set_bucket_maxima(bm)
        got = img.copy()
        rf = np.zeros(48, np.uint8)
        res = ctx.optimize_batch([got], [rf], 20, 2)
        assert res[0]["status"] == 0
        assert np.array_equal(got, want_px) and np.array_equal(rf, want_rf), (lanes, bm)
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_4k_golden_eight_images_per_cta_bucket_maxima(ctx, oracle):
    """The bench's kernel variant (one lane per channel, 8 images per CTA, bucket maxima) on the full-size
    4K golden vector: 8 replicas in one CTA, every one must hash to the reference's output."""
    c = [c for c in cases("large") if c["src"]["w"] == 3840 and c["strength"] == 20][0]
    src = c["src"]
    n = 8
    ctx.set_lanes(1)
    ctx.set_bucket_maxima(1)
    batch = pngloss_b200.Batch(ctx, [src["w"]] * n, [src["h"]] * n)
    for i in range(n):
        batch.synth(i, src["seed"])
    batch.run(c["strength"], c["bleed"])
    st, bpp, _ = batch.finish()
    assert (st == 0).all() and (bpp == 4).all()
    out = np.zeros((src["h"], src["w"], 4), np.uint8)
    rf = np.zeros(src["h"], np.uint8)
    for i in range(n):
        batch.download(i, out, rf)
        ctx.sync()
        assert sha16(out) == c["px_sha"] and sha16(rf) == c["filt_sha"], i
    info = batch.launch_info()
    assert info["images_per_cta"] == 8 and info["k2_ctas"] == 1
    batch.close()
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_invalid_arguments(ctx):
    img = np.zeros((4, 4, 4), np.uint8)
    with pytest.raises(pngloss_b200.PnglossError):
        ctx.optimize_batch([img], None, 300, 2)
    with pytest.raises(pngloss_b200.PnglossError):
        ctx.optimize_batch([img], None, 20, 0)
    with pytest.raises(pngloss_b200.PnglossError):
        pngloss_b200.Batch(ctx, [0], [4])



def test_in_place_device_batch(ctx, oracle):
    """PNGLOSS_B200_BATCH_IN_PLACE: the kernel overwrites the uploaded rows (it keeps the one original row
    it still needs in a scratch row); results must not change."""
    imgs = [to_bpp(oracle.synth(97, 33, 50 + i), (i % 4) + 1) for i in range(12)]
    for lanes, bm in ((8, 0), (2, 1), (1, 1)):
        ctx.set_lanes(lanes)
        ctx.set_bucket_maxima(bm)
        batch = pngloss_b200.Batch(ctx, [97] * 12, [33] * 12, in_place=True)
        for i, im in enumerate(imgs):
            batch.upload(i, im)
        batch.run(20, 2)
        st, _, _ = batch.finish()
        assert (st == 0).all()
        for i, im in enumerate(imgs):
            want_px, want_rf = oracle.optimize(im, 20, 2, True)
            out, again = np.zeros_like(im), np.zeros_like(im)
            rf = np.zeros(33, np.uint8)
            batch.download(i, out, rf)
            batch.download_input(i, again)
            ctx.sync()
            assert np.array_equal(out, want_px) and np.array_equal(rf, want_rf), (lanes, bm, i)
            assert np.array_equal(again, want_px)
        batch.close()
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_jobs_in_flight_with_separate_outputs(ctx, oracle):
    """pngloss_b200_submit / _wait: three jobs in flight (more than the two device batches the library
    keeps), results into separate output buffers, inputs untouched, waited out of order."""
    rng = np.random.default_rng(99)
    jobs = []
    for j in range(3):
        imgs = []
        for i in range(10):
            im = oracle.synth(120, 40, 700 + 10 * j + i) if i % 2 == 0 else \
                rng.integers(0, 256, (40, 120, 4), dtype=np.uint8)
            imgs.append(to_bpp(im, (i % 4) + 1))
        pinned_in = ctx.pinned_empty((10, 40, 120, 4))
        pinned_out = ctx.pinned_empty((10, 40, 120, 4))
        for i, im in enumerate(imgs):
            pinned_in[i] = im
        rfs = [np.zeros(40, np.uint8) if i % 3 else None for i in range(10)]
        job = ctx.submit([pinned_in[i] for i in range(10)], rfs, 20 + j, 2,
                         outputs=[pinned_out[i] for i in range(10)])
        jobs.append((job, imgs, pinned_in, pinned_out, rfs, 20 + j))
    for k in (1, 0, 2):
        job, imgs, pinned_in, pinned_out, rfs, s = jobs[k]
        res = job.wait()
        for i, im in enumerate(imgs):
            want_px, want_rf = oracle.optimize(im, s, 2, rfs[i] is not None)
            assert res[i]["status"] == 0
            assert np.array_equal(pinned_in[i], im), "input buffer was modified"
            assert np.array_equal(pinned_out[i], want_px), (k, i)
            if rfs[i] is not None:
                assert np.array_equal(rfs[i], want_rf), (k, i)
    for _, _, a, b, _, _ in jobs:
        ctx.free_pinned(a)
        ctx.free_pinned(b)



def test_k4_filtered_scanlines(ctx, oracle):
    """pngloss_b200_batch_scanlines: the filtered scanlines of the results equal a numpy restatement of the
    PNG filters applied to the downloaded results, and a PNG decoder turns them back into those pixels."""
    import io

    from PIL import Image

    from checkers import png_from_scanlines, png_scanlines
    imgs = [to_bpp(oracle.synth(131, 29, 80 + i), (i % 4) + 1) for i in range(8)]
    imgs.append(np.full((29, 131, 4), 200, np.uint8))        # flat gray, opaque
    batch = pngloss_b200.Batch(ctx, [131] * len(imgs), [29] * len(imgs))
    for i, im in enumerate(imgs):
        batch.upload(i, im)
    batch.run(20, 2)
    batch.scanlines()
    st, _, _ = batch.finish()
    assert (st == 0).all()
    for i, im in enumerate(imgs):
        out = np.zeros_like(im)
        rf = np.zeros(29, np.uint8)
        batch.download(i, out, rf)
        ctx.sync()
        info = batch.scanline_info(i)
        got = batch.download_scanlines(i)
        want_bpp, want_f0, want = png_scanlines(out, rf)
        assert (info["bytes_per_pixel"], info["row0_filter"]) == (want_bpp, want_f0), i
        assert np.array_equal(got, want), i
        dec = np.asarray(Image.open(io.BytesIO(png_from_scanlines(131, 29, want_bpp, got))).convert("RGBA"))
        assert np.array_equal(dec, out), i
    batch.close()
    # full-size 4K image
    c = [c for c in cases("large") if c["src"]["w"] == 3840 and c["strength"] == 20][0]
    src = c["src"]
    batch = pngloss_b200.Batch(ctx, [src["w"]], [src["h"]])
    batch.synth(0, src["seed"])
    batch.run(c["strength"], c["bleed"])
    batch.scanlines()
    batch.finish()
    out = np.zeros((src["h"], src["w"], 4), np.uint8)
    rf = np.zeros(src["h"], np.uint8)
    batch.download(0, out, rf)
    ctx.sync()
    assert sha16(out) == c["px_sha"]
    got = batch.download_scanlines(0)
    want_bpp, want_f0, want = png_scanlines(out, rf)
    assert batch.scanline_info(0)["bytes_per_pixel"] == want_bpp == 4
    assert np.array_equal(got, want)
    batch.close()


def test_scanlines_through_the_host_buffer_call(ctx, oracle):
    """pngloss_b200_image.scanlines: the blocking batch call returns filtered scanlines next to (or instead of)
    the pixels."""
    from checkers import png_scanlines
    imgs = [to_bpp(oracle.synth(64, 20, 300 + i), (i % 4) + 1) for i in range(6)]
    for no_pixels in (False, True):
        work = [im.copy() for im in imgs]
        rfs = [np.zeros(20, np.uint8) for _ in imgs]
        scans = [np.zeros(20 * (1 + 4 * 64), np.uint8) for _ in imgs]
        res = ctx.optimize_batch(work, rfs, 20, 2, scanlines=scans, no_pixels=no_pixels)
        for i, im in enumerate(imgs):
            want_px, want_rf = oracle.optimize(im, 20, 2, True)
            want_bpp, want_f0, want = png_scanlines(want_px, want_rf)
            assert res[i]["status"] == 0
            assert (res[i]["scan_bytes_per_pixel"], res[i]["scan_row0_filter"]) == (want_bpp, want_f0)
            assert res[i]["scan_bytes"] == want.size
            assert np.array_equal(scans[i][:want.size].reshape(want.shape), want), i
            assert np.array_equal(rfs[i], want_rf)
            assert np.array_equal(work[i], im if no_pixels else want_px), "pixels"


def test_very_wide_rows_fall_back_to_the_scan(ctx, oracle):
    """Widths of 16384 and more do not fit the bucket table's relative counts: same results, scan kernel
    (chosen by the host) or in-kernel fall-back (variant forced)."""
    img = oracle.synth(16400, 3, 55)
    want_px, want_rf = oracle.optimize(img, 20, 2, True)
    for lanes, bm in ((1, -1), (1, 1), (8, -1)):
        ctx.set_lanes(lanes)
        ctx.set_bucket_maxima(bm)
        got = img.copy()
        rf = np.zeros(3, np.uint8)
        res = ctx.optimize_batch([got], [rf], 20, 2)
        assert res[0]["status"] == 0
        assert np.array_equal(got, want_px) and np.array_equal(rf, want_rf), (lanes, bm)
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)


def test_host_calls_with_changing_shapes_and_small_memory_budget(oracle):
    """The host-buffer calls recycle device batches: alternate shapes between calls (the pool must not hand out
    a batch of the wrong shape), then squeeze the memory budget so that one call runs as several pipelined
    groups - with scanlines requested, whose device buffer counts against the budget too."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import pngloss_b200
from checkers import Oracle, png_scanlines
o = Oracle()
ctx = pngloss_b200.Context(0)
want = {}
for rep in range(3):
    for (w, h, n) in [(64, 20, 5), (48, 31, 3), (64, 20, 5), (200, 9, 7)]:
        imgs = [o.synth(w, h, 1000 + 10 * i + w) for i in range(n)]
        work = [im.copy() for im in imgs]
        rfs = [np.zeros(h, np.uint8) for _ in imgs]
        scans = [np.zeros(h * (1 + 4 * w), np.uint8) for _ in imgs]
        res = ctx.optimize_batch(work, rfs, 20, 2, scanlines=scans)
        for i, im in enumerate(imgs):
            key = (w, h, i)
            if key not in want:
                px, rf = o.optimize(im, 20, 2, True)
                want[key] = (px, rf, png_scanlines(px, rf))
            px, rf, (bpp, f0, sc) = want[key]
            assert res[i]["status"] == 0 and np.array_equal(work[i], px) and np.array_equal(rfs[i], rf), key
            assert res[i]["scan_bytes_per_pixel"] == bpp and res[i]["scan_bytes"] == sc.size
            assert np.array_equal(scans[i][:sc.size].reshape(sc.shape), sc), key
print("pool ok")
""" % (ROOT_DIR, os.path.join(ROOT_DIR, "tests"))
    for budget in (None, "1"):
        env = dict(os.environ)
        if budget:
            env["PNGLOSS_B200_MEM_BUDGET_MB"] = budget
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert r.returncode == 0 and "pool ok" in r.stdout, (budget, r.stderr[-2000:])


# ---- the large-batch kernel (pl_k2_lean.cuh) ---------------------------------------------------------------
@pytest.mark.parametrize("lean", [1, 0], ids=["lean", "generic"])
def test_lean_kernel_golden_and_random(ctx, oracle, lean):
    """The large-batch kernel (pl_k2_lean: bulk-copy tile ring, 16-bit count increments, three CTAs per SM)
    against the reference's goldens and the oracle: the goldens whose width is a multiple of 4 in one batch per
    (strength, bleed) (mixed sizes, colour modes, NULL filters), 19 random images of every mode in full and
    ragged CTAs, and the same through the generic kernel."""
    ctx.set_lanes(1)
    ctx.set_bucket_maxima(1)
    ctx.set_lean(lean)
    groups = {}
    for c in cases("small", "medium"):
        groups.setdefault((c["strength"], c["bleed"]), []).append(c)
    for (s, b), cs in groups.items():
        cs = [c for c in cs if load_input(c, oracle).shape[1] % 4 == 0]
        if not cs:
            continue
        imgs = [load_input(c, oracle).copy() for c in cs]
        rfs = [np.zeros(im.shape[0], np.uint8) if c["filters"] else None for c, im in zip(cs, imgs)]
        res = ctx.optimize_batch(imgs, rfs, s, b)
        for c, im, rf, r in zip(cs, imgs, rfs, res):
            assert r["status"] == 0
            assert sha16(im) == c["px_sha"], case_id(c)
            if c["filters"]:
                assert sha16(rf) == c["filt_sha"], case_id(c)
    rng = np.random.default_rng(99)
    ran_lean = 0
    for (w, h, s, b) in [(64, 40, 20, 2), (100, 17, 19, 1), (36, 33, 63, 3), (260, 9, 126, 2), (48, 12, 15, 2)]:
        n = 19
        imgs = []
        for i in range(n):
            kind = i % 3
            if kind == 0:
                a = oracle.synth(w, h, 500 + i)
            elif kind == 1:
                a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
                a[rng.random((h, w)) < 0.2, 3] = 0
            else:
                a = (rng.integers(0, 6, (h, w, 4)) * 51).astype(np.uint8)
            imgs.append(to_bpp(a, 4 if i < 9 else (i % 4) + 1))
        batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
        for i, a in enumerate(imgs):
            batch.upload(i, a)
        batch.run(s, b)
        st, _, _ = batch.finish()
        assert (st == 0).all()
        info = batch.launch_info()
        assert info["lean"] == bool(lean) and info["images_per_cta"] == 8
        ran_lean += info["lean"]
        out = np.zeros((h, w, 4), np.uint8)
        rf = np.zeros(h, np.uint8)
        for i, a in enumerate(imgs):
            batch.download(i, out, rf)
            ctx.sync()
            px, want_rf = oracle.optimize(a, s, b, True)
            assert np.array_equal(out, px) and np.array_equal(rf, want_rf), (w, h, s, b, i)
        batch.close()
    assert ran_lean == (5 if lean else 0)
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)
    ctx.set_lean(-1)


def test_lean_kernel_4k_golden_24_images(ctx, oracle):
    """Three full CTAs of the lean kernel on the full-size 4K golden vector, as an in-place batch: replicas must
    hash to the reference's output (pixels and row filters)."""
    c = [c for c in cases("large") if c["src"]["w"] == 3840 and c["strength"] == 20][0]
    src = c["src"]
    n = 24
    ctx.set_lanes(1)
    ctx.set_bucket_maxima(1)
    ctx.set_lean(1)
    batch = pngloss_b200.Batch(ctx, [src["w"]] * n, [src["h"]] * n, in_place=True)
    for i in range(n):
        batch.synth(i, src["seed"])
    batch.run(c["strength"], c["bleed"])
    st, bpp, _ = batch.finish()
    assert (st == 0).all() and (bpp == 4).all()
    assert batch.launch_info()["lean"]
    out = np.zeros((src["h"], src["w"], 4), np.uint8)
    rf = np.zeros(src["h"], np.uint8)
    for i in range(0, n, 5):
        batch.download(i, out, rf)
        ctx.sync()
        assert sha16(out) == c["px_sha"] and sha16(rf) == c["filt_sha"], i
    batch.close()
    ctx.set_lanes(0)
    ctx.set_bucket_maxima(-1)
    ctx.set_lean(-1)


# ---- the latency kernel (pl_k2_solo.cuh) -------------------------------------------------------------------------
@pytest.mark.parametrize("solo", [1, 2, 3], ids=["one-chain-warp", "five-chain-warps", "four-warp-cta"])
def test_solo_kernel_golden_and_random(ctx, oracle, solo):
    """The latency kernel (pl_k2_solo: one image per CTA, chain / producer / post warps, fast path + general path)
    against the reference's goldens - every small and medium vector of strength 0 .. 126, through the
    drop-in entry and in batches - and against the oracle on random images of every mode (noise, transparent
    holes, few grey levels, dark and bright: the clamped bands and the channel replay)."""
    ctx.set_lanes(8)
    ctx.set_solo(solo)
    ran = 0
    for c in cases("small", "medium"):
        if c["strength"] > 126:
            continue
        img = load_input(c, oracle).copy()
        rf = np.zeros(img.shape[0], np.uint8) if c["filters"] else None
        res, = ctx.optimize_batch([img], [rf], c["strength"], c["bleed"])
        assert res["status"] == 0
        assert sha16(img) == c["px_sha"], case_id(c)
        if c["filters"]:
            assert sha16(rf) == c["filt_sha"], case_id(c)
        ran += 1
    assert ran >= 20
    rng = np.random.default_rng(199)
    for (w, h, s, b) in [(64, 40, 20, 2), (101, 17, 19, 1), (37, 33, 63, 3), (260, 9, 126, 2), (48, 12, 15, 2),
                         (300, 20, 85, 2), (31, 50, 40, 1), (90, 12, 0, 2), (64, 16, 3, 1), (75, 10, 14, 2)]:
        n = 10
        imgs = []
        for i in range(n):
            kind = i % 5
            if kind == 0:
                a = oracle.synth(w, h, 500 + i)
            elif kind == 1:
                a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
                a[rng.random((h, w)) < 0.2, 3] = 0
            elif kind == 2:
                a = (rng.integers(0, 6, (h, w, 4)) * 51).astype(np.uint8)
            elif kind == 3:
                a = rng.integers(0, 30, (h, w, 4)).astype(np.uint8)
            else:
                a = (255 - rng.integers(0, 30, (h, w, 4))).astype(np.uint8)
            imgs.append(to_bpp(a, 4 if i < 5 else (i % 4) + 1))
        batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n, in_place=bool(w & 1))
        for i, a in enumerate(imgs):
            batch.upload(i, a)
        batch.run(s, b)
        st, _, _ = batch.finish()
        assert (st == 0).all()
        info = batch.launch_info()
        assert info["solo"] and info["images_per_cta"] == 1
        out = np.zeros((h, w, 4), np.uint8)
        rf = np.zeros(h, np.uint8)
        for i, a in enumerate(imgs):
            batch.download(i, out, rf)
            ctx.sync()
            px, want_rf = oracle.optimize(a, s, b, True)
            assert np.array_equal(out, px) and np.array_equal(rf, want_rf), (w, h, s, b, i)
        batch.close()
    # NULL filters (every row adaptive: the retry path, also below the table's strength range)
    tiny = [to_bpp(rng.integers(0, 256, (3, 8, 4), dtype=np.uint8), int(rng.integers(1, 5))) for _ in range(24)]
    outs = [t.copy() for t in tiny]
    res = ctx.optimize_batch(outs, [None] * len(tiny), 20, 2)
    for t, o, r in zip(tiny, outs, res):
        px, _ = oracle.optimize(t, 20, 2, False)
        assert r["status"] == 0 and np.array_equal(o, px)
    ctx.set_lanes(0)
    ctx.set_solo(-1)


@pytest.mark.parametrize("solo", [1, 3], ids=["one-chain-warp", "four-warp-cta"])
def test_solo_kernel_suite_images(ctx, oracle, solo):
    """The eight full suite images (tier "suite" goldens of the unmodified reference) through the latency kernel."""
    ctx.set_lanes(8)
    ctx.set_solo(solo)
    cs = [c for c in cases("suite") if c["strength"] <= 126]
    assert cs
    for c in cs:
        img = load_input(c, oracle).copy()
        rf = np.zeros(img.shape[0], np.uint8) if c["filters"] else None
        res, = ctx.optimize_batch([img], [rf], c["strength"], c["bleed"])
        assert res["status"] == 0
        assert sha16(img) == c["px_sha"], case_id(c)
        if c["filters"]:
            assert sha16(rf) == c["filt_sha"], case_id(c)
    ctx.set_lanes(0)
    ctx.set_solo(-1)


def test_kernel_choice_by_batch_size(ctx, oracle):
    """The library's own choice of the quantise kernel (pl_api.cu use_solo / choose_lpc): the latency kernel while
    every image can have a CTA of its own with at most four CTAs per SM and the strength is at most 126, the generic
    kernel (several images per CTA) beyond, and whenever a lane mapping is set explicitly."""
    sms = 148
    img = oracle.synth(16, 4, 3)
    want_px, want_rf = oracle.optimize(img, 20, 2, True)

    def run(n, strength=20):
        batch = pngloss_b200.Batch(ctx, [16] * n, [4] * n)
        for i in range(n):
            batch.upload(i, img)
        batch.run(strength, 2)
        st, _, _ = batch.finish()
        assert (st == 0).all()
        info = batch.launch_info()
        if strength == 20:
            out = np.zeros_like(img)
            rf = np.zeros(4, np.uint8)
            batch.download(n - 1, out, rf)
            ctx.sync()
            assert np.array_equal(out, want_px) and np.array_equal(rf, want_rf)
        batch.close()
        return info

    ctx.set_lanes(0)
    ctx.set_solo(-1)
    assert run(1)["solo"] and run(2 * sms)["solo"] and run(4 * sms)["solo"]
    assert run(1, 0)["solo"] and run(1, 126)["solo"]
    big = run(4 * sms + 1)
    assert not big["solo"] and big["images_per_cta"] > 1
    assert not run(1, 127)["solo"]
    ctx.set_lanes(8)
    assert not run(1)["solo"]            # an explicit lane mapping keeps the generic kernel
    ctx.set_solo(0)
    ctx.set_lanes(0)
    assert not run(1)["solo"]
    ctx.set_solo(-1)


def test_full_8192x8192_golden(ctx, oracle):
    """BASELINE configs[4]'s image size in full: one 8192 x 8192 synthetic image (seed 1000, strength 20) through
    the device-resident batch must hash to what the unmodified reference produced (tier "huge" golden, made by
    tests/golden/make_golden_r2.py; 268 MB in, 268 MB out)."""
    huge = cases("huge")
    if not huge:
        pytest.skip("no 8192 x 8192 golden in golden.json")
    c = huge[0]
    src = c["src"]
    batch = pngloss_b200.Batch(ctx, [src["w"]], [src["h"]], in_place=True)
    batch.synth(0, src["seed"])
    batch.run(c["strength"], c["bleed"])
    st, bpp, _ = batch.finish()
    assert (st == 0).all() and (bpp == 4).all()
    out = np.zeros((src["h"], src["w"], 4), np.uint8)
    rf = np.zeros(src["h"], np.uint8)
    batch.download(0, out, rf)
    ctx.sync()
    batch.close()
    assert sha16(rf) == c["filt_sha"] and filter_counts(rf) == c["nsuap"]
    assert sha16(out) == c["px_sha"]


def test_k1_original_histogram_against_the_reference(ctx, oracle):
    """The histogram kernel K1 on the device against the reference's own optimize_state_init
    (src/optimize_state.c:66-83, called in oracle/_ref; the restatement where that library is absent): per filter,
    K1's per-channel counts summed over the channels of the colour mode are original_frequency[5][256]."""
    from checkers import Reference, have_reference
    ref = Reference() if have_reference() else None
    for bpp in (1, 2, 3, 4):
        img = to_bpp(oracle.synth(203, 57, 80 + bpp), bpp)
        batch = pngloss_b200.Batch(ctx, [203], [57])
        batch.upload(0, img)
        batch.run(20, 2)
        st, got_bpp, _ = batch.finish()
        assert st[0] == 0 and got_bpp[0] == bpp
        chan = batch.original_histogram(0)
        batch.close()
        chans = {1: [1], 2: [1, 3], 3: [0, 1, 2], 4: [0, 1, 2, 3]}[bpp]
        folded = chan[:, chans, :].sum(axis=1)
        packed = np.ascontiguousarray(img[:, :, chans]).reshape(img.shape[0], -1)
        want = ref.original_frequency(packed, bpp) if ref else oracle.original_frequency(packed, bpp)
        assert np.array_equal(folded, want), bpp


def test_dropin_calls_reuse_their_batch_and_run_from_threads(oracle):
    """optimize_with_rows keeps a context and the last device batch per calling thread: repeated calls with the
    same size, a size change in between, and two threads calling at once must all give the oracle's result."""
    import threading
    a = oracle.synth(64, 40, 900)
    b = oracle.synth(64, 40, 901)
    c = oracle.synth(52, 33, 902)
    want = {id(x): oracle.optimize(x, 20, 2, True) for x in (a, b, c)}
    for img in (a, b, a, c, b):
        px, rf = run_dropin(img, 20, 2)
        assert np.array_equal(px, want[id(img)][0]) and np.array_equal(rf, want[id(img)][1])
    errors = []

    def worker(imgs):
        try:
            for _ in range(3):
                for img in imgs:
                    px, rf = run_dropin(img, 20, 2)
                    assert np.array_equal(px, want[id(img)][0]) and np.array_equal(rf, want[id(img)][1])
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(imgs,)) for imgs in ((a, c), (b, a))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
