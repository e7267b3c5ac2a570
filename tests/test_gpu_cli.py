"""GPU tests of the host program (pngloss_b200/host/pngloss): the reference's
command line surface end to end - PNG in, batched quantise + filter search on the GPU, PNG out - checked
against the oracle for pixels and per-row filters."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

from checkers import Oracle
from golden_cases import GOLDEN, load_input

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pngloss_b200", "host")
CLI = os.path.join(HOST, "pngloss")
MASK_TO_TYPE = {0x08: 0, 0x10: 1, 0x20: 2, 0x40: 3, 0x80: 4}


@pytest.fixture(scope="module", autouse=True)
def _build():
    subprocess.run(["make", "-s", "-C", HOST], check=True)


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def png_filter_bytes(path):
    data = open(path, "rb").read()
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        n, name = struct.unpack(">I4s", data[pos:pos + 8])
        if name == b"IHDR":
            ihdr = data[pos + 8:pos + 8 + n]
        if name == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    w, h, _, ctype = struct.unpack(">IIBB", ihdr[:10])
    bpp = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    raw = zlib.decompress(idat)
    return [raw[y * (1 + w * bpp)] for y in range(h)], ctype


def fixture_images(oracle):
    out = {}
    for c in GOLDEN:
        if c["src"]["kind"] == "fixture" and c["src"]["key"] not in out and c["src"]["key"] != "lena":
            out[c["src"]["key"]] = load_input(c, oracle)
    return out


def test_cli_batch_of_files_matches_oracle(tmp_path, oracle):
    imgs = fixture_images(oracle)
    paths = []
    for key, rgba in imgs.items():
        p = str(tmp_path / f"{key}.png")
        Image.fromarray(rgba, "RGBA").save(p)       # every input is an RGBA PNG; the path narrows it itself
        paths.append(p)
    # --jobs / --gpus are additions of this build: CPU threads for PNG decode/encode, files sharded over
    # all visible GPUs (one on the test box; the sharding itself is host logic)
    r = subprocess.run([CLI, "-v", "-s", "19", "-b", "2", "--jobs", "3", "--gpus", "0", "--", *paths],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"Compressed {len(paths)} images." in r.stderr
    for key, rgba in imgs.items():
        out = str(tmp_path / f"{key}-loss.png")
        want_px, want_rf = oracle.optimize(rgba, 19, 2, True)
        got = np.array(Image.open(out).convert("RGBA"))
        assert np.array_equal(got, want_px), key
        filt, ctype = png_filter_bytes(out)
        assert filt == [MASK_TO_TYPE[m] for m in want_rf], key    # row 0 too: heuristic == chosen filter
        gray = (want_px[..., 0] == want_px[..., 1]).all() and (want_px[..., 1] == want_px[..., 2]).all()
        opaque = (want_px[..., 3] == 255).all()
        assert ctype == {(1, 1): 0, (1, 0): 4, (0, 1): 2, (0, 0): 6}[(int(gray), int(opaque))], key
        assert os.path.getsize(out) < os.path.getsize(str(tmp_path / f"{key}.png"))


def test_cli_stdin_stdout_and_output_options(tmp_path, oracle):
    rgba = load_input([c for c in GOLDEN if c["name"] == "rose"][0], oracle)
    src = str(tmp_path / "rose.png")
    Image.fromarray(rgba, "RGBA").save(src)
    want_px, _ = oracle.optimize(rgba, 30, 1, True)
    # the website front end's invocation: pngloss -sN -bN - on stdin/stdout (reference pnglossapi.go:543-556)
    r = subprocess.run([CLI, "-s30", "-b1", "-"], input=open(src, "rb").read(), capture_output=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "stdout.png")
    open(out, "wb").write(r.stdout)
    assert np.array_equal(np.array(Image.open(out).convert("RGBA")), want_px)
    # -o and --ext
    dst = str(tmp_path / "explicit.png")
    assert subprocess.run([CLI, "-s", "30", "-b", "1", "-o", dst, src]).returncode == 0
    assert np.array_equal(np.array(Image.open(dst).convert("RGBA")), want_px)
    assert subprocess.run([CLI, "-s", "30", "-b", "1", "--ext", "-q.png", src]).returncode == 0
    assert np.array_equal(np.array(Image.open(str(tmp_path / "rose-q.png")).convert("RGBA")), want_px)
    # second run without --force refuses to overwrite, with --force it succeeds
    assert subprocess.run([CLI, "-s", "30", "-b", "1", "--ext", "-q.png", src],
                          capture_output=True).returncode == 15
    assert subprocess.run([CLI, "-f", "-s", "30", "-b", "1", "--ext", "-q.png", src]).returncode == 0


def test_cli_skip_if_larger(tmp_path, oracle):
    """--skip-if-larger: re-compressing the tool's own strength-0 output cannot make it smaller -> exit 98,
    no output file (reference src/pngloss.c:268-270, rwpng.c:85-105,626-628)."""
    img = oracle.synth(48, 32, 21)
    src = str(tmp_path / "in.png")
    Image.fromarray(img, "RGBA").save(src)
    first = str(tmp_path / "first.png")
    assert subprocess.run([CLI, "-s", "0", "-o", first, src]).returncode == 0
    assert np.array_equal(np.array(Image.open(first).convert("RGBA")), img)        # strength 0 is lossless
    r = subprocess.run([CLI, "-v", "-s", "0", "--skip-if-larger", first], capture_output=True, text=True)
    assert r.returncode == 98, r.stderr
    assert not os.path.exists(str(tmp_path / "first-loss.png"))
    assert not os.path.exists(str(tmp_path / "first-loss.png.tmp"))
    assert "Skipped 1 file" in r.stderr
    # and a clearly compressible case passes the same option
    assert subprocess.run([CLI, "-s", "40", "--skip-if-larger", src]).returncode == 0
    assert os.path.getsize(str(tmp_path / "in-loss.png")) < os.path.getsize(src)


def test_cli_gpu_scanlines_give_the_same_files_as_cpu_filtering(tmp_path, oracle):
    """The encoder normally receives filtered scanlines from the GPU (K4); with
    PNGLOSS_CPU_FILTER=1 it narrows and filters the quantised pixels itself, the way the reference's libpng
    does.  Both must produce byte-identical files, for every colour type."""
    from checkers import to_bpp
    imgs = dict(fixture_images(oracle))
    for bpp in (1, 2, 3, 4):
        imgs[f"synth{bpp}"] = to_bpp(oracle.synth(300, 41, 90 + bpp), bpp)
    outs = {}
    for mode in ("gpu", "cpu"):
        d = tmp_path / mode
        d.mkdir()
        paths = []
        for key, rgba in imgs.items():
            p = str(d / f"{key}.png")
            Image.fromarray(rgba, "RGBA").save(p)
            paths.append(p)
        env = dict(os.environ)
        if mode == "cpu":
            env["PNGLOSS_CPU_FILTER"] = "1"
        r = subprocess.run([CLI, "-s", "20", "--", *paths], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs[mode] = {key: open(str(d / f"{key}-loss.png"), "rb").read() for key in imgs}
    for key in imgs:
        assert outs["gpu"][key] == outs["cpu"][key], key


# ---- the reference's OWN main(): src/pngloss.c + src/pngloss_opts.c, unmodified, linked with the product --------
REFMAIN = os.path.join(ROOT, "oracle", "_ref", "pngloss_refmain")


def test_reference_main_linked_against_the_product(tmp_path, oracle):
    """oracle/_ref/pngloss_refmain is the reference's unmodified command line (its main(), its per-file loop,
    its optimize_with_rows call site src/pngloss.c:266, its option parser) compiled against the reference's
    own headers and linked with libpngloss_b200.so + the product's PNG reader / writer (oracle/Makefile refcli;
    built in the dev container, where /root/reference exists, and shipped prebuilt).  What it writes must
    decode to the oracle's pixels and carry the oracle's filter on every row - the drop-in, proven through the
    reference's own call site."""
    if not os.path.exists(REFMAIN):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refcli"], check=False)
    if not os.path.exists(REFMAIN):
        pytest.skip("oracle/_ref/pngloss_refmain not built (needs /root/reference)")
    imgs = fixture_images(oracle)
    paths = []
    for key, rgba in imgs.items():
        p = str(tmp_path / f"{key}.png")
        Image.fromarray(rgba, "RGBA").save(p)
        paths.append(p)
    r = subprocess.run([REFMAIN, "-v", "-s", "19", "-b", "2", *paths], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for key, rgba in imgs.items():
        out = str(tmp_path / f"{key}-loss.png")
        want_px, want_rf = oracle.optimize(rgba, 19, 2, True)
        got = np.array(Image.open(out).convert("RGBA"))
        assert np.array_equal(got, want_px), key
        filt, _ = png_filter_bytes(out)
        assert filt == [MASK_TO_TYPE[m] for m in want_rf], key
    # the stdin -> stdout mode the website front end execs (reference website/pnglossapi.go:543-556)
    rose = [c for c in GOLDEN if c["name"] == "rose"][0]
    rgba = load_input(rose, oracle)
    src = str(tmp_path / "rose_in.png")
    Image.fromarray(rgba, "RGBA").save(src)
    r = subprocess.run([REFMAIN, "-s30", "-b1", "-"], input=open(src, "rb").read(), capture_output=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "rose_out.png")
    open(out, "wb").write(r.stdout)
    want_px, _ = oracle.optimize(rgba, 30, 1, True)
    assert np.array_equal(np.array(Image.open(out).convert("RGBA")), want_px)


def test_cli_stdout_skip_if_larger_sends_one_complete_png(tmp_path, oracle):
    """stdout + --skip-if-larger with a result that is not smaller: stdout must carry exactly one complete PNG
    (the original), not the beginning of the rejected file followed by the original."""
    img = oracle.synth(48, 32, 21)
    src = str(tmp_path / "in.png")
    Image.fromarray(img, "RGBA").save(src)
    first = str(tmp_path / "first.png")
    assert subprocess.run([CLI, "-s", "0", "-o", first, src]).returncode == 0
    data = open(first, "rb").read()
    r = subprocess.run([CLI, "-s", "0", "--skip-if-larger", "-"], input=data, capture_output=True)
    assert r.returncode == 98, r.stderr
    assert r.stdout.count(b"\x89PNG\r\n\x1a\n") == 1 and r.stdout.count(b"IEND") == 1
    out = str(tmp_path / "stdout.png")
    open(out, "wb").write(r.stdout)
    assert np.array_equal(np.array(Image.open(out).convert("RGBA")), img)


def test_cli_chunks_under_a_small_host_budget(tmp_path, oracle):
    """PNGLOSS_HOST_BUDGET_MB=1 forces the file list through several decode / GPU / encode chunks; results and
    the verbose batch summary are those of a single pass."""
    imgs = {f"s{i}": oracle.synth(160 + 4 * i, 120, 300 + i) for i in range(7)}
    paths = []
    for key, rgba in imgs.items():
        p = str(tmp_path / f"{key}.png")
        Image.fromarray(rgba, "RGBA").save(p)
        paths.append(p)
    env = dict(os.environ, PNGLOSS_HOST_BUDGET_MB="1")
    r = subprocess.run([CLI, "-v", "-s", "20", "--", *paths], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "Compressed 7 images." in r.stderr
    assert "batch of 7 images on 1 GPU: used" in r.stderr
    total = 0
    for key, rgba in imgs.items():
        want_px, want_rf, tr = oracle.optimize(rgba, 20, 2, True, trace=True)
        got = np.array(Image.open(str(tmp_path / f"{key}-loss.png")).convert("RGBA"))
        assert np.array_equal(got, want_px), key
        total += int(tr["final_frequency"].sum())
    assert f"in {total} bytes" in r.stderr


@pytest.mark.skipif(__import__("pngloss_b200").device_count() < 2, reason="needs two GPUs")
def test_cli_two_gpus_reduce_the_batch_histogram_with_nccl(tmp_path, oracle):
    """--gpus 2: files sharded over two GPUs (one host thread and context each), results identical to the oracle,
    and the verbose batch line reports the symbol histogram summed over the GPUs by the library's NCCL call."""
    imgs = {f"g{i}": oracle.synth(96 + 8 * i, 64, 700 + i) for i in range(6)}
    paths = []
    for key, rgba in imgs.items():
        p = str(tmp_path / f"{key}.png")
        Image.fromarray(rgba, "RGBA").save(p)
        paths.append(p)
    r = subprocess.run([CLI, "-v", "-s", "20", "--gpus", "2", "--", *paths], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "batch of 6 images on 2 GPUs: used" in r.stderr
    assert "summed over the GPUs by NCCL" in r.stderr, r.stderr
    total = 0
    for key, rgba in imgs.items():
        want_px, want_rf, tr = oracle.optimize(rgba, 20, 2, True, trace=True)
        got = np.array(Image.open(str(tmp_path / f"{key}-loss.png")).convert("RGBA"))
        assert np.array_equal(got, want_px), key
        total += int(tr["final_frequency"].sum())
    assert f"in {total} bytes" in r.stderr
