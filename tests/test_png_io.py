"""CPU tests of the host program's PNG reader / writer (pngloss_b200/host/pl_png.c)
through pl_pngtool: decoding equals Pillow's for every colour type / bit depth / interlace, the writer
honours explicit per-row filters and the colour-type auto-detection of the reference
(src/rwpng.c:488-495, 557-613), chunks pass through, and option handling of the CLI keeps the
reference's exit codes (src/pngloss.c:94-160)."""
import io
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image, PngImagePlugin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pngloss_b200", "host")
TOOL = os.path.join(HOST, "pl_pngtool")
CLI = os.path.join(HOST, "pngloss")


@pytest.fixture(scope="module", autouse=True)
def _build():
    if not os.path.exists(os.path.join(ROOT, "pngloss_b200", "libpngloss_b200.so")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "pngloss_b200", "csrc")], check=True)
    subprocess.run(["make", "-s", "-C", HOST], check=True)


def decode(path, tmp_path):
    out = str(tmp_path / "d.rgba")
    subprocess.run([TOOL, "decode", path, out], check=True, stderr=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    w, h = struct.unpack("<II", raw[:8])
    return np.frombuffer(raw[8:], np.uint8).reshape(h, w, 4)


def encode(rgba, path, tmp_path, filters=None):
    src = str(tmp_path / "e.rgba")
    with open(src, "wb") as f:
        f.write(struct.pack("<II", rgba.shape[1], rgba.shape[0]))
        f.write(np.ascontiguousarray(rgba).tobytes())
    cmd = [TOOL, "encode", src, path]
    if filters is not None:
        fp = str(tmp_path / "f.bin")
        open(fp, "wb").write(bytes(filters))
        cmd.append(fp)
    subprocess.run(cmd, check=True)


def chunks_of(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, out = 8, []
    while pos < len(data):
        n, name = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(name + body)
        out.append((name.decode(), body))
        pos += 12 + n
    return out


def filter_bytes(path):
    ch = chunks_of(path)
    w, h, depth, ctype = struct.unpack(">IIBB", dict(ch)["IHDR"][:10])
    bpp = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    raw = zlib.decompress(b"".join(b for n, b in ch if n == "IDAT"))
    stride = 1 + w * bpp
    return [raw[y * stride] for y in range(h)], ctype


def pil_rgba(path):
    return np.array(Image.open(path).convert("RGBA"))


def save_interlaced(arr, path, ctype, depth=8):
    """Adam7 encoder for the tests (Pillow cannot write interlaced PNGs); filter type 0 everywhere."""
    h, w = arr.shape[:2]
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    a = arr.reshape(h, w, ch)
    raw = b""
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2),
                           (0, 1, 1, 2)]:
        sub = a[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        for row in sub:
            raw += b"\x00" + (row.astype(">u2").tobytes() if depth == 16 else row.astype(np.uint8).tobytes())
    def chunk(name, body):
        return struct.pack(">I", len(body)) + name + body + struct.pack(">I", zlib.crc32(name + body))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1)) +
                chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_decode_matches_pillow_all_types(tmp_path):
    rng = np.random.default_rng(3)
    h, w = 37, 53
    cases = {}
    cases["L8"] = Image.fromarray(rng.integers(0, 256, (h, w), dtype=np.uint8), "L")
    cases["L1"] = Image.fromarray((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), "L").convert("1")
    cases["LA"] = Image.fromarray(rng.integers(0, 256, (h, w, 2), dtype=np.uint8), "LA")
    cases["RGB"] = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
    cases["RGBA"] = Image.fromarray(rng.integers(0, 256, (h, w, 4), dtype=np.uint8), "RGBA")
    pal = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB").quantize(17)
    cases["P"] = pal
    for name, im in cases.items():
        p = str(tmp_path / f"{name}.png")
        im.save(p)
        assert np.array_equal(decode(p, tmp_path), pil_rgba(p)), name
    # palette with tRNS, RGB / gray with a transparent colour key
    p = str(tmp_path / "Pt.png")
    pal.save(p, transparency=bytes(rng.integers(0, 256, 9, dtype=np.uint8)))
    assert np.array_equal(decode(p, tmp_path), pil_rgba(p))
    rgb = np.array(cases["RGB"])
    p = str(tmp_path / "RGBt.png")
    cases["RGB"].save(p, transparency=tuple(int(v) for v in rgb[3, 4]))
    got = decode(p, tmp_path)
    assert np.array_equal(got, pil_rgba(p)) and got[3, 4, 3] == 0
    p = str(tmp_path / "Lt.png")
    cases["L8"].save(p, transparency=int(np.array(cases["L8"])[2, 2]))
    assert np.array_equal(decode(p, tmp_path), pil_rgba(p))
    # low bit depth palette / gray (Pillow packs 2- and 4-bit palettes with bits=)
    for bits in (1, 2, 4):
        q = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB").quantize(1 << bits)
        p = str(tmp_path / f"P{bits}.png")
        q.save(p, bits=bits)
        assert np.array_equal(decode(p, tmp_path), pil_rgba(p)), bits


def test_decode_16_bit_and_interlaced(tmp_path):
    rng = np.random.default_rng(4)
    h, w = 19, 23
    # 16-bit samples keep their high byte (png_set_strip_16, reference src/rwpng.c:226-228)
    g16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    p = str(tmp_path / "g16.png")
    Image.fromarray(g16, "I;16").save(p)
    got = decode(p, tmp_path)
    assert np.array_equal(got[..., 0], (g16 >> 8).astype(np.uint8)) and (got[..., 3] == 255).all()
    for ctype, ch in [(0, 1), (2, 3), (4, 2), (6, 4)]:
        for depth in (8, 16):
            a = rng.integers(0, 1 << depth, (h, w, ch)).astype(np.uint16 if depth == 16 else np.uint8)
            p = str(tmp_path / f"i{ctype}_{depth}.png")
            save_interlaced(a, p, ctype, depth)
            a8 = (a >> 8).astype(np.uint8) if depth == 16 else a
            want = np.zeros((h, w, 4), np.uint8)
            want[..., 3] = 255
            if ctype in (0, 4):
                want[..., 0] = want[..., 1] = want[..., 2] = a8[..., 0]
                if ctype == 4:
                    want[..., 3] = a8[..., 1]
            else:
                want[..., :ch] = a8
            assert np.array_equal(decode(p, tmp_path), want), (ctype, depth)


def test_writer_filters_and_colour_type(tmp_path):
    rng = np.random.default_rng(5)
    h, w = 21, 40
    base = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    masks = [0x08, 0x10, 0x20, 0x40, 0x80]
    for bpp, ctype in [(4, 6), (3, 2), (2, 4), (1, 0)]:
        img = base.copy()
        if bpp in (1, 2):
            img[..., 0] = img[..., 2] = img[..., 1]
        if bpp in (1, 3):
            img[..., 3] = 255
        filt = [masks[int(v)] for v in rng.integers(0, 5, h)]
        p = str(tmp_path / f"w{bpp}.png")
        encode(img, p, tmp_path, filt)
        got_filters, got_ctype = filter_bytes(p)
        assert got_ctype == ctype
        # rows >= 1 carry the caller's filter; row 0 is the heuristic's choice (reference rwpng.c:488-495)
        assert got_filters[1:] == [masks.index(m) for m in filt[1:]]
        assert np.array_equal(pil_rgba(p), img)
        assert np.array_equal(decode(p, tmp_path), img)
    # no explicit filters: every row by the heuristic; still a valid, identical image
    p = str(tmp_path / "heur.png")
    encode(base, p, tmp_path)
    assert np.array_equal(pil_rgba(p), base)


def test_chunks_pass_through_and_srgb(tmp_path):
    rng = np.random.default_rng(6)
    im = Image.fromarray(rng.integers(0, 256, (8, 9, 3), dtype=np.uint8), "RGB")
    meta = PngImagePlugin.PngInfo()
    meta.add_text("Comment", "kept by pngloss")
    p = str(tmp_path / "meta.png")
    im.save(p, pnginfo=meta, dpi=(300, 300))
    data = open(p, "rb").read()
    srgb = struct.pack(">I", 1) + b"sRGB\x00" + struct.pack(">I", zlib.crc32(b"sRGB\x00"))
    open(p, "wb").write(data[:33] + srgb + data[33:])       # right after IHDR
    out = str(tmp_path / "copy.png")
    subprocess.run([TOOL, "copy", p, out], check=True, stderr=subprocess.DEVNULL)
    names = [n for n, _ in chunks_of(out)]
    assert "tEXt" in names and "pHYs" in names
    assert names.index("gAMA") < names.index("IDAT") and "sRGB" in names    # reference rwpng.c:501-509
    assert dict(chunks_of(out))["gAMA"] == struct.pack(">I", 45455)
    assert np.array_equal(pil_rgba(out), pil_rgba(p))


def test_rejects_garbage(tmp_path):
    p = str(tmp_path / "bad.png")
    open(p, "wb").write(b"not a png at all")
    r = subprocess.run([TOOL, "decode", p, str(tmp_path / "o")], capture_output=True)
    assert r.returncode == 25                                   # LIBPNG_FATAL_ERROR
    good = str(tmp_path / "good.png")
    Image.fromarray(np.zeros((4, 4, 3), np.uint8), "RGB").save(good)
    data = bytearray(open(good, "rb").read())
    data[20] ^= 0xFF                                            # corrupt IHDR -> CRC mismatch on a critical chunk
    open(p, "wb").write(data)
    assert subprocess.run([TOOL, "decode", p, str(tmp_path / "o")], capture_output=True).returncode == 25


def test_cli_option_validation(tmp_path):
    """Exit codes and messages of the reference's option handling (src/pngloss.c:94-160,
    src/pngloss_opts.c:38-136); none of these reach the GPU."""
    def run(*args, **kw):
        return subprocess.run([CLI, *args], capture_output=True, text=True, **kw)
    assert run("-V").stdout.strip().startswith("1.0.1")
    r = run()
    assert r.returncode == 1 and "usage:" in r.stderr                        # MISSING_ARGUMENT
    r = run("-h")
    assert r.returncode == 0 and "--strength" in r.stdout
    r = run("-s", "300", "x.png")
    assert r.returncode == 4 and "range 0-255" in r.stderr                   # INVALID_ARGUMENT
    assert run("-s", "abc", "x.png").returncode == 4
    assert run("-b", "0", "x.png").returncode == 4
    assert run("-b", "40000", "x.png").returncode == 4
    r = run("--ext", "-a.png", "-o", "out.png", "x.png")
    assert r.returncode == 4 and "can't be used at the same time" in r.stderr
    assert run("-o", "a.png", "x.png", "y.png").returncode == 4
    assert run("-o", "-", "x.png", "y.png").returncode == 4
    r = run("-s", "5")
    assert r.returncode == 1 and "No input files specified." in r.stderr
    assert run("--bogus").returncode == 4
    r = run(str(tmp_path / "missing.png"))
    assert r.returncode == 2 and "cannot open" in r.stderr                   # READ_ERROR
    # existing output is not overwritten without --force (NOT_OVERWRITING_ERROR)
    src = str(tmp_path / "a.png")
    Image.fromarray(np.zeros((4, 4, 3), np.uint8), "RGB").save(src)
    open(str(tmp_path / "a-loss.png"), "wb").write(b"x")
    r = run(src)
    assert r.returncode == 15 and "not overwriting" in r.stderr
    assert open(str(tmp_path / "a-loss.png"), "rb").read() == b"x"
