"""ctypes access to the parity checkers (TEST INFRASTRUCTURE).

* ``Oracle``    -> oracle/liboracle.so, the CPU restatement (always available).
* ``Reference`` -> oracle/_ref/libpngloss_ref.so, the unmodified reference sources compiled by
  oracle/Makefile.  Built in the dev container; travels to the GPU box as a prebuilt file.

Nothing in the product imports this module.
"""
import ctypes
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpngloss_ref.so")
SUITE_DIR = "/root/reference/suite"

PNG_MASKS = (0x08, 0x10, 0x20, 0x40, 0x80)


def sha16(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def row_pointers(a: np.ndarray):
    h = a.shape[0]
    return (ctypes.c_void_p * h)(*[a.ctypes.data + y * a.strides[0] for y in range(h)])


def filter_counts(filters: np.ndarray):
    return [int((filters == m).sum()) for m in PNG_MASKS]


class OracleTrace(ctypes.Structure):
    _fields_ = [("row_costs", ctypes.c_void_p), ("row_strength", ctypes.c_void_p),
                ("final_frequency", ctypes.c_uint32 * 256),
                ("original_frequency", (ctypes.c_uint32 * 256) * 5)]


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_optimize_with_rows.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_void_p, ctypes.c_uint8, ctypes.c_long,
                                                ctypes.c_void_p]
        L.oracle_optimize_with_rows.restype = ctypes.c_int
        L.oracle_optimize_image.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p,
                                            ctypes.c_uint8, ctypes.c_long, ctypes.c_void_p]
        L.oracle_optimize_image.restype = ctypes.c_int
        L.oracle_original_frequency.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p]
        L.oracle_original_frequency.restype = None
        L.oracle_adaptive_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                             ctypes.c_uint32]
        L.oracle_adaptive_filter.restype = ctypes.c_int
        L.oracle_synth_rgba.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint64]
        L.oracle_synth_rgba.restype = None

    def synth(self, w, h, seed) -> np.ndarray:
        a = np.zeros((h, w, 4), np.uint8)
        self.lib.oracle_synth_rgba(a.ctypes.data, w, h, seed)
        return a

    def optimize(self, rgba, strength, bleed, want_filters=True, trace=False):
        """Returns (pixels, row_filters or None[, trace dict])."""
        a = np.ascontiguousarray(rgba).copy()
        h, w, _ = a.shape
        rf = np.zeros(h, np.uint8)
        tr = OracleTrace()
        costs = np.zeros((h, 5), np.uint64)
        strengths = np.zeros(h, np.uint8)
        if trace:
            tr.row_costs = costs.ctypes.data
            tr.row_strength = strengths.ctypes.data
        rc = self.lib.oracle_optimize_with_rows(row_pointers(a), w, h,
                                                rf.ctypes.data if want_filters else None,
                                                strength, bleed, ctypes.addressof(tr))
        if rc != 0:
            raise RuntimeError(f"oracle rc={rc}")
        out = (a, rf if want_filters else None)
        if trace:
            out += ({"row_costs": costs, "row_strength": strengths,
                     "final_frequency": np.ctypeslib.as_array(tr.final_frequency).copy(),
                     "original_frequency": np.ctypeslib.as_array(tr.original_frequency).copy()},)
        return out

    def original_frequency(self, packed: np.ndarray, bpp: int) -> np.ndarray:
        p = np.ascontiguousarray(packed)
        h = p.shape[0]
        w = p.shape[1] // bpp if p.ndim == 2 else p.shape[1]
        out = np.zeros((5, 256), np.uint32)
        self.lib.oracle_original_frequency(p.ctypes.data, w, h, bpp, w * bpp, out.ctypes.data)
        return out


class Reference:
    """The unmodified reference hot path (reference src/pngloss_image.h:21-25)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.optimize_with_rows.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_void_p, ctypes.c_bool, ctypes.c_uint8,
                                                ctypes.c_long]
        self.lib.optimize_with_rows.restype = ctypes.c_int
        self.lib.optimize_with_stride.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                  ctypes.c_uint32, ctypes.c_bool, ctypes.c_uint8,
                                                  ctypes.c_long]
        self.lib.optimize_with_stride.restype = None

    def original_frequency(self, packed: np.ndarray, bpp: int) -> np.ndarray:
        """original_frequency[5][256] as the reference's own optimize_state_init builds it
        (src/optimize_state.c:28-86) over a packed bytes_per_pixel image of shape (h, w * bpp)."""
        class RefImage(ctypes.Structure):          # reference src/pngloss_image.h:7-11
            _fields_ = [("rows", ctypes.POINTER(ctypes.c_void_p)), ("width", ctypes.c_uint32),
                        ("height", ctypes.c_uint32), ("bytes_per_pixel", ctypes.c_uint8)]

        class RefState(ctypes.Structure):          # reference src/optimize_state.h:9-16
            _fields_ = [("x", ctypes.c_uint32), ("y", ctypes.c_uint32), ("pixels", ctypes.c_void_p),
                        ("color_error", ctypes.c_void_p), ("symbol_frequency", ctypes.c_void_p),
                        ("symbol_count", ctypes.c_uint64),
                        ("original_frequency", ctypes.POINTER(ctypes.c_uint32) * 5)]
        p = np.ascontiguousarray(packed)
        h, wb = p.shape
        rows = (ctypes.c_void_p * h)(*[p.ctypes.data + y * wb for y in range(h)])
        img = RefImage(rows, wb // bpp, h, bpp)
        st = RefState()
        self.lib.optimize_state_init.argtypes = [ctypes.POINTER(RefState), ctypes.POINTER(RefImage)]
        self.lib.optimize_state_init.restype = ctypes.c_int
        self.lib.optimize_state_destroy.argtypes = [ctypes.POINTER(RefState)]
        self.lib.optimize_state_destroy.restype = None
        assert self.lib.optimize_state_init(ctypes.byref(st), ctypes.byref(img)) == 0
        out = np.stack([np.ctypeslib.as_array(st.original_frequency[f], shape=(256,)).copy() for f in range(5)])
        self.lib.optimize_state_destroy(ctypes.byref(st))
        return out

    def optimize(self, rgba, strength, bleed, want_filters=True):
        a = np.ascontiguousarray(rgba).copy()
        h, w, _ = a.shape
        rf = np.zeros(h, np.uint8)
        rc = self.lib.optimize_with_rows(row_pointers(a), w, h,
                                         rf.ctypes.data if want_filters else None,
                                         False, strength, bleed)
        if rc != 0:
            raise RuntimeError(f"reference rc={rc}")
        return a, (rf if want_filters else None)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def have_suite() -> bool:
    return os.path.isdir(SUITE_DIR)


def load_suite_rgba(name: str) -> np.ndarray:
    """Decode a reference suite PNG the way the reference reader hands it to the hot path
    (8-bit RGBA, reference src/rwpng.c:238-256)."""
    from PIL import Image
    return np.array(Image.open(os.path.join(SUITE_DIR, name)).convert("RGBA"), dtype=np.uint8)


# ---- input transforms that force each bytes-per-pixel path -------------------------
def to_bpp(rgba: np.ndarray, bpp: int) -> np.ndarray:
    a = rgba.copy()
    if bpp in (1, 2):
        a[..., 0] = a[..., 1]
        a[..., 2] = a[..., 1]
    if bpp in (1, 3):
        a[..., 3] = 255
    return a


# ---- PNG scanlines (K4): plain restatement of the PNG specification's filters -----------------------------
def png_narrow(rgba: np.ndarray):
    """RGBA8 (h, w, 4) -> (bytes_per_pixel, packed (h, w * bpp)) with the colour type the pixels allow, the
    way the reference's writer detects it (src/rwpng.c:557-613)."""
    gray = bool((rgba[..., 0] == rgba[..., 1]).all() and (rgba[..., 1] == rgba[..., 2]).all())
    opaque = bool((rgba[..., 3] == 255).all())
    chans = {(True, True): [1], (True, False): [1, 3], (False, True): [0, 1, 2], (False, False): [0, 1, 2, 3]}[
        (gray, opaque)]
    return len(chans), np.ascontiguousarray(rgba[..., chans]).reshape(rgba.shape[0], -1)


def png_filter_row(ftype: int, row: np.ndarray, prev, bpp: int) -> np.ndarray:
    row = row.astype(np.int32)
    up = np.zeros_like(row) if prev is None else prev.astype(np.int32)
    left = np.concatenate([np.zeros(bpp, np.int32), row[:-bpp]]) if row.size > bpp else np.zeros_like(row)
    ul = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if row.size > bpp else np.zeros_like(row)
    if ftype == 0:
        pred = np.zeros_like(row)
    elif ftype == 1:
        pred = left
    elif ftype == 2:
        pred = up
    elif ftype == 3:
        pred = (left + up) >> 1
    else:
        p = left + up - ul
        pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
        pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
    return ((row - pred) & 255).astype(np.uint8)


def png_heuristic_filter(row: np.ndarray, prev, bpp: int) -> int:
    """libpng's default: smallest sum of |signed residual|, first minimum in the order none .. paeth."""
    sums = []
    for f in range(5):
        r = png_filter_row(f, row, prev, bpp).astype(np.int64)
        sums.append(int(np.where(r < 128, r, 256 - r).sum()))
    return int(np.argmin(sums))


def png_scanlines(rgba: np.ndarray, row_filters: np.ndarray):
    """What the reference's writer hands to deflate for a quantised image: -> (bpp, row0 filter, (h, 1 + w*bpp))."""
    bpp, packed = png_narrow(rgba)
    h = packed.shape[0]
    out = np.zeros((h, 1 + packed.shape[1]), np.uint8)
    types = {0x08: 0, 0x10: 1, 0x20: 2, 0x40: 3, 0x80: 4}
    f0 = png_heuristic_filter(packed[0], None, bpp)
    for y in range(h):
        t = f0 if y == 0 else types[int(row_filters[y])]
        out[y, 0] = t
        out[y, 1:] = png_filter_row(t, packed[y], packed[y - 1] if y else None, bpp)
    return bpp, f0, out


def png_from_scanlines(w: int, h: int, bpp: int, scan: np.ndarray) -> bytes:
    """A complete PNG file around filtered scanlines (zlib here; the product's writer is pl_png.c)."""
    import struct
    import zlib

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[bpp]
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(scan.tobytes(), 6)) + chunk(b"IEND", b""))
