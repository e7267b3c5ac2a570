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
