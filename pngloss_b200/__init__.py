"""pngloss_b200 - Python access to the B200-native quantise + filter-search path of pngloss.

This package is a thin ctypes binding over the C ABI in include/pngloss_b200.h (the product is
pngloss_b200/libpngloss_b200.so: hand-written sm_100a kernels + a C host shim).  It exists for the
tests and bench.py; C callers link the library directly.

There is no CPU implementation here: importing works without a GPU (so the ABI can be inspected),
but every compute entry point raises if the library or a CUDA device is missing.
"""
import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNGLOSS_B200_LIB selects an alternative build of the same library (tuning experiments only)
LIB_PATH = os.environ.get("PNGLOSS_B200_LIB") or os.path.join(_HERE, "libpngloss_b200.so")

SUCCESS = 0
INVALID_ARGUMENT = 4
OUT_OF_MEMORY = 17
DEVICE_ERROR = 40
NO_ACCEPTABLE_ROW = 41

# every symbol include/pngloss_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "optimize_with_rows", "optimize_with_stride", "optimizeForAverageFilter", "optimize_image",
    "pngloss_b200_device_count", "pngloss_b200_ctx_create", "pngloss_b200_ctx_destroy",
    "pngloss_b200_ctx_error", "pngloss_b200_ctx_set_lanes", "pngloss_b200_ctx_set_bucket_maxima",
    "pngloss_b200_ctx_set_lean", "pngloss_b200_ctx_set_solo", "pngloss_b200_ctx_timer_start",
    "pngloss_b200_ctx_timer_stop", "pngloss_b200_ctx_sync", "pngloss_b200_host_alloc",
    "pngloss_b200_host_free", "pngloss_b200_optimize_batch", "pngloss_b200_submit", "pngloss_b200_wait",
    "pngloss_b200_batch_create", "pngloss_b200_batch_create_ex",
    "pngloss_b200_batch_destroy", "pngloss_b200_batch_set_mode", "pngloss_b200_batch_upload",
    "pngloss_b200_batch_upload_rows", "pngloss_b200_batch_synth", "pngloss_b200_batch_run",
    "pngloss_b200_batch_download", "pngloss_b200_batch_download_rows",
    "pngloss_b200_batch_download_input", "pngloss_b200_batch_finish",
    "pngloss_b200_batch_image_histogram", "pngloss_b200_batch_image_original_histogram",
    "pngloss_b200_batch_histogram",
    "pngloss_b200_batch_histogram_device", "pngloss_b200_batch_timings",
    "pngloss_b200_batch_launch_info", "pngloss_b200_batch_scanlines", "pngloss_b200_batch_scanline_info",
    "pngloss_b200_batch_download_scanlines",
    "pngloss_b200_comm_unique_id", "pngloss_b200_comm_init_rank", "pngloss_b200_comm_init_all",
    "pngloss_b200_comm_destroy", "pngloss_b200_comm_size", "pngloss_b200_batch_allreduce_histogram",
    "pngloss_b200_comm_allreduce_u64", "pngloss_b200_ctx_flush_l2", "pngloss_b200_ctx_set_pipeline", "pngloss_b200_ctx_symbol_histogram",
]


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """NCCL unique id for pngloss_b200_comm_init_rank (rank 0 creates it and hands it to the other ranks)."""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = load_library().pngloss_b200_comm_unique_id(buf)
    if rc:
        raise PnglossError(rc, "pngloss_b200_comm_unique_id (is libnccl.so.2 loadable?)")
    return buf.raw


def comm_init_all(contexts: "Sequence[Context]"):
    """One process driving several GPUs: one communicator over the given contexts (one per device)."""
    arr = (ctypes.c_void_p * len(contexts))(*[c.handle for c in contexts])
    rc = load_library().pngloss_b200_comm_init_all(arr, len(contexts))
    if rc:
        raise PnglossError(rc, contexts[0].lib.pngloss_b200_ctx_error(contexts[0].handle).decode())


class PnglossError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"pngloss_b200 error {code}: {msg}")
        self.code = code


class PnglossImage(ctypes.Structure):
    """struct pngloss_image (reference src/pngloss_image.h:7-11)."""
    _fields_ = [("rows", ctypes.POINTER(ctypes.c_void_p)), ("width", ctypes.c_uint32),
                ("height", ctypes.c_uint32), ("bytes_per_pixel", ctypes.c_uint8)]


class ImageDesc(ctypes.Structure):
    """struct pngloss_b200_image (include/pngloss_b200.h)."""
    _fields_ = [("pixels", ctypes.c_void_p), ("stride", ctypes.c_size_t),
                ("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("row_filters", ctypes.c_void_p), ("force_bytes_per_pixel", ctypes.c_uint32),
                ("bytes_per_pixel", ctypes.c_uint32), ("retried_rows", ctypes.c_uint32),
                ("status", ctypes.c_int), ("out_pixels", ctypes.c_void_p), ("out_stride", ctypes.c_size_t),
                ("scanlines", ctypes.c_void_p), ("flags", ctypes.c_uint32),
                ("scan_bytes_per_pixel", ctypes.c_uint32), ("scan_row0_filter", ctypes.c_uint32),
                ("scan_bytes", ctypes.c_size_t)]


_lib = None
u32t = ctypes.c_uint32


def load_library() -> ctypes.CDLL:
    """Load libpngloss_b200.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnglossError(DEVICE_ERROR, f"{LIB_PATH} is missing: run `make -C pngloss_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, sz, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_int
    L.optimize_with_rows.argtypes = [vp, u32, u32, vp, ctypes.c_bool, ctypes.c_uint8, ctypes.c_long]
    L.optimize_with_rows.restype = i32
    L.optimize_with_stride.argtypes = [vp, u32, u32, u32, ctypes.c_bool, ctypes.c_uint8, ctypes.c_long]
    L.optimize_with_stride.restype = None
    L.optimize_image.argtypes = [ctypes.POINTER(PnglossImage), vp, ctypes.c_bool, ctypes.c_uint8,
                                 ctypes.c_long]
    L.optimize_image.restype = i32
    L.optimizeForAverageFilter.argtypes = [vp, i32, i32, i32]
    L.optimizeForAverageFilter.restype = None
    L.pngloss_b200_device_count.restype = i32
    L.pngloss_b200_ctx_create.argtypes = [ctypes.POINTER(vp), i32, vp]
    L.pngloss_b200_ctx_destroy.argtypes = [vp]
    L.pngloss_b200_ctx_destroy.restype = None
    L.pngloss_b200_ctx_error.argtypes = [vp]
    L.pngloss_b200_ctx_error.restype = ctypes.c_char_p
    L.pngloss_b200_ctx_set_lanes.argtypes = [vp, i32]
    L.pngloss_b200_ctx_set_bucket_maxima.argtypes = [vp, i32]
    L.pngloss_b200_ctx_set_lean.argtypes = [vp, i32]
    L.pngloss_b200_ctx_set_solo.argtypes = [vp, i32]
    L.pngloss_b200_ctx_timer_start.argtypes = [vp]
    L.pngloss_b200_ctx_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.pngloss_b200_ctx_sync.argtypes = [vp]
    L.pngloss_b200_host_alloc.argtypes = [sz]
    L.pngloss_b200_host_alloc.restype = vp
    L.pngloss_b200_host_free.argtypes = [vp]
    L.pngloss_b200_host_free.restype = None
    L.pngloss_b200_optimize_batch.argtypes = [vp, ctypes.POINTER(ImageDesc), sz, ctypes.c_uint,
                                              ctypes.c_long]
    L.pngloss_b200_batch_create.argtypes = [vp, sz, vp, vp, ctypes.POINTER(vp)]
    L.pngloss_b200_submit.argtypes = [vp, ctypes.POINTER(ImageDesc), sz, ctypes.c_uint, ctypes.c_long,
                                      ctypes.POINTER(vp)]
    L.pngloss_b200_wait.argtypes = [vp]
    L.pngloss_b200_batch_create_ex.argtypes = [vp, sz, vp, vp, ctypes.c_uint, ctypes.POINTER(vp)]
    L.pngloss_b200_batch_destroy.argtypes = [vp]
    L.pngloss_b200_batch_destroy.restype = None
    L.pngloss_b200_batch_set_mode.argtypes = [vp, sz, i32, u32]
    L.pngloss_b200_batch_upload.argtypes = [vp, sz, vp, sz]
    L.pngloss_b200_batch_upload_rows.argtypes = [vp, sz, vp]
    L.pngloss_b200_batch_synth.argtypes = [vp, sz, ctypes.c_uint64]
    L.pngloss_b200_batch_run.argtypes = [vp, ctypes.c_uint, ctypes.c_long]
    L.pngloss_b200_batch_download.argtypes = [vp, sz, vp, sz, vp]
    L.pngloss_b200_batch_download_rows.argtypes = [vp, sz, vp, vp]
    L.pngloss_b200_batch_download_input.argtypes = [vp, sz, vp, sz]
    L.pngloss_b200_batch_finish.argtypes = [vp, vp, vp, vp]
    L.pngloss_b200_batch_image_histogram.argtypes = [vp, sz, vp]
    L.pngloss_b200_batch_image_original_histogram.argtypes = [vp, sz, vp]
    L.pngloss_b200_batch_histogram.argtypes = [vp, vp]
    L.pngloss_b200_batch_histogram_device.argtypes = [vp]
    L.pngloss_b200_batch_histogram_device.restype = vp
    L.pngloss_b200_batch_timings.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.pngloss_b200_batch_launch_info.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    L.pngloss_b200_batch_scanlines.argtypes = [vp]
    L.pngloss_b200_batch_scanline_info.argtypes = [vp, sz, ctypes.POINTER(u32), ctypes.POINTER(u32),
                                                   ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_float)]
    L.pngloss_b200_batch_download_scanlines.argtypes = [vp, sz, vp, sz]
    L.pngloss_b200_comm_unique_id.argtypes = [vp]
    L.pngloss_b200_comm_init_rank.argtypes = [vp, i32, i32, vp]
    L.pngloss_b200_comm_init_all.argtypes = [ctypes.POINTER(vp), i32]
    L.pngloss_b200_comm_destroy.argtypes = [vp]
    L.pngloss_b200_comm_destroy.restype = None
    L.pngloss_b200_comm_size.argtypes = [vp]
    L.pngloss_b200_batch_allreduce_histogram.argtypes = [vp]
    L.pngloss_b200_comm_allreduce_u64.argtypes = [vp, vp, sz, i32]
    L.pngloss_b200_ctx_flush_l2.argtypes = [vp]
    L.pngloss_b200_ctx_set_pipeline.argtypes = [vp, i32]
    L.pngloss_b200_ctx_symbol_histogram.argtypes = [vp, vp, i32]
    _lib = L
    return L


def device_count() -> int:
    return load_library().pngloss_b200_device_count()


# ---------------------------------------------------------------------------------------------------
# Mirror of the reference interface (src/pngloss_image.h): same names, argument meaning, in-place.
# ---------------------------------------------------------------------------------------------------
def _row_pointers(a: np.ndarray):
    h = a.shape[0]
    return (ctypes.c_void_p * h)(*[a.ctypes.data + y * a.strides[0] for y in range(h)])


def optimize_with_rows(rgba: np.ndarray, row_filters: Optional[np.ndarray], verbose: bool,
                       quantization_strength: int, bleed_divider: int) -> int:
    """reference src/pngloss_image.c:52 - `rgba` (h, w, 4) uint8 is quantised in place, row_filters
    (h,) uint8 receives the libpng masks (or None)."""
    assert rgba.dtype == np.uint8 and rgba.ndim == 3 and rgba.shape[2] == 4 and rgba.strides[1] == 4
    h, w, _ = rgba.shape
    rf = row_filters.ctypes.data if row_filters is not None else None
    return load_library().optimize_with_rows(_row_pointers(rgba), w, h, rf, verbose,
                                             quantization_strength, bleed_divider)


def optimize_image(packed: np.ndarray, bytes_per_pixel: int, row_filters: Optional[np.ndarray],
                   verbose: bool, quantization_strength: int, bleed_divider: int) -> int:
    """reference src/pngloss_image.c:159 - `packed` (h, w * bytes_per_pixel) uint8, in place."""
    assert packed.dtype == np.uint8 and packed.ndim == 2 and packed.strides[1] == 1
    h = packed.shape[0]
    w = packed.shape[1] // bytes_per_pixel
    rows = (ctypes.c_void_p * h)(*[packed.ctypes.data + y * packed.strides[0] for y in range(h)])
    img = PnglossImage(rows, w, h, bytes_per_pixel)
    rf = row_filters.ctypes.data if row_filters is not None else None
    return load_library().optimize_image(ctypes.byref(img), rf, verbose, quantization_strength,
                                         bleed_divider)


def optimize_with_stride(rgba: np.ndarray, verbose: bool, quantization_strength: int,
                         bleed_divider: int) -> None:
    """reference src/pngloss_image.c:40"""
    h, w, _ = rgba.shape
    load_library().optimize_with_stride(rgba.ctypes.data, w, h, rgba.strides[0], verbose,
                                        quantization_strength, bleed_divider)


def optimizeForAverageFilter(rgba: np.ndarray, quantization_strength: int) -> None:
    """reference src/pngloss_image.c:29 (tight stride, bleed 2)"""
    assert rgba.flags["C_CONTIGUOUS"]
    h, w, _ = rgba.shape
    load_library().optimizeForAverageFilter(rgba.ctypes.data, w, h, quantization_strength)


# ---------------------------------------------------------------------------------------------------
# Batch objects
# ---------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        self.handle = ctypes.c_void_p()
        rc = self.lib.pngloss_b200_ctx_create(ctypes.byref(self.handle), device, stream)
        if rc:
            raise PnglossError(rc, f"cannot create a context on CUDA device {device} "
                               "(no CPU fallback exists)")
        self.device = device

    def _check(self, rc):
        if rc:
            raise PnglossError(rc, self.lib.pngloss_b200_ctx_error(self.handle).decode())

    def set_lanes(self, lanes_per_channel: int):
        self._check(self.lib.pngloss_b200_ctx_set_lanes(self.handle, lanes_per_channel))

    def set_bucket_maxima(self, mode: int):
        """K2's candidate choice: 1 winner table + scan fall-back, 0 scan only, -1 from the strength."""
        self._check(self.lib.pngloss_b200_ctx_set_bucket_maxima(self.handle, mode))

    def set_lean(self, mode: int):
        """1: the lean kernel wherever it applies, -1 (default): only for grids beyond two CTAs per SM, 0: never"""
        self._check(self.lib.pngloss_b200_ctx_set_lean(self.handle, mode))

    def set_solo(self, mode: int):
        """The latency kernel for one-image-per-CTA grids: 1 one chain warp (eight warps per CTA), 2 five chain warps,
        3 one chain warp in a four-warp CTA (up to four CTAs per SM), -1 library's choice, 0 never."""
        self._check(self.lib.pngloss_b200_ctx_set_solo(self.handle, mode))

    # ---- multi-GPU: the library's own NCCL communicator (pl_comm.cuh) --------------------------------------
    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == COMM_ID_BYTES
        buf = ctypes.create_string_buffer(unique_id, COMM_ID_BYTES)
        self._check(self.lib.pngloss_b200_comm_init_rank(self.handle, nranks, rank, buf))

    def comm_size(self) -> int:
        return self.lib.pngloss_b200_comm_size(self.handle)

    def comm_allreduce(self, values, op: str = "sum") -> np.ndarray:
        """Up to 256 host integers reduced over the ranks (blocking); op in sum / max / min."""
        a = np.ascontiguousarray(np.asarray(values, np.uint64).reshape(-1)).copy()
        self._check(self.lib.pngloss_b200_comm_allreduce_u64(self.handle, a.ctypes.data, a.size,
                                                            {"sum": 0, "max": 1, "min": 2}[op]))
        return a

    def barrier(self):
        if self.comm_size() > 1:
            self.comm_allreduce([0])

    def symbol_histogram(self, across_ranks: bool = False) -> np.ndarray:
        """Symbol counts of every image the host-buffer calls of this context finished (NCCL sum over ranks)."""
        out = np.zeros(256, np.uint64)
        self._check(self.lib.pngloss_b200_ctx_symbol_histogram(self.handle, out.ctypes.data, int(across_ranks)))
        return out

    def set_pipeline(self, jobs_in_flight: int):
        """Device batches the job API keeps at once; > 2: every job computes on its own stream (see the header)."""
        self._check(self.lib.pngloss_b200_ctx_set_pipeline(self.handle, jobs_in_flight))

    def flush_l2(self):
        self._check(self.lib.pngloss_b200_ctx_flush_l2(self.handle))

    def timer_start(self):
        self._check(self.lib.pngloss_b200_ctx_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        self._check(self.lib.pngloss_b200_ctx_timer_stop(self.handle, ctypes.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self.lib.pngloss_b200_ctx_sync(self.handle))

    def pinned_empty(self, shape, dtype=np.uint8) -> np.ndarray:
        """numpy array backed by cudaMallocHost memory (kept alive by the array's base)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.lib.pngloss_b200_host_alloc(nbytes)
        if not p:
            raise PnglossError(OUT_OF_MEMORY, f"cudaMallocHost({nbytes})")
        buf = (ctypes.c_uint8 * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        _PINNED[arr.ctypes.data] = (p, self.lib)
        return arr

    def free_pinned(self, arr: np.ndarray):
        p, lib = _PINNED.pop(arr.ctypes.data)
        lib.pngloss_b200_host_free(p)

    @staticmethod
    def _descs(images, row_filters, force_bpp, outputs=None, scanlines=None, no_pixels=False):
        n = len(images)
        descs = (ImageDesc * n)()
        for i, a in enumerate(images):
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4 and a.strides[1] == 4
            rf = row_filters[i] if row_filters is not None else None
            descs[i].pixels = a.ctypes.data
            descs[i].stride = a.strides[0]
            descs[i].width = a.shape[1]
            descs[i].height = a.shape[0]
            descs[i].row_filters = rf.ctypes.data if rf is not None else None
            descs[i].force_bytes_per_pixel = force_bpp
            if outputs is not None:
                o = outputs[i]
                assert o.dtype == np.uint8 and o.shape == a.shape and o.strides[1] == 4
                descs[i].out_pixels = o.ctypes.data
                descs[i].out_stride = o.strides[0]
            if scanlines is not None:
                sc = scanlines[i]
                assert sc.dtype == np.uint8 and sc.size >= a.shape[0] * (1 + 4 * a.shape[1])
                descs[i].scanlines = sc.ctypes.data
                descs[i].flags = 1 if no_pixels else 0
        return descs

    def optimize_batch(self, images: Sequence[np.ndarray],
                       row_filters: Optional[Sequence[Optional[np.ndarray]]], strength: int,
                       bleed: int, force_bpp: int = 0,
                       outputs: Optional[Sequence[np.ndarray]] = None,
                       scanlines: Optional[Sequence[np.ndarray]] = None, no_pixels: bool = False) -> List[dict]:
        """Host-buffer batch: images are quantised in place (or into `outputs`).  row_filters: list of
        (h,) uint8 arrays, entries (or the list) may be None for the reference's row_filters == NULL
        semantics."""
        descs = self._descs(images, row_filters, force_bpp, outputs, scanlines, no_pixels)
        rc = self.lib.pngloss_b200_optimize_batch(self.handle, descs, len(images), strength, bleed)
        self._check(rc)
        return [dict(status=d.status, bytes_per_pixel=d.bytes_per_pixel, retried_rows=d.retried_rows,
                     scan_bytes_per_pixel=d.scan_bytes_per_pixel, scan_row0_filter=d.scan_row0_filter,
                     scan_bytes=d.scan_bytes) for d in descs]

    def submit(self, images, row_filters, strength: int, bleed: int, force_bpp: int = 0, outputs=None) -> "Job":
        """Asynchronous optimize_batch (pngloss_b200_submit): returns a Job; Job.wait() blocks and returns
        the per-image results.  The arrays must stay alive (and untouched) until then."""
        descs = self._descs(images, row_filters, force_bpp, outputs)
        handle = ctypes.c_void_p()
        self._check(self.lib.pngloss_b200_submit(self.handle, descs, len(images), strength, bleed,
                                                 ctypes.byref(handle)))
        return Job(self, handle, descs, (images, row_filters, outputs))

    def close(self):
        if self.handle:
            self.lib.pngloss_b200_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Job:
    """One batch in flight (pngloss_b200_submit / pngloss_b200_wait)."""

    def __init__(self, ctx, handle, descs, keepalive):
        self.ctx, self.handle, self.descs, self.keepalive = ctx, handle, descs, keepalive

    def wait(self) -> List[dict]:
        rc = self.ctx.lib.pngloss_b200_wait(self.handle)
        self.handle = None
        if rc and rc != NO_ACCEPTABLE_ROW:
            self.ctx._check(rc)
        return [dict(status=d.status, bytes_per_pixel=d.bytes_per_pixel, retried_rows=d.retried_rows)
                for d in self.descs]


_PINNED = {}


class Batch:
    """Device-resident batch (pngloss_b200_batch_*)."""

    def __init__(self, ctx: Context, widths: Sequence[int], heights: Sequence[int], in_place: bool = False):
        """in_place: the quantised rows overwrite the uploaded ones on the device (half the memory; every
        run needs a fresh upload and download_input returns the result afterwards)."""
        self.ctx = ctx
        self.lib = ctx.lib
        self.n = len(widths)
        self.widths = np.asarray(widths, np.uint32)
        self.heights = np.asarray(heights, np.uint32)
        self.handle = ctypes.c_void_p()
        ctx._check(self.lib.pngloss_b200_batch_create_ex(ctx.handle, self.n, self.widths.ctypes.data,
                                                         self.heights.ctypes.data, 1 if in_place else 0,
                                                         ctypes.byref(self.handle)))

    def set_mode(self, i, adaptive_all=False, force_bpp=0):
        self.ctx._check(self.lib.pngloss_b200_batch_set_mode(self.handle, i, int(adaptive_all), force_bpp))

    def upload(self, i, rgba: np.ndarray):
        assert rgba.shape == (self.heights[i], self.widths[i], 4) and rgba.strides[1] == 4
        self.ctx._check(self.lib.pngloss_b200_batch_upload(self.handle, i, rgba.ctypes.data,
                                                           rgba.strides[0]))

    def synth(self, i, seed):
        self.ctx._check(self.lib.pngloss_b200_batch_synth(self.handle, i, seed))

    def run(self, strength, bleed):
        self.ctx._check(self.lib.pngloss_b200_batch_run(self.handle, strength, bleed))

    def download(self, i, rgba: Optional[np.ndarray] = None, row_filters: Optional[np.ndarray] = None):
        self.ctx._check(self.lib.pngloss_b200_batch_download(
            self.handle, i, rgba.ctypes.data if rgba is not None else None,
            rgba.strides[0] if rgba is not None else 0,
            row_filters.ctypes.data if row_filters is not None else None))

    def download_input(self, i, rgba: np.ndarray):
        self.ctx._check(self.lib.pngloss_b200_batch_download_input(self.handle, i, rgba.ctypes.data,
                                                                   rgba.strides[0]))

    def finish(self):
        st = np.zeros(self.n, np.int32)
        bpp = np.zeros(self.n, np.uint32)
        rt = np.zeros(self.n, np.uint32)
        rc = self.lib.pngloss_b200_batch_finish(self.handle, st.ctypes.data, bpp.ctypes.data,
                                                rt.ctypes.data)
        if rc and rc != NO_ACCEPTABLE_ROW:
            self.ctx._check(rc)
        return st, bpp, rt

    def image_histogram(self, i) -> np.ndarray:
        out = np.zeros(256, np.uint32)
        self.ctx._check(self.lib.pngloss_b200_batch_image_histogram(self.handle, i, out.ctypes.data))
        return out

    def original_histogram(self, i) -> np.ndarray:
        """K1's output for image i: counts of (byte - predictor) per filter and RGBA channel, (5, 4, 256)."""
        out = np.zeros((5, 4, 256), np.uint32)
        self.ctx._check(self.lib.pngloss_b200_batch_image_original_histogram(self.handle, i, out.ctypes.data))
        return out

    def histogram(self) -> np.ndarray:
        out = np.zeros(256, np.uint64)
        self.ctx._check(self.lib.pngloss_b200_batch_histogram(self.handle, out.ctypes.data))
        return out

    def allreduce_histogram(self):
        """NCCL sum of the batch histogram over all ranks, in place on the device (asynchronous)."""
        self.ctx._check(self.lib.pngloss_b200_batch_allreduce_histogram(self.handle))

    def histogram_device_ptr(self) -> int:
        return self.lib.pngloss_b200_batch_histogram_device(self.handle)

    def timings(self):
        ms = (ctypes.c_float * 4)()
        self.ctx._check(self.lib.pngloss_b200_batch_timings(self.handle, ms))
        return dict(k1_hist_ms=ms[0], k2_quantize_ms=ms[1], k3_batch_hist_ms=ms[2], run_ms=ms[3])

    def launch_info(self):
        info = (ctypes.c_uint32 * 4)()
        self.ctx._check(self.lib.pngloss_b200_batch_launch_info(self.handle, info))
        return dict(k2_ctas=info[0], images_per_cta=info[1], k2_smem_bytes=info[2], launches=info[3] & 0xff,
                    bucket_maxima=bool(info[3] & 0x100), lean=bool(info[3] & 0x200), solo=bool(info[3] & 0x400))

    def scanlines(self):
        """K4: filtered PNG scanlines of the results, on the device (asynchronous)."""
        self.ctx._check(self.lib.pngloss_b200_batch_scanlines(self.handle))

    def scanline_info(self, i):
        bpp, f0, nbytes, ms = u32t(), u32t(), ctypes.c_size_t(), (ctypes.c_float * 2)()
        self.ctx._check(self.lib.pngloss_b200_batch_scanline_info(
            self.handle, i, ctypes.byref(bpp), ctypes.byref(f0), ctypes.byref(nbytes), ms))
        return dict(bytes_per_pixel=bpp.value, row0_filter=f0.value, bytes=nbytes.value,
                    k4_scan_ms=ms[0], k4_filter_ms=ms[1])

    def download_scanlines(self, i) -> np.ndarray:
        """(height, 1 + width * bytes_per_pixel) uint8: filter-type byte, then the filtered row."""
        info = self.scanline_info(i)
        buf = np.empty(info["bytes"], np.uint8)
        self.ctx._check(self.lib.pngloss_b200_batch_download_scanlines(self.handle, i, buf.ctypes.data, buf.size))
        self.ctx.sync()
        return buf.reshape(int(self.heights[i]), -1)

    def close(self):
        if self.handle:
            self.lib.pngloss_b200_batch_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
