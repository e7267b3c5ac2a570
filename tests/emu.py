"""ctypes access to tests/simt_emu/libpl_emu.so: the product kernels executed on CPU fibers."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "simt_emu")


class Emu:
    def __init__(self):
        subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
        self.lib = ctypes.CDLL(os.path.join(EMU_DIR, "libpl_emu.so"))
        self.lib.emu_optimize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                                          ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]
        self.lib.emu_optimize.restype = ctypes.c_int
        self.lib.emu_synth.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint64]
        self.lib.emu_synth.restype = None

    def optimize(self, imgs, strength, bleed, adaptive_all, lpc):
        """imgs: list of equally sized (h, w, 4) uint8 arrays -> dict of outputs.
        lpc: lanes per channel (8, 4, 2, 1), + 16 selects the bucket-maxima variant of K2, + 32 an
        in-place batch (output buffer = input buffer)."""
        n = len(imgs)
        h, w, _ = imgs[0].shape
        buf = np.ascontiguousarray(np.stack(imgs)).copy()
        filt = np.zeros((n, h), np.uint8)
        final = np.zeros((n, 256), np.uint32)
        status = np.zeros((n, 3), np.uint32)
        batch = np.zeros(256, np.uint64)
        chan = np.zeros((n, 5, 4, 256), np.uint32)
        rc = self.lib.emu_optimize(buf.ctypes.data, n, w, h, filt.ctypes.data, strength, bleed,
                                   int(adaptive_all), lpc, final.ctypes.data, status.ctypes.data,
                                   batch.ctypes.data, chan.ctypes.data)
        assert rc == 0
        return dict(pixels=buf, filters=filt, final_hist=final, status=status, batch_hist=batch,
                    chan_hist=chan)

    def scanlines(self, imgs, filters):
        """K4 on quantised images (list of equally sized (h, w, 4) arrays) and their row filters (n, h):
        -> list of (bytes_per_pixel, row0_filter, (h, 1 + w * bpp) uint8 array)."""
        n = len(imgs)
        h, w, _ = imgs[0].shape
        buf = np.ascontiguousarray(np.stack(imgs))
        filt = np.ascontiguousarray(np.asarray(filters, np.uint8).reshape(n, h))
        room = h * (1 + 4 * w)
        scan = np.full((n, room), 0x5A, np.uint8)
        oflags = np.zeros((n, 4), np.uint32)
        self.lib.emu_scanlines.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        assert self.lib.emu_scanlines(buf.ctypes.data, n, w, h, filt.ctypes.data, scan.ctypes.data,
                                      oflags.ctypes.data) == 0
        out = []
        for i in range(n):
            bpp = (4 if oflags[i, 1] else 3) if oflags[i, 0] else (2 if oflags[i, 1] else 1)
            assert (scan[i, h * (1 + w * bpp):] == 0x5A).all(), "K4 wrote past its rows"
            out.append((bpp, int(oflags[i, 2]), scan[i, :h * (1 + w * bpp)].reshape(h, 1 + w * bpp).copy()))
        return out

    def counters(self):
        """Sierra taps from the table / computed, channel fix-up replays / skips executed so far (per lane)."""
        out = (ctypes.c_ulonglong * 12)()
        self.lib.emu_counters(out)
        return dict(taps_table=out[0], taps_computed=out[1], fixup_replay=out[2], fixup_skipped=out[3],
                    bm_lookup=out[4], bm_scan=out[5], bm_fast_update=out[6], bm_general_update=out[7],
                    solo_fast=out[8], solo_general=out[9])

    def synth(self, w, h, seed):
        a = np.zeros((h, w, 4), np.uint8)
        self.lib.emu_synth(a.ctypes.data, w, h, seed)
        return a
