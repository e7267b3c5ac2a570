// Structures shared by the host shim (pl_api.cu) and the kernels (pl_kernels.cuh).
#pragma once
#include <stdint.h>

#define PL_FILTERS 5          // none, sub, up, average, paeth  (reference src/optimize_state.h:18-25)
#define PL_ERR_PAD 8          // error rows hold width + PL_ERR_PAD cells (reference uses width + 5)
#define PL_K2_WARPS PL_FILTERS
#define PL_K2_THREADS (32 * PL_K2_WARPS)

// image status written by the quantise kernel
#define PL_ST_OK 0
#define PL_ST_NO_ROW 1        // no acceptable row even at strength 0 (reference aborts, pngloss_image.c:268)

// One image of a batch, as the device sees it.  Pixels are tight RGBA8 rows (width * 4 bytes) for
// every colour mode; the narrowed 1/2/3 byte-per-pixel modes of the reference are run on the same
// layout through a channel mask (DESIGN.md "Data layout").
struct PlImageDev {
    const uchar4 *in;        // original image
    uchar4 *out;             // quantised image, row y written once the winner of row y is known; may be
                             // the same buffer as `in` (in-place batch): the kernel never reads an
                             // original row again after its winner is committed
    uchar4 *oprev;           // scratch [width]: the original row above the row in flight
    unsigned char *filters;  // height libpng filter masks (0x08..0x80)
    uint32_t *chan_hist;     // [5][4][256] per-channel original-image histograms (K1 output)
    uint32_t *flags;         // [0] != 0: some pixel is not gray; [1] != 0: some pixel is not opaque
    uint32_t *final_hist;    // [256] symbol histogram after the last row
    short4 *err;             // scratch [2 parity][5 filters][2 rows][width + PL_ERR_PAD]
    uchar4 *cand;            // scratch [5 filters][width] candidate rows of the current row
    uint32_t *status;        // [0] PL_ST_*, [1] bytes-per-pixel mode used, [2] rows that needed a retry
    uint32_t width, height;
    uint32_t adaptive_all;   // 1: caller passed row_filters == NULL, every row is "adaptive"
                             //    (reference src/pngloss_image.c:210)
    uint32_t force_mode;     // 0: detect gray/opaque like optimize_with_rows; 1..4: explicit
                             //    bytes-per-pixel like the reference's optimize_image
};

// active RGBA channels of a colour mode: 1 gray (G), 2 gray+alpha (G,A), 3 rgb, 4 rgba
#define PL_MODE_MASK(mode) ((mode) == 1 ? 0x2 : (mode) == 2 ? 0xA : (mode) == 3 ? 0x7 : 0xF)

// One image as the scanline kernels (K4) see it.
struct PlScanDev {
    const uchar4 *px;              // quantised image (K2's output), tight RGBA8 rows
    const unsigned char *filters;  // K2's row filters (libpng masks)
    unsigned char *scan;           // out: height x (1 + width * bytes_per_pixel) bytes, filter byte first
    uint32_t *oflags;              // [0] != 0: some output pixel is not gray; [1] != 0: not opaque;
                                   // [2] out: filter type used for row 0
    uint32_t width, height;
};
