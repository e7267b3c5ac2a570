/* PNG reader / writer for the pngloss_b200 host program.
 *
 * Replaces the reference's src/rwpng.c (a libpng wrapper) for SURVEY 8(f) row 2: this image has zlib but
 * no libpng headers, so the container format is handled here directly on top of zlib.  Names and the
 * png24_image fields follow the reference's src/rwpng.h:56-88 so that the driver reads like the
 * reference's; the implementation is written from the PNG specification, not from libpng.
 *
 * What is kept from the reference's behaviour (src/rwpng.c):
 *   read  (:179-400)  any colour type / bit depth / interlace is expanded to 8-bit RGBA (16-bit samples
 *                     keep their high byte, tRNS becomes alpha, gray becomes R=G=B); unknown ancillary
 *                     chunks and pHYs/iTXt/tEXt/zTXt are captured for pass-through unless `strip`;
 *                     sRGB input is remembered so that the output is tagged again.
 *   write (:515-637)  colour type is auto-detected from the pixels (gray / gray+alpha / rgb / rgba),
 *                     8 bits per sample; row 0 is filtered by the min-sum-of-absolute-differences
 *                     heuristic, rows >= 1 by the caller's row_filters[] (libpng masks 0x08..0x80), or
 *                     by the heuristic when row_filters is NULL; zlib level 9, memLevel 9, Z_FILTERED;
 *                     an optional size cap returns TOO_LARGE_FILE.
 */
#ifndef PL_PNG_H
#define PL_PNG_H

#include <setjmp.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

/* Types below are layout-identical to the reference's src/rwpng.h (same members, order and enum values), so
 * that the reference's own src/pngloss.c, compiled against ITS rwpng.h, links against this reader / writer
 * unchanged (oracle/Makefile target refcli; tests/test_gpu_cli.py runs that binary). */

/* reference src/rwpng.h:23-38 */
typedef enum {
    SUCCESS = 0,
    MISSING_ARGUMENT = 1,
    READ_ERROR = 2,
    INVALID_ARGUMENT = 4,
    NOT_OVERWRITING_ERROR = 15,
    CANT_WRITE_ERROR = 16,
    OUT_OF_MEMORY_ERROR = 17,
    WRONG_ARCHITECTURE = 18,
    PNG_OUT_OF_MEMORY_ERROR = 24,
    LIBPNG_FATAL_ERROR = 25,
    WRONG_INPUT_COLOR_TYPE = 26,
    LIBPNG_INIT_ERROR = 35,
    PNGLOSS_DEVICE_ERROR = 40,   /* new: libpngloss_b200 found no usable GPU */
    TOO_LARGE_FILE = 98,
    TOO_LOW_QUALITY = 99,
} pngloss_error;

/* where a passed-through chunk sat in the input */
enum { RWPNG_AFTER_IHDR = 0x01, RWPNG_AFTER_PLTE = 0x02, RWPNG_AFTER_IDAT = 0x08 };

struct rwpng_chunk {
    struct rwpng_chunk *next;
    unsigned char *data;
    size_t size;
    unsigned char name[5];
    unsigned char location;
};

/* reference src/rwpng.h:52-60; this reader only ever reports NONE, SRGB and GAMA_ONLY (no LCMS / Cocoa) */
typedef enum {
    RWPNG_NONE,
    RWPNG_SRGB,           /* sRGB chunk present: written back as gAMA + sRGB */
    RWPNG_ICCP,
    RWPNG_ICCP_WARN_GRAY,
    RWPNG_GAMA_CHRM,
    RWPNG_GAMA_ONLY,      /* gAMA only (or nothing): no colour chunk is written */
    RWPNG_COCOA,
} rwpng_color_transform;

/* reference src/rwpng.h:62-75 */
typedef struct {
    jmp_buf jmpbuf;       /* libpng's error exit in the reference; unused here, kept for the layout */
    uint32_t width;
    uint32_t height;
    size_t file_size;
    size_t maximum_file_size;
    size_t metadata_size;
    double gamma;
    unsigned char **row_pointers;
    unsigned char *rgba_data;
    struct rwpng_chunk *chunks;
    rwpng_color_transform input_color;
    rwpng_color_transform output_color;
} png24_image;

void rwpng_version_info(FILE *fp);
pngloss_error rwpng_read_image24(FILE *infile, png24_image *out, bool strip, bool verbose);
pngloss_error rwpng_write_image24(FILE *outfile, png24_image *image, unsigned char *row_filters);
/* The same file from rows that are already narrowed and filtered (pngloss_b200_image.scanlines: height rows of
 * one filter-type byte + width * bytes_per_pixel bytes); image supplies size, colour tags and chunks. */
pngloss_error rwpng_write_scanlines(FILE *outfile, png24_image *image, unsigned bytes_per_pixel,
                                    const unsigned char *scanlines);
void rwpng_free_image24(png24_image *image);

/* The filter (0..4) libpng's default heuristic picks for a row of `rowbytes` bytes with `bpp` bytes per
 * pixel; `prev` may be NULL.  First minimum in the order none, sub, up, average, paeth. */
int rwpng_heuristic_filter(const unsigned char *prev, const unsigned char *row, size_t rowbytes, unsigned bpp);

#endif
