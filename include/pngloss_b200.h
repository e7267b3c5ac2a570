/* pngloss_b200 - C ABI of the B200-native quantise + filter-search path of pngloss.
 *
 * The library (pngloss_b200/libpngloss_b200.so) is plain C at the boundary: pointers and sizes only,
 * no CUDA or torch types.  It contains no CPU implementation of the path: every entry point runs the
 * sm_100a kernels and fails with PNGLOSS_B200_DEVICE_ERROR when no usable device is present.
 *
 * Part 1 is the drop-in: the four public symbols of the reference's src/pngloss_image.h with the
 * same names, argument meaning, in-place behaviour and error codes, so that the reference's
 * src/pngloss.c:266 call site links against this library unchanged.
 * Part 2 adds what the reference does not have: a batch entry (the reference loops over files one
 * call at a time, src/pngloss.c:173-205) and a device-resident batch object for pipelines and
 * benchmarking.
 */
#ifndef PNGLOSS_B200_H
#define PNGLOSS_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: the subset of the reference's pngloss_error (src/rwpng.h:23-38) this path can
 * return, plus one value of our own for device failures. */
#define PNGLOSS_B200_SUCCESS 0
#define PNGLOSS_B200_INVALID_ARGUMENT 4   /* reference INVALID_ARGUMENT */
#define PNGLOSS_B200_OUT_OF_MEMORY 17     /* reference OUT_OF_MEMORY_ERROR (host or device) */
#define PNGLOSS_B200_DEVICE_ERROR 40      /* no CUDA device / kernel or copy failed (new) */
#define PNGLOSS_B200_NO_ACCEPTABLE_ROW 41 /* the reference abort()s here, src/pngloss_image.c:268 */

/* ------------------------------------------------------------------------------------------------
 * Part 1 - drop-in for reference src/pngloss_image.h.
 * When the reference's own header is included first these prototypes are skipped; the symbols the
 * library exports are ABI-identical (pngloss_error is a 4-byte enum, uint_fast8_t is unsigned char,
 * int_fast16_t is long on x86-64 SysV).
 * ---------------------------------------------------------------------------------------------- */
#ifndef PNGLOSS_IMAGE_H
/* replaces reference src/pngloss_image.c:52 (declared src/pngloss_image.h:21-25).
 * rows[y] -> width*4 RGBA8 bytes, quantised in place; row_filters[height] receives libpng filter
 * masks 0x08/0x10/0x20/0x40/0x80, or may be NULL, in which case every row is checked against
 * libpng's own filter heuristic.  Returns 0 or 17; other failures abort() like the reference. */
int optimize_with_rows(unsigned char **rows, uint32_t width, uint32_t height,
                       unsigned char *row_filters, bool verbose,
                       uint_fast8_t quantization_strength, int_fast16_t bleed_divider);
/* replaces reference src/pngloss_image.c:40 (row_filters = NULL) */
void optimize_with_stride(unsigned char *pixels, uint32_t width, uint32_t height, uint32_t stride,
                          bool verbose, uint_fast8_t quantization_strength,
                          int_fast16_t bleed_divider);
/* replaces reference src/pngloss_image.c:29 (bleed 2, tight stride, row_filters = NULL) */
void optimizeForAverageFilter(unsigned char pixels[], int width, int height, int quantization);
/* replaces reference src/pngloss_image.c:159 (declared src/pngloss_image.h:26-29): the inner entry on
 * a packed image with an explicit bytes_per_pixel of 1 (gray), 2 (gray+alpha), 3 (rgb) or 4 (rgba). */
typedef struct {
    unsigned char **rows;
    uint32_t width, height;
    uint_fast8_t bytes_per_pixel;
} pngloss_image;
int optimize_image(pngloss_image *image, unsigned char *row_filters, bool verbose,
                   uint_fast8_t quantization_strength, int_fast16_t bleed_divider);
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 2 - batch and device-resident entry points (new).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pngloss_b200_ctx pngloss_b200_ctx;
typedef struct pngloss_b200_batch pngloss_b200_batch;

int pngloss_b200_device_count(void);
/* One context per GPU and host thread.  cuda_stream: a cudaStream_t to enqueue on, or NULL to let
 * the context create its own non-blocking stream. */
int pngloss_b200_ctx_create(pngloss_b200_ctx **out, int device, void *cuda_stream);
void pngloss_b200_ctx_destroy(pngloss_b200_ctx *ctx);
/* Human readable description of the last failure on this context ("" if none). */
const char *pngloss_b200_ctx_error(const pngloss_b200_ctx *ctx);
/* Lane mapping of the quantise kernel: lanes per colour channel = 8, 4, 2 or 1, i.e. 1, 2, 4 or 8
 * images per CTA; 0 = choose from the batch size.  A tuning knob, results never depend on it. */
int pngloss_b200_ctx_set_lanes(pngloss_b200_ctx *ctx, int lanes_per_channel);
/* Candidate choice of the quantise kernel: 1 = per-band winner table ("bucket maxima") with the scan as
 * fall-back, 0 = scan every band, -1 = choose from the strength (default).  A tuning knob, results
 * never depend on it. */
int pngloss_b200_ctx_set_bucket_maxima(pngloss_b200_ctx *ctx, int mode);
/* Kernel variant for batches that run one lane per channel with the winner table: 1 = the lean kernel (bulk-copy
 * tile ring, three CTAs per SM) wherever its conditions hold (every width a multiple of 4), -1 (default) = the
 * same, but only for grids of more than two CTAs per SM (where it is the faster one), 0 = always the generic
 * kernel.  A tuning knob, results never depend on it. */
int pngloss_b200_ctx_set_lean(pngloss_b200_ctx *ctx, int mode);
/* Kernel variant for batches small enough that every image gets a CTA (and an SM) of its own - a single image above
 * all, the reference's one-file-per-call use (src/pngloss.c:173-205,266): the latency kernel (warp-specialised: chain /
 * producer / post warps) for strengths up to 126.  -1 (default) = for batches of at
 * most four images per SM, unless a lane mapping was set explicitly; 1 = one chain warp for the five filter candidates
 * in a CTA of eight warps, 2 = one chain warp per candidate, 3 = one chain warp in a CTA of four warps (1 .. 3:
 * whenever the lane mapping is 0 or 8); 0 = never.  A tuning knob, results never depend on it. */
int pngloss_b200_ctx_set_solo(pngloss_b200_ctx *ctx, int mode);
/* CUDA-event stopwatch on the context's stream (what bench.py times with). */
int pngloss_b200_ctx_timer_start(pngloss_b200_ctx *ctx);
int pngloss_b200_ctx_timer_stop(pngloss_b200_ctx *ctx, float *milliseconds);
int pngloss_b200_ctx_sync(pngloss_b200_ctx *ctx);
/* Pinned host memory for fast uploads / downloads. */
void *pngloss_b200_host_alloc(size_t bytes);
void pngloss_b200_host_free(void *p);

typedef struct {
    unsigned char *pixels;      /* RGBA8, quantised in place */
    size_t stride;              /* bytes between rows, >= width * 4 */
    uint32_t width, height;
    unsigned char *row_filters; /* height bytes, or NULL = all rows adaptive (see above) */
    uint32_t force_bytes_per_pixel; /* 0 = detect gray / opaque like optimize_with_rows;
                                       1..4 = explicit mode like the reference's optimize_image
                                       (src/pngloss_image.c:159) on RGBA-laid-out input */
    uint32_t bytes_per_pixel;   /* out: colour mode used */
    uint32_t retried_rows;      /* out: strength decrements that were needed (normally 0) */
    int status;                 /* out: PNGLOSS_B200_* for this image */
    unsigned char *out_pixels;  /* NULL: quantise `pixels` in place; else the result goes here and */
    size_t out_stride;          /* `pixels` is left alone */
    /* Optional: the result as filtered PNG scanlines, ready for deflate (see pngloss_b200_batch_scanlines).
     * scanlines: NULL, or room for height * (1 + 4 * width) bytes. */
    unsigned char *scanlines;
    uint32_t flags;             /* PNGLOSS_B200_IMAGE_* */
    uint32_t scan_bytes_per_pixel; /* out: 1 gray, 2 gray+alpha, 3 rgb, 4 rgba - what the output pixels allow */
    uint32_t scan_row0_filter;  /* out: PNG filter type libpng's heuristic picks for row 0 */
    size_t scan_bytes;          /* out: height * (1 + width * scan_bytes_per_pixel) */
} pngloss_b200_image;
#define PNGLOSS_B200_IMAGE_NO_PIXELS 1u   /* with `scanlines`: do not copy the quantised pixels back */

/* Upload, run, download a batch of independent images.  Blocking.  Returns the first non-zero
 * per-image status, or 0.  A batch that does not fit the device runs as consecutive groups whose
 * copies overlap each other's kernels.  The pixels of an image that ends with
 * PNGLOSS_B200_NO_ACCEPTABLE_ROW are unspecified (the reference abort()s at that point). */
int pngloss_b200_optimize_batch(pngloss_b200_ctx *ctx, pngloss_b200_image *images, size_t n,
                                unsigned strength, long bleed);

/* The same, asynchronous: submit enqueues the uploads, kernels and downloads of one batch and returns;
 * wait blocks until its results are in the host buffers, fills the out fields of `images` and frees the
 * job.  Several jobs may be in flight: their kernels run one after the other, while the PCIe copies of
 * one job overlap the kernels of another (use pinned host memory, pngloss_b200_host_alloc, or the copies
 * are not asynchronous).  `images` and the buffers it points to must stay valid until wait returns.
 * When the device cannot hold another batch, submit first waits for the oldest job in flight. */
typedef struct pngloss_b200_job pngloss_b200_job;
/* How many jobs the context keeps on the device at once (default 2: the copies of one job overlap the kernels of
 * the other, kernels run back to back - right when every job fills the GPU by itself).  With more than 2 every
 * job also computes on a stream of its own, so that many small jobs share the SMs: as one job's CTAs finish,
 * the next job's CTAs take their places, and uploads, kernels and downloads of different jobs overlap
 * continuously - the streaming set-up for inputs that together exceed the device memory.  Call before the first
 * submit. */
int pngloss_b200_ctx_set_pipeline(pngloss_b200_ctx *ctx, int jobs_in_flight);
int pngloss_b200_submit(pngloss_b200_ctx *ctx, pngloss_b200_image *images, size_t n, unsigned strength,
                        long bleed, pngloss_b200_job **job);
int pngloss_b200_wait(pngloss_b200_job *job);

/* Device-resident batch: allocate once, then upload / run / download as often as wanted. */
int pngloss_b200_batch_create(pngloss_b200_ctx *ctx, size_t n, const uint32_t *widths,
                              const uint32_t *heights, pngloss_b200_batch **out);
/* flags: PNGLOSS_B200_BATCH_IN_PLACE - the quantised rows overwrite the uploaded ones on the device
 * (half the memory; a second run needs a fresh upload, and download_input then returns the result). */
#define PNGLOSS_B200_BATCH_IN_PLACE 1u
int pngloss_b200_batch_create_ex(pngloss_b200_ctx *ctx, size_t n, const uint32_t *widths,
                                 const uint32_t *heights, unsigned flags, pngloss_b200_batch **out);
void pngloss_b200_batch_destroy(pngloss_b200_batch *b);
/* adaptive_all / force_bytes_per_pixel per image; defaults 0 / 0 */
int pngloss_b200_batch_set_mode(pngloss_b200_batch *b, size_t i, int adaptive_all,
                                uint32_t force_bytes_per_pixel);
int pngloss_b200_batch_upload(pngloss_b200_batch *b, size_t i, const unsigned char *pixels,
                              size_t stride);                     /* async on the ctx stream */
int pngloss_b200_batch_upload_rows(pngloss_b200_batch *b, size_t i, unsigned char *const *rows);
int pngloss_b200_batch_synth(pngloss_b200_batch *b, size_t i, uint64_t seed); /* SURVEY 8d generator */
int pngloss_b200_batch_run(pngloss_b200_batch *b, unsigned strength, long bleed); /* async */
int pngloss_b200_batch_download(pngloss_b200_batch *b, size_t i, unsigned char *pixels,
                                size_t stride, unsigned char *row_filters);  /* async */
int pngloss_b200_batch_download_rows(pngloss_b200_batch *b, size_t i, unsigned char *const *rows,
                                     unsigned char *row_filters);
int pngloss_b200_batch_download_input(pngloss_b200_batch *b, size_t i, unsigned char *pixels,
                                      size_t stride);
/* Waits for the stream; fills per-image status / mode / retries (each may be NULL). */
int pngloss_b200_batch_finish(pngloss_b200_batch *b, int *status, uint32_t *bytes_per_pixel,
                              uint32_t *retried_rows);
/* Per-image final symbol histogram [256] and the batch sum [256] (valid after finish). */
int pngloss_b200_batch_image_histogram(pngloss_b200_batch *b, size_t i, uint32_t *out256);
int pngloss_b200_batch_histogram(pngloss_b200_batch *b, uint64_t *out256);
/* The histogram kernel's output for one image (valid after finish): counts of (byte - predictor) over the
 * original image per filter and RGBA channel, [5][4][256]; summed over the channels of the image's colour mode
 * it is the reference's original_frequency table (src/optimize_state.c:66-83). */
int pngloss_b200_batch_image_original_histogram(pngloss_b200_batch *b, size_t i, uint32_t *out5x4x256);
/* Device address of the batch sum (uint64[256]) for the caller's NCCL all-reduce. */
void *pngloss_b200_batch_histogram_device(pngloss_b200_batch *b);
/* Durations of the last run in milliseconds: [0] histogram kernel, [1] quantise kernel,
 * [2] batch-histogram kernel, [3] whole run.  Valid after finish. */
int pngloss_b200_batch_timings(pngloss_b200_batch *b, float ms[4]);
/* Launch geometry of the last run: [0] quantise CTAs, [1] images per CTA, [2] dynamic smem bytes,
 * [3] kernels launched (bits 0-7), bit 8: the quantise kernel used the bucket-maxima table, bit 9: it was the
 * lean kernel. */
int pngloss_b200_batch_launch_info(pngloss_b200_batch *b, uint32_t info[4]);

/* Filtered PNG scanlines of the batch's results, ready for deflate (what the reference leaves to libpng
 * after the hot path, src/rwpng.c:488-495,557-613): for every image the colour type its output pixels allow
 * (1 gray, 2 gray+alpha, 3 rgb, 4 rgba bytes per pixel), and height rows of one filter-type byte plus
 * width * bytes_per_pixel filtered bytes - row 0 filtered by libpng's minimum-sum-of-absolute-differences
 * heuristic, every other row by the filter the search chose.  Asynchronous; call after pngloss_b200_batch_run.
 * The first call allocates a second device buffer of height * (1 + 4 * width) bytes per image. */
int pngloss_b200_batch_scanlines(pngloss_b200_batch *b);
/* Waits for the stream.  Each out pointer may be NULL; milliseconds = device time of the two scanline
 * kernels of the last call ([0] colour-type scan, [1] filtering). */
int pngloss_b200_batch_scanline_info(pngloss_b200_batch *b, size_t i, uint32_t *bytes_per_pixel,
                                     uint32_t *row0_filter, size_t *bytes, float milliseconds[2]);
int pngloss_b200_batch_download_scanlines(pngloss_b200_batch *b, size_t i, unsigned char *dst,
                                          size_t capacity);   /* async copy of `bytes` bytes */

/* ------------------------------------------------------------------------------------------------
 * Part 3 - multi-GPU (new).  Images are the shard unit and need no exchange; the one collective of the path
 * is the sum of the batch symbol histograms over all GPUs (256 x u64) - the batch-level form of the
 * reference's "used N unique symbols" report (src/pngloss_image.c:315-325).  The library issues it itself as
 * an NCCL all-reduce on the context's stream; NCCL is loaded at run time (libnccl.so.2, or the path in
 * PNGLOSS_B200_NCCL_LIB), so single-GPU use does not need it.
 * ---------------------------------------------------------------------------------------------- */
#define PNGLOSS_B200_COMM_ID_BYTES 128
/* One process per GPU: rank 0 creates an id, hands it to the other ranks by whatever means the launcher has,
 * and every rank joins with its context. */
int pngloss_b200_comm_unique_id(unsigned char id[PNGLOSS_B200_COMM_ID_BYTES]);
int pngloss_b200_comm_init_rank(pngloss_b200_ctx *ctx, int nranks, int rank,
                                const unsigned char id[PNGLOSS_B200_COMM_ID_BYTES]);
/* One process, one context (and host thread) per GPU. */
int pngloss_b200_comm_init_all(pngloss_b200_ctx **ctxs, int n);
void pngloss_b200_comm_destroy(pngloss_b200_ctx *ctx);
int pngloss_b200_comm_size(const pngloss_b200_ctx *ctx);   /* 0 = no communicator */
/* Sum of the batch histogram over all ranks, in place on the device, asynchronous on the context's stream
 * (call after pngloss_b200_batch_run; read with pngloss_b200_batch_histogram after _finish).  Every rank must
 * call it once per run. */
int pngloss_b200_batch_allreduce_histogram(pngloss_b200_batch *b);
/* Up to 256 host values reduced over the ranks, blocking (a barrier when the result is ignored).
 * op: 0 sum, 1 max, 2 min. */
int pngloss_b200_comm_allreduce_u64(pngloss_b200_ctx *ctx, uint64_t *values, size_t n, int op);

/* Symbol counts (256 x u64) of every image this context's host-buffer calls (optimize_batch, submit / wait) have
 * finished so far.  across_ranks != 0 and a communicator on the context: summed over all ranks by the NCCL
 * all-reduce (blocking; every rank must call). */
int pngloss_b200_ctx_symbol_histogram(pngloss_b200_ctx *ctx, uint64_t out256[256], int across_ranks);

/* Benchmark helper: overwrite a buffer larger than the L2 cache on the context's stream. */
int pngloss_b200_ctx_flush_l2(pngloss_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PNGLOSS_B200_H */
