/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 *
 * CPU restatement ("oracle port") of the pngloss quantise + filter-search hot
 * path, written from the behaviour of the reference sources:
 *   src/pngloss_image.c   (optimize_with_rows :52-156, optimize_image :159-333)
 *   src/optimize_state.c  (init :28-86, run :114-290, row :292-361,
 *                          diffusion :390-467, libpng heuristic :492-562,
 *                          ulog2 :564-572, predictors :575-613)
 *   src/color_delta.c     (:4-66)
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file byte-for-byte
 * against oracle/_ref/libpngloss_ref.so (the unmodified reference sources
 * compiled where they lie, see oracle/Makefile) and against the golden hashes
 * of SURVEY.md section 8c that were produced by the reference itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may link or load this.  The product (libpngloss_b200.so) never does.
 */
#ifndef PNGLOSS_ORACLE_H
#define PNGLOSS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_OK 0
#define ORACLE_OUT_OF_MEMORY 17          /* reference src/rwpng.h:30 */
#define ORACLE_NO_ACCEPTABLE_ROW (-1)    /* reference abort()s, src/pngloss_image.c:268-271 */

/* Optional per-run trace used by tests to localise a first divergence. */
typedef struct {
    uint64_t *row_costs;      /* [height][5] cost of every candidate of the accepted pass, or NULL */
    uint8_t *row_strength;    /* [height] strength at which the row was accepted, or NULL */
    uint32_t final_frequency[256]; /* symbol histogram after the last row */
    uint32_t original_frequency[5][256];
} oracle_trace;

/* Same contract as the reference's optimize_with_rows (src/pngloss_image.c:52):
 * rows[y] -> width*4 RGBA bytes, modified in place; row_filters (may be NULL,
 * which makes every row "adaptive") receives libpng filter masks. */
int oracle_optimize_with_rows(uint8_t **rows, uint32_t width, uint32_t height,
                              uint8_t *row_filters, uint8_t strength, long bleed,
                              oracle_trace *trace);

/* Same contract as optimize_image (src/pngloss_image.c:159) on a packed
 * bytes_per_pixel image with an explicit stride. */
int oracle_optimize_image(uint8_t *pixels, uint32_t width, uint32_t height,
                          uint32_t bytes_per_pixel, uint64_t stride,
                          uint8_t *row_filters, uint8_t strength, long bleed,
                          oracle_trace *trace);

/* The 5 x 256 histogram optimize_state_init builds over the untouched image
 * (src/optimize_state.c:66-83). */
void oracle_original_frequency(const uint8_t *pixels, uint32_t width, uint32_t height,
                               uint32_t bytes_per_pixel, uint64_t stride,
                               uint32_t out[5][256]);

/* libpng's min-sum-of-abs heuristic as re-implemented by the reference
 * (src/optimize_state.c:492-562).  above may be NULL. */
int oracle_adaptive_filter(const uint8_t *above, const uint8_t *row,
                           uint32_t width, uint32_t bytes_per_pixel);

/* Test instrumentation: number of int16 error-cell stores that wrapped since the last reset (not thread safe). */
uint64_t oracle_int16_wraps(int reset);

/* Stateless synthetic image generator shared by tests and bench. */
void oracle_synth_rgba(uint8_t *dst, uint32_t width, uint32_t height, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
