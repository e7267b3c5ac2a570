/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.  See pngloss_oracle.h.
 *
 * Plain, single-threaded C restatement of the reference hot path.  It is kept
 * deliberately close to the *semantics* of the reference (packed 1/2/3/4
 * byte-per-pixel working image, three whole error rows, a straight sweep for
 * the row cost) so that it checks the restructurings the CUDA path relies on
 * (RGBA layout + channel mask, histogram-delta cost, streamed error windows,
 * post-pass derivative error) instead of sharing them.
 */
#include "pngloss_oracle.h"

#include <stdlib.h>
#include <string.h>

enum { F_NONE = 0, F_SUB = 1, F_UP = 2, F_AVG = 3, F_PAETH = 4, F_COUNT = 5 };
enum { ERR_PAD = 5, ERR_ROWS = 3 };

/* reference src/optimize_state.c:575-613 */
static inline int predict(int filter, int above, int diag, int left) {
    switch (filter) {
    case F_SUB: return left;
    case F_UP: return above;
    case F_AVG: return (above + left) / 2;
    case F_PAETH: {
        int p = above - diag;
        int q = left - diag;
        int dist_left = p < 0 ? -p : p;
        int dist_above = q < 0 ? -q : q;
        int dist_diag = (p + q) < 0 ? -(p + q) : (p + q);
        if (dist_left <= dist_above && dist_left <= dist_diag) return left;
        if (dist_above <= dist_diag) return above;
        return diag;
    }
    default: return 0;
    }
}

/* reference src/optimize_state.c:564-572 (returns the bit length) */
static inline unsigned bit_length_u64(uint64_t v) {
    unsigned n = 0;
    while (v) { v >>= 1; n++; }
    return n;
}

/* reference src/color_delta.c:4-41; result = b - a, narrowed to int16 on store */
static inline void lane_delta(uint32_t bpp, int16_t out[4], const int a[4], const int b[4]) {
    switch (bpp) {
    case 1:
        out[0] = out[1] = out[2] = (int16_t)(b[0] - a[0]);
        out[3] = 0;
        break;
    case 2:
        out[0] = out[1] = out[2] = (int16_t)(b[0] - a[0]);
        out[3] = (int16_t)(b[1] - a[1]);
        break;
    case 3:
        out[0] = (int16_t)(b[0] - a[0]);
        out[1] = (int16_t)(b[1] - a[1]);
        out[2] = (int16_t)(b[2] - a[2]);
        out[3] = 0;
        break;
    default:
        out[0] = (int16_t)(b[0] - a[0]);
        out[1] = (int16_t)(b[1] - a[1]);
        out[2] = (int16_t)(b[2] - a[2]);
        out[3] = (int16_t)(b[3] - a[3]);
        break;
    }
}

/* reference src/color_delta.c:43-66 as used at src/optimize_state.c:266-285 */
static inline uint32_t neighbour_error(uint32_t bpp, const int orig[4], const int back[4],
                                       const int old_n[4], const int new_n[4]) {
    int16_t old_partial[4], new_partial[4];
    lane_delta(bpp, old_partial, orig, old_n);
    lane_delta(bpp, new_partial, back, new_n);
    uint32_t total = 0;
    for (int i = 0; i < 4; i++) {
        int16_t d2 = (int16_t)(old_partial[i] - new_partial[i]);
        total += (uint32_t)((int)d2 * (int)d2);
    }
    return total;
}

typedef struct {
    uint8_t *row;      /* candidate output row, width*bpp bytes */
    int16_t *err;      /* ERR_ROWS * (width + ERR_PAD) cells of 4 lanes */
    uint32_t freq[256];
} candidate;

typedef struct {
    uint8_t *pixels;   /* working image, modified in place row by row */
    uint64_t stride;
    uint32_t width, height, bpp;
    const uint8_t *old_above; /* original bytes of row y-1 (reference: last_row_pixels) */
    uint32_t (*orig_freq)[256];
    long bleed;
} job;

static int candidate_alloc(candidate *c, const job *j) {
    c->row = calloc((size_t)j->width, j->bpp);
    c->err = calloc((size_t)ERR_ROWS * (j->width + ERR_PAD) * 4, sizeof(int16_t));
    memset(c->freq, 0, sizeof c->freq);
    return (c->row && c->err) ? 0 : -1;
}

static void candidate_free(candidate *c) {
    free(c->row);
    free(c->err);
}

static void candidate_assign(candidate *to, const candidate *from, const job *j) {
    memcpy(to->row, from->row, (size_t)j->width * j->bpp);
    memcpy(to->err, from->err, (size_t)ERR_ROWS * (j->width + ERR_PAD) * 4 * sizeof(int16_t));
    memcpy(to->freq, from->freq, sizeof to->freq);
}

/* reference src/optimize_state.c:390-467 (the live "sierra dithering" block) */
/* Test instrumentation: how often a store into an int16 error cell did not fit 16 bits and wrapped (the reference
 * relies on that narrowing, src/color_delta.h:6; tests use the count to prove that a vector exercises it). */
static uint64_t g_int16_wraps;
uint64_t oracle_int16_wraps(int reset) {
    const uint64_t n = g_int16_wraps;
    if (reset) g_int16_wraps = 0;
    return n;
}
static inline int16_t narrow16(long v) {
    if (v < -32768 || v > 32767) g_int16_wraps++;
    return (int16_t)v;
}

static void diffuse(candidate *c, const job *j, uint32_t x, const int16_t diff[4]) {
    const size_t ew = (size_t)j->width + ERR_PAD;
    int16_t *r0 = c->err + (0 * ew + x) * 4;
    int16_t *r1 = c->err + (1 * ew + x) * 4;
    int16_t *r2 = c->err + (2 * ew + x) * 4;
    for (int lane = 0; lane < 4; lane++) {
        long d = diff[lane];
        d = d / j->bleed;

        long twos = d / 16;
        d -= twos * 4;
        r1[0 * 4 + lane] = narrow16(r1[0 * 4 + lane] + twos);
        r1[4 * 4 + lane] = narrow16(r1[4 * 4 + lane] + twos);
        r2[1 * 4 + lane] = narrow16(r2[1 * 4 + lane] + twos);
        r2[3 * 4 + lane] = narrow16(r2[3 * 4 + lane] + twos);

        long threes = d / 8;
        d -= threes * 2;
        r0[4 * 4 + lane] = narrow16(r0[4 * 4 + lane] + threes);
        r2[2 * 4 + lane] = narrow16(r2[2 * 4 + lane] + threes);

        long fours = d * 2 / 9;
        d -= fours * 2;
        r1[1 * 4 + lane] = narrow16(r1[1 * 4 + lane] + fours);
        r1[3 * 4 + lane] = narrow16(r1[3 * 4 + lane] + fours);

        long five = d / 2;
        d -= five;
        r1[2 * 4 + lane] = narrow16(r1[2 * 4 + lane] + five);

        r0[3 * 4 + lane] = narrow16(r0[3 * 4 + lane] + d);
    }
}

/* reference src/optimize_state.c:492-562 */
int oracle_adaptive_filter(const uint8_t *above, const uint8_t *row,
                           uint32_t width, uint32_t bpp) {
    uint32_t sums[F_COUNT] = {0, 0, 0, 0, 0};
    const uint32_t n = width * bpp;
    for (uint32_t i = 0; i < n; i++) {
        int a = above ? above[i] : 0;
        int l = 0, d = 0;
        if (i >= bpp) {
            l = row[i - bpp];
            if (above) d = above[i - bpp];
        }
        int here = row[i];
        for (int f = 0; f < F_COUNT; f++) {
            uint8_t r = (uint8_t)(here - predict(f, a, d, l));
            sums[f] += (r < 128) ? r : (256u - r);
        }
    }
    uint32_t lowest = sums[0];
    for (int f = 1; f < F_COUNT; f++)
        if (sums[f] < lowest) lowest = sums[f];
    for (int f = 0; f < F_COUNT; f++)
        if (lowest >= sums[f]) return f;
    return F_COUNT;
}

/* One pixel: reference src/optimize_state.c:114-290.  Returns the derivative error. */
static uint64_t quantise_pixel(candidate *c, const job *j, uint32_t x, uint32_t y,
                               int filter, int strength) {
    const uint32_t bpp = j->bpp;
    const uint8_t *orig_row = j->pixels + (size_t)y * j->stride;
    const uint8_t *new_above_row = y ? j->pixels + (size_t)(y - 1) * j->stride : NULL;
    const size_t ew = (size_t)j->width + ERR_PAD;
    const int step = strength + 1;

    int orig[4] = {0}, here[4] = {0}, back[4] = {0};
    int old_above[4] = {0}, new_above[4] = {0}, old_diag[4] = {0}, new_diag[4] = {0};
    int old_left[4] = {0}, new_left[4] = {0};

    const int transparent = (bpp % 2 == 0) && orig_row[(size_t)x * bpp + bpp - 1] == 0;

    for (uint32_t ch = 0; ch < bpp; ch++) {
        const size_t off = (size_t)x * bpp + ch;
        orig[ch] = orig_row[off];
        if (y > 0) {
            new_above[ch] = new_above_row[off];
            old_above[ch] = j->old_above[off];
            if (x > 0) {
                new_diag[ch] = new_above_row[off - bpp];
                old_diag[ch] = j->old_above[off - bpp];
            }
        }
        if (x > 0) {
            new_left[ch] = c->row[off - bpp];
            old_left[ch] = orig_row[off - bpp];
        }

        int predicted = predict(filter, new_above[ch], new_diag[ch], new_left[ch]);
        uint8_t chosen;
        if (transparent && ch == bpp - 1) {
            /* :158-164 keep fully transparent pixels fully transparent */
            here[ch] = 0;
            back[ch] = 0;
            chosen = (uint8_t)(0 - predicted);
        } else {
            const int lane = (bpp == 2 && ch == 1) ? 3 : (int)ch;      /* :167-171 */
            const int carried = c->err[(0 * ew + x + 2) * 4 + lane];   /* :172 */
            here[ch] = orig[ch] + carried;

            int exact = orig[ch] - predicted;                          /* :175-182 */
            if (exact < -128) {
                predicted -= 256;
                exact = orig[ch] - predicted;
            } else if (exact > 127) {
                predicted += 256;
                exact = orig[ch] - predicted;
            }
            const int wanted = here[ch] - predicted;

            int lo, hi;                                                /* :186-193 */
            if (wanted < 0) {
                hi = -(-wanted - (-wanted % step));
                lo = hi - strength;
            } else {
                lo = wanted - (wanted % step);
                hi = lo + strength;
            }
            if (lo + predicted < 0) lo = -predicted;                   /* :195-200 */
            if (hi + predicted > 255) hi = 255 - predicted;
            if (hi < lo) {                                             /* :201-210 */
                if (wanted + predicted > 255) lo = hi = 255 - predicted;
                if (wanted + predicted < 0) lo = hi = -predicted;
            }

            /* :212-244 ascending scan, replace only on a strictly better key */
            int have = 0;
            uint32_t top_freq = 0;
            int top_symbol = 0;
            for (int s = lo; s <= hi; s++) {
                const uint32_t f = c->freq[(uint8_t)s];
                int better = 0;
                if (!have) {
                    better = 1;
                } else if (f > top_freq) {
                    better = 1;
                } else if (f == top_freq) {
                    const uint32_t prior_top = j->orig_freq[filter][(uint8_t)top_symbol];
                    const uint32_t prior = j->orig_freq[filter][(uint8_t)s];
                    if (prior > prior_top) better = 1;
                    else if (prior == prior_top && s == exact) better = 1;
                }
                if (better) {
                    have = 1;
                    top_freq = f;
                    top_symbol = s;
                }
            }
            if (!have) abort();                                        /* :245-248 */
            back[ch] = top_symbol + predicted;
            if (back[ch] < 0 || back[ch] > 255) abort();               /* :216-219 */
            chosen = (uint8_t)top_symbol;
        }
        c->row[off] = (uint8_t)back[ch];
        c->freq[chosen]++;                                             /* :253 */
    }

    int16_t diff[4];
    lane_delta(bpp, diff, back, here);                                 /* :258-259 */
    diffuse(c, j, x, diff);

    uint64_t e = neighbour_error(bpp, orig, back, old_above, new_above);
    e += neighbour_error(bpp, orig, back, old_diag, new_diag);
    e += neighbour_error(bpp, orig, back, old_left, new_left);
    return e;
}

/* One candidate row: reference src/optimize_state.c:292-361 */
static uint64_t quantise_row(candidate *c, const job *j, uint32_t y,
                             int filter, int strength, int adaptive) {
    uint64_t distortion = 0;
    for (uint32_t x = 0; x < j->width; x++)
        distortion += quantise_pixel(c, j, x, y, filter, strength);

    const uint8_t *above = y ? j->pixels + (size_t)(y - 1) * j->stride : NULL;
    if (adaptive && oracle_adaptive_filter(above, c->row, j->width, j->bpp) != filter)
        return UINT64_MAX;

    uint32_t bits = 0;
    for (uint32_t x = 0; x < j->width; x++) {
        for (uint32_t ch = 0; ch < j->bpp; ch++) {
            const size_t off = (size_t)x * j->bpp + ch;
            int l = x ? c->row[off - j->bpp] : 0;
            int a = above ? above[off] : 0;
            int d = (above && x) ? above[off - j->bpp] : 0;
            uint8_t symbol = (uint8_t)(c->row[off] - (uint8_t)predict(filter, a, d, l));
            uint32_t f = c->freq[symbol];
            if (f) bits += bit_length_u64(UINT64_MAX / f);
        }
    }

    const size_t ew = (size_t)j->width + ERR_PAD;
    memmove(c->err, c->err + ew * 4, (ERR_ROWS - 1) * ew * 4 * sizeof(int16_t));
    memset(c->err + (ERR_ROWS - 1) * ew * 4, 0, ew * 4 * sizeof(int16_t));

    return distortion / 128 + bits;
}

void oracle_original_frequency(const uint8_t *pixels, uint32_t width, uint32_t height,
                               uint32_t bpp, uint64_t stride, uint32_t out[5][256]) {
    memset(out, 0, 5 * 256 * sizeof(uint32_t));
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t *row = pixels + (size_t)y * stride;
        const uint8_t *above = y ? pixels + (size_t)(y - 1) * stride : NULL;
        for (uint32_t i = 0; i < width * bpp; i++) {
            int l = i >= bpp ? row[i - bpp] : 0;
            int a = above ? above[i] : 0;
            int d = (above && i >= bpp) ? above[i - bpp] : 0;
            for (int f = 0; f < F_COUNT; f++)
                out[f][(uint8_t)(row[i] - (uint8_t)predict(f, a, d, l))]++;
        }
    }
}

static const uint8_t png_mask[F_COUNT] = {0x08, 0x10, 0x20, 0x40, 0x80};

int oracle_optimize_image(uint8_t *pixels, uint32_t width, uint32_t height,
                          uint32_t bpp, uint64_t stride,
                          uint8_t *row_filters, uint8_t strength, long bleed,
                          oracle_trace *trace) {
    job j = {.pixels = pixels, .stride = stride, .width = width, .height = height,
             .bpp = bpp, .bleed = bleed};
    candidate current = {0}, best = {0}, trial = {0};
    uint32_t (*orig_freq)[256] = malloc(5 * 256 * sizeof(uint32_t));
    uint8_t *old_above = calloc((size_t)width, bpp);
    int rc = ORACLE_OK;

    j.orig_freq = orig_freq;
    j.old_above = old_above;
    if (!orig_freq || !old_above || candidate_alloc(&current, &j) ||
        candidate_alloc(&best, &j) || candidate_alloc(&trial, &j)) {
        rc = ORACLE_OUT_OF_MEMORY;
        goto out;
    }
    oracle_original_frequency(pixels, width, height, bpp, stride, orig_freq);
    if (trace) memcpy(trace->original_frequency, orig_freq, 5 * 256 * sizeof(uint32_t));

    for (uint32_t y = 0; y < height && rc == ORACLE_OK; y++) {
        const int adaptive = (!row_filters || y == 0);      /* src/pngloss_image.c:210 */
        int s = strength;
        int winner = -1;
        uint64_t lowest = UINT64_MAX;
        for (;;) {
            for (int f = 0; f < F_COUNT; f++) {
                candidate_assign(&trial, &current, &j);
                uint64_t cost = quantise_row(&trial, &j, y, f, s, adaptive);
                if (trace && trace->row_costs) trace->row_costs[(size_t)y * 5 + f] = cost;
                if (cost < lowest) {                        /* :257 strict, lowest index wins ties */
                    lowest = cost;
                    winner = f;
                    candidate_assign(&best, &trial, &j);
                }
            }
            if (winner >= 0) break;
            if (s == 0) { rc = ORACLE_NO_ACCEPTABLE_ROW; break; }   /* :268-271 */
            s--;
        }
        if (rc != ORACLE_OK) break;
        if (trace && trace->row_strength) trace->row_strength[y] = (uint8_t)s;
        uint8_t *image_row = pixels + (size_t)y * stride;
        memcpy(old_above, image_row, (size_t)width * bpp);  /* :277-281 */
        memcpy(image_row, best.row, (size_t)width * bpp);   /* :282-286 */
        candidate_assign(&current, &best, &j);
        if (row_filters) row_filters[y] = png_mask[winner];
    }
    if (trace) memcpy(trace->final_frequency, current.freq, sizeof current.freq);

out:
    candidate_free(&current);
    candidate_free(&best);
    candidate_free(&trial);
    free(orig_freq);
    free(old_above);
    return rc;
}

/* reference src/pngloss_image.c:52-156 */
int oracle_optimize_with_rows(uint8_t **rows, uint32_t width, uint32_t height,
                              uint8_t *row_filters, uint8_t strength, long bleed,
                              oracle_trace *trace) {
    int gray = 1, opaque = 1;
    for (uint32_t y = 0; y < height && (gray || opaque); y++) {
        for (uint32_t x = 0; x < width; x++) {
            const uint8_t *p = rows[y] + (size_t)x * 4;
            if (p[0] != p[1] || p[1] != p[2]) gray = 0;
            if (p[3] < 255) opaque = 0;
        }
    }
    const uint32_t bpp = gray ? (opaque ? 1 : 2) : (opaque ? 3 : 4);
    const size_t stride = (size_t)width * bpp;
    uint8_t *packed = malloc(stride * height + 1);
    if (!packed) return ORACLE_OUT_OF_MEMORY;

    for (uint32_t y = 0; y < height; y++) {
        for (uint32_t x = 0; x < width; x++) {
            const uint8_t *p = rows[y] + (size_t)x * 4;
            uint8_t *q = packed + y * stride + (size_t)x * bpp;
            switch (bpp) {
            case 1: q[0] = p[1]; break;
            case 2: q[0] = p[1]; q[1] = p[3]; break;
            case 3: q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; break;
            default: memcpy(q, p, 4); break;
            }
        }
    }
    int rc = oracle_optimize_image(packed, width, height, bpp, stride, row_filters,
                                   strength, bleed, trace);
    if (rc == ORACLE_OK) {
        for (uint32_t y = 0; y < height; y++) {
            for (uint32_t x = 0; x < width; x++) {
                uint8_t *p = rows[y] + (size_t)x * 4;
                const uint8_t *q = packed + y * stride + (size_t)x * bpp;
                switch (bpp) {
                case 1: p[0] = p[1] = p[2] = q[0]; p[3] = 255; break;
                case 2: p[0] = p[1] = p[2] = q[0]; p[3] = q[1]; break;
                case 3: p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; p[3] = 255; break;
                default: memcpy(p, q, 4); break;
                }
            }
        }
    }
    free(packed);
    return rc;
}

static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* gradient + +-8 noise, alpha 128..255 with fully transparent holes. */
void oracle_synth_rgba(uint8_t *dst, uint32_t w, uint32_t h, uint64_t seed) {
    const uint64_t dx = w > 1 ? w - 1 : 1, dy = h > 1 ? h - 1 : 1;
    const uint64_t dxy = (w + h > 2) ? (uint64_t)w + h - 2 : 1;
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            int base[4];
            base[0] = (int)((uint64_t)x * 255 / dx);
            base[1] = (int)((uint64_t)y * 255 / dy);
            base[2] = (int)(((uint64_t)x + y) * 255 / dxy);
            base[3] = 255 - base[0] / 2;
            for (int c = 0; c < 4; c++) {
                uint64_t i = ((uint64_t)y * w + x) * 4 + c;
                int n = (int)(splitmix64((seed << 40) + i) % 17) - 8;
                int v = base[c] + n;
                v = v < 0 ? 0 : v > 255 ? 255 : v;
                if (c == 3 && (y / 64) % 4 == 0 && x < w / 16) v = 0;
                dst[i] = (uint8_t)v;
            }
        }
    }
}
