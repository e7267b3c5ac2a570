/* PNG reader / writer on zlib.  See pl_png.h. */
#include "pl_png.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static const unsigned char PNG_SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

static uint32_t be32(const unsigned char *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static void put_be32(unsigned char *p, uint32_t v) {
    p[0] = (unsigned char)(v >> 24);
    p[1] = (unsigned char)(v >> 16);
    p[2] = (unsigned char)(v >> 8);
    p[3] = (unsigned char)v;
}

void rwpng_version_info(FILE *fp) {
    fprintf(fp, "   PNG container handled by pngloss_b200 (pl_png.c) on zlib %s.\n", zlibVersion());
}

/* ---- PNG predictors ------------------------------------------------------------------------------- */
static inline int paeth(int a /*left*/, int b /*above*/, int c /*upper left*/) {
    int p = a + b - c;
    int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc) ? b : c;
}

static int unfilter_row(int type, unsigned char *row, const unsigned char *prev, size_t n, unsigned bpp) {
    switch (type) {
    case 0: break;
    case 1:
        for (size_t i = bpp; i < n; i++) row[i] = (unsigned char)(row[i] + row[i - bpp]);
        break;
    case 2:
        if (prev) for (size_t i = 0; i < n; i++) row[i] = (unsigned char)(row[i] + prev[i]);
        break;
    case 3:
        for (size_t i = 0; i < n; i++) {
            int l = i >= bpp ? row[i - bpp] : 0, u = prev ? prev[i] : 0;
            row[i] = (unsigned char)(row[i] + ((l + u) >> 1));
        }
        break;
    case 4:
        for (size_t i = 0; i < n; i++) {
            int l = i >= bpp ? row[i - bpp] : 0, u = prev ? prev[i] : 0, ul = (prev && i >= bpp) ? prev[i - bpp] : 0;
            row[i] = (unsigned char)(row[i] + paeth(l, u, ul));
        }
        break;
    default: return -1;
    }
    return 0;
}

static void filter_row(int type, unsigned char *dst, const unsigned char *row, const unsigned char *prev,
                       size_t n, unsigned bpp) {
    for (size_t i = 0; i < n; i++) {
        int l = i >= bpp ? row[i - bpp] : 0, u = prev ? prev[i] : 0, ul = (prev && i >= bpp) ? prev[i - bpp] : 0;
        int pred = type == 1 ? l : type == 2 ? u : type == 3 ? ((l + u) >> 1) : type == 4 ? paeth(l, u, ul) : 0;
        dst[i] = (unsigned char)(row[i] - pred);
    }
}

int rwpng_heuristic_filter(const unsigned char *prev, const unsigned char *row, size_t n, unsigned bpp) {
    unsigned long best = ULONG_MAX;
    int pick = 0;
    for (int type = 0; type < 5; type++) {
        unsigned long sum = 0;
        for (size_t i = 0; i < n; i++) {
            int l = i >= bpp ? row[i - bpp] : 0, u = prev ? prev[i] : 0, ul = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int pred = type == 1 ? l : type == 2 ? u : type == 3 ? ((l + u) >> 1) : type == 4 ? paeth(l, u, ul) : 0;
            unsigned char v = (unsigned char)(row[i] - pred);
            sum += v < 128 ? v : 256u - v;
        }
        if (sum < best) { best = sum; pick = type; }
    }
    return pick;
}

/* ---- reader -------------------------------------------------------------------------------------------- */
static const char *const KNOWN_DROPPED[] = {"bKGD", "cHRM", "eXIf", "gAMA", "hIST", "iCCP", "oFFs", "pCAL",
                                            "sBIT", "sCAL", "sPLT", "sRGB", "sTER", "tIME", "tRNS", "PLTE",
                                            "IHDR", "IDAT", "IEND", NULL};

static bool chunk_is_passed_through(const unsigned char *name) {
    for (int i = 0; KNOWN_DROPPED[i]; i++)
        if (memcmp(name, KNOWN_DROPPED[i], 4) == 0) return false;
    return true;   /* pHYs, iTXt, tEXt, zTXt and everything libpng does not know (src/rwpng.c:129-156,263) */
}

static unsigned char *read_all(FILE *f, size_t *len) {
    size_t cap = 1 << 16, n = 0;
    unsigned char *buf = malloc(cap);
    if (!buf) return NULL;
    for (;;) {
        if (n == cap) {
            unsigned char *nb = realloc(buf, cap *= 2);
            if (!nb) { free(buf); return NULL; }
            buf = nb;
        }
        size_t got = fread(buf + n, 1, cap - n, f);
        n += got;
        if (!got) break;
    }
    *len = n;
    return buf;
}

/* Adam7 pass geometry */
static const int A7_X0[7] = {0, 4, 0, 2, 0, 1, 0}, A7_Y0[7] = {0, 0, 4, 0, 2, 0, 1};
static const int A7_DX[7] = {8, 8, 4, 4, 2, 2, 1}, A7_DY[7] = {8, 8, 8, 4, 4, 2, 2};

typedef struct {
    uint32_t w, h;
    int depth, ctype, interlace, channels;
    unsigned char plte[256 * 3];
    int nplte;
    unsigned char trns[256];
    int ntrns;
    bool has_trns;
    unsigned trns_gray, trns_r, trns_g, trns_b;
} png_header;

/* sample k (0-based, depth bits each) of a row, as stored */
static inline unsigned sample_at(const unsigned char *row, size_t k, int depth) {
    if (depth == 8) return row[k];
    if (depth == 16) return ((unsigned)row[2 * k] << 8) | row[2 * k + 1];
    size_t bit = k * (size_t)depth;
    return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
}

static void expand_pixel(const png_header *hd, const unsigned char *row, size_t x, unsigned char *out) {
    const int d = hd->depth;
    switch (hd->ctype) {
    case 0: {
        unsigned v = sample_at(row, x, d);
        unsigned char g = d == 16 ? (unsigned char)(v >> 8) : d == 8 ? (unsigned char)v
                          : (unsigned char)(v * (255u / ((1u << d) - 1u)));
        out[0] = out[1] = out[2] = g;
        out[3] = (hd->has_trns && v == hd->trns_gray) ? 0 : 255;
        break;
    }
    case 2: {
        unsigned r = sample_at(row, 3 * x, d), g = sample_at(row, 3 * x + 1, d), b = sample_at(row, 3 * x + 2, d);
        out[0] = (unsigned char)(d == 16 ? r >> 8 : r);
        out[1] = (unsigned char)(d == 16 ? g >> 8 : g);
        out[2] = (unsigned char)(d == 16 ? b >> 8 : b);
        out[3] = (hd->has_trns && r == hd->trns_r && g == hd->trns_g && b == hd->trns_b) ? 0 : 255;
        break;
    }
    case 3: {
        unsigned i = sample_at(row, x, d);
        if ((int)i >= hd->nplte) { out[0] = out[1] = out[2] = 0; out[3] = 255; break; }
        out[0] = hd->plte[3 * i];
        out[1] = hd->plte[3 * i + 1];
        out[2] = hd->plte[3 * i + 2];
        out[3] = (int)i < hd->ntrns ? hd->trns[i] : 255;
        break;
    }
    case 4: {
        unsigned g = sample_at(row, 2 * x, d), a = sample_at(row, 2 * x + 1, d);
        out[0] = out[1] = out[2] = (unsigned char)(d == 16 ? g >> 8 : g);
        out[3] = (unsigned char)(d == 16 ? a >> 8 : a);
        break;
    }
    default: {
        for (int c = 0; c < 4; c++) {
            unsigned v = sample_at(row, 4 * x + c, d);
            out[c] = (unsigned char)(d == 16 ? v >> 8 : v);
        }
        break;
    }
    }
}

static void free_chunks(struct rwpng_chunk *c) {
    while (c) {
        struct rwpng_chunk *n = c->next;
        free(c->data);
        free(c);
        c = n;
    }
}

void rwpng_free_image24(png24_image *image) {
    free(image->row_pointers);
    image->row_pointers = NULL;
    free(image->rgba_data);
    image->rgba_data = NULL;
    free_chunks(image->chunks);
    image->chunks = NULL;
}

pngloss_error rwpng_read_image24(FILE *infile, png24_image *out, bool strip, bool verbose) {
    size_t len = 0;
    unsigned char *file = read_all(infile, &len);
    if (!file) return PNG_OUT_OF_MEMORY_ERROR;
    pngloss_error rc = LIBPNG_FATAL_ERROR;
    unsigned char *zdata = NULL, *raw = NULL;
    size_t zlen = 0, zcap = 0;
    struct rwpng_chunk *chunks = NULL, **chunk_tail = &chunks;
    png_header hd;
    memset(&hd, 0, sizeof hd);
    bool have_ihdr = false, have_plte = false, have_idat = false, have_iend = false, srgb = false;
    double gamma = 0.45455;

    memset(out, 0, sizeof *out);
    if (len < 8 || memcmp(file, PNG_SIG, 8) != 0) {
        if (verbose) fprintf(stderr, "  error: not a PNG file\n");
        goto done;
    }
    for (size_t pos = 8; pos + 12 <= len && !have_iend;) {
        const uint32_t clen = be32(file + pos);
        const unsigned char *name = file + pos + 4, *data = file + pos + 8;
        if (clen > 0x7fffffffu || pos + 12 + (size_t)clen > len) goto done;
        const bool crc_ok = be32(data + clen) == (uint32_t)crc32(crc32(0, name, 4), data, clen);
        const bool critical = !(name[0] & 0x20);
        pos += 12 + (size_t)clen;
        if (!crc_ok) {
            if (critical) goto done;
            if (verbose) fprintf(stderr, "  libpng warning: CRC error in %.4s\n", name);
            continue;
        }
        if (!have_ihdr && memcmp(name, "IHDR", 4) != 0) goto done;
        if (memcmp(name, "IHDR", 4) == 0) {
            if (clen != 13 || have_ihdr) goto done;
            hd.w = be32(data);
            hd.h = be32(data + 4);
            hd.depth = data[8];
            hd.ctype = data[9];
            hd.interlace = data[12];
            if (!hd.w || !hd.h || hd.w > 0x7fffffffu || hd.h > 0x7fffffffu || data[10] || data[11] || hd.interlace > 1)
                goto done;
            static const int chan_of[7] = {1, 0, 3, 1, 2, 0, 4};
            if (hd.ctype > 6 || !chan_of[hd.ctype]) goto done;
            hd.channels = chan_of[hd.ctype];
            const int d = hd.depth;
            const bool depth_ok = hd.ctype == 0 ? (d == 1 || d == 2 || d == 4 || d == 8 || d == 16)
                                  : hd.ctype == 3 ? (d == 1 || d == 2 || d == 4 || d == 8) : (d == 8 || d == 16);
            if (!depth_ok) goto done;
            have_ihdr = true;
        } else if (memcmp(name, "PLTE", 4) == 0) {
            if (clen % 3 || clen > 768 || have_idat) goto done;
            memcpy(hd.plte, data, clen);
            hd.nplte = (int)(clen / 3);
            have_plte = true;
        } else if (memcmp(name, "tRNS", 4) == 0) {
            if (hd.ctype == 3) {
                hd.ntrns = clen > 256 ? 256 : (int)clen;
                memcpy(hd.trns, data, (size_t)hd.ntrns);
            } else if (hd.ctype == 0 && clen >= 2) {
                hd.has_trns = true;
                hd.trns_gray = ((unsigned)data[0] << 8 | data[1]) & ((1u << hd.depth) - 1u);
            } else if (hd.ctype == 2 && clen >= 6) {
                const unsigned m = (1u << hd.depth) - 1u;
                hd.has_trns = true;
                hd.trns_r = ((unsigned)data[0] << 8 | data[1]) & m;
                hd.trns_g = ((unsigned)data[2] << 8 | data[3]) & m;
                hd.trns_b = ((unsigned)data[4] << 8 | data[5]) & m;
            }
        } else if (memcmp(name, "sRGB", 4) == 0) {
            srgb = true;
        } else if (memcmp(name, "gAMA", 4) == 0 && clen == 4) {
            gamma = be32(data) / 100000.0;
        } else if (memcmp(name, "IDAT", 4) == 0) {
            if (zlen + clen > zcap) {
                zcap = (zlen + clen) * 2 + 4096;
                unsigned char *nz = realloc(zdata, zcap);
                if (!nz) { rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }
                zdata = nz;
            }
            memcpy(zdata + zlen, data, clen);
            zlen += clen;
            have_idat = true;
        } else if (memcmp(name, "IEND", 4) == 0) {
            have_iend = true;
        } else if (!strip && chunk_is_passed_through(name)) {
            struct rwpng_chunk *c = calloc(1, sizeof *c);
            if (!c) { rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }
            memcpy(c->name, name, 4);
            c->size = clen;
            c->location = have_idat ? RWPNG_AFTER_IDAT : have_plte ? RWPNG_AFTER_PLTE : RWPNG_AFTER_IHDR;
            if (clen) {
                c->data = malloc(clen);
                if (!c->data) { free(c); rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }
                memcpy(c->data, data, clen);
            }
            *chunk_tail = c;
            chunk_tail = &c->next;
        } else if (critical && chunk_is_passed_through(name)) {
            goto done;   /* unknown critical chunk */
        }
    }
    if (!have_ihdr || !have_idat || (hd.ctype == 3 && !have_plte)) goto done;
    /* For overflow safety reject images that won't fit in 32-bit (reference src/rwpng.c:286-290) */
    if ((uint64_t)hd.w * 4 > (uint64_t)INT_MAX / hd.h) { rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }

    /* inflate */
    const size_t bits_pp = (size_t)hd.depth * hd.channels;
    size_t raw_len = 0;
    if (hd.interlace) {
        for (int p = 0; p < 7; p++) {
            const size_t pw = (hd.w + A7_DX[p] - 1 - A7_X0[p]) / A7_DX[p], ph = (hd.h + A7_DY[p] - 1 - A7_Y0[p]) / A7_DY[p];
            if (pw && ph) raw_len += ph * (1 + (pw * bits_pp + 7) / 8);
        }
    } else {
        raw_len = (size_t)hd.h * (1 + ((size_t)hd.w * bits_pp + 7) / 8);
    }
    raw = malloc(raw_len ? raw_len : 1);
    out->rgba_data = malloc((size_t)hd.w * hd.h * 4);
    out->row_pointers = malloc((size_t)hd.h * sizeof out->row_pointers[0]);
    if (!raw || !out->rgba_data || !out->row_pointers) { rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }
    {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit(&zs) != Z_OK) { rc = PNG_OUT_OF_MEMORY_ERROR; goto done; }
        size_t in_off = 0, out_off = 0;
        int zr = Z_OK;
        while (zr == Z_OK && out_off < raw_len) {
            const size_t in_now = zlen - in_off > 0x40000000u ? 0x40000000u : zlen - in_off;
            const size_t out_now = raw_len - out_off > 0x40000000u ? 0x40000000u : raw_len - out_off;
            zs.next_in = zdata + in_off;
            zs.avail_in = (uInt)in_now;
            zs.next_out = raw + out_off;
            zs.avail_out = (uInt)out_now;
            zr = inflate(&zs, Z_NO_FLUSH);
            in_off += in_now - zs.avail_in;
            out_off += out_now - zs.avail_out;
            if (zr == Z_OK && in_now == zs.avail_in && out_now == zs.avail_out) break;   /* no progress */
        }
        inflateEnd(&zs);
        if (out_off != raw_len || (zr != Z_OK && zr != Z_STREAM_END)) {
            if (verbose) fprintf(stderr, "  error: corrupt image data (libpng failed)\n");
            goto done;
        }
    }
    for (uint32_t y = 0; y < hd.h; y++) out->row_pointers[y] = out->rgba_data + (size_t)y * hd.w * 4;

    /* unfilter + expand */
    {
        const unsigned fbpp = (unsigned)((bits_pp + 7) / 8);
        const unsigned char *src = raw;
        const int npass = hd.interlace ? 7 : 1;
        for (int p = 0; p < npass; p++) {
            const int x0 = hd.interlace ? A7_X0[p] : 0, y0 = hd.interlace ? A7_Y0[p] : 0;
            const int dx = hd.interlace ? A7_DX[p] : 1, dy = hd.interlace ? A7_DY[p] : 1;
            const size_t pw = (hd.w + dx - 1 - x0) / dx, ph = (hd.h + dy - 1 - y0) / dy;
            if (!pw || !ph) continue;
            const size_t rb = (pw * bits_pp + 7) / 8;
            unsigned char *prev = NULL;
            for (size_t r = 0; r < ph; r++) {
                unsigned char *line = (unsigned char *)src + 1;
                if (unfilter_row(src[0], line, prev, rb, fbpp)) goto done;
                unsigned char *dst = out->row_pointers[y0 + r * dy];
                for (size_t i = 0; i < pw; i++) expand_pixel(&hd, line, i, dst + ((size_t)x0 + i * dx) * 4);
                prev = line;
                src += 1 + rb;
            }
        }
    }

    out->width = hd.w;
    out->height = hd.h;
    out->file_size = len;
    out->chunks = chunks;
    chunks = NULL;
    /* colour tagging as the reference decides it (src/rwpng.c:238-256) */
    if (srgb) {
        out->input_color = out->output_color = RWPNG_SRGB;
        gamma = 0.45455;
    } else if (gamma > 0 && gamma <= 1.0) {
        out->input_color = out->output_color = RWPNG_GAMA_ONLY;
    } else {
        fprintf(stderr, "pngloss readpng:  ignored out-of-range gamma %f\n", gamma);
        out->input_color = out->output_color = RWPNG_NONE;
        gamma = 0.45455;
    }
    out->gamma = gamma;
    rc = SUCCESS;

done:
    if (rc != SUCCESS) {
        free(out->rgba_data);
        free(out->row_pointers);
        memset(out, 0, sizeof *out);
    }
    free_chunks(chunks);
    free(zdata);
    free(raw);
    free(file);
    return rc;
}

/* ---- writer -------------------------------------------------------------------------------------------- */
typedef struct {
    FILE *f;
    size_t written, cap;
    pngloss_error rc;
} sink;

static void sink_write(sink *s, const void *p, size_t n) {
    /* stop writing once the size cap is exceeded (reference src/rwpng.c:85-105) */
    if (s->cap && s->written + n > s->cap) s->rc = TOO_LARGE_FILE;
    if (s->rc == SUCCESS && n && !fwrite(p, n, 1, s->f)) s->rc = CANT_WRITE_ERROR;
    s->written += n;
}

static void write_chunk(sink *s, const char *name, const unsigned char *data, size_t n) {
    unsigned char hdr[8], crc[4];
    put_be32(hdr, (uint32_t)n);
    memcpy(hdr + 4, name, 4);
    uint32_t c = (uint32_t)crc32(0, hdr + 4, 4);
    if (n) c = (uint32_t)crc32(c, data, (uInt)n);
    put_be32(crc, c);
    sink_write(s, hdr, 8);
    sink_write(s, data, n);
    sink_write(s, crc, 4);
}

static void write_passthrough(sink *s, png24_image *img, int location) {
    for (struct rwpng_chunk *c = img->chunks; c; c = c->next)
        if (c->location == location) {
            write_chunk(s, (const char *)c->name, c->data, c->size);
            img->metadata_size += c->size + 12;
        }
}

#define IDAT_BYTES 8192   /* libpng's default compression buffer: one IDAT chunk per 8 KB of deflate output */

/* The container around a stream of filtered scanlines: signature, IHDR, colour tags, passed-through chunks,
 * IDAT (zlib level 9, memLevel 9 - reference src/rwpng.c:471-472 - with the filtered-data strategy and a
 * window no larger than the data needs, 8 KB chunks), IEND.  next_row(ctx, y) returns row y: one filter-type
 * byte and width * bpp filtered bytes. */
typedef const unsigned char *(*row_source)(void *ctx, uint32_t y);

static pngloss_error write_png_stream(FILE *outfile, png24_image *img, unsigned bpp, row_source next_row, void *ctx) {
    const uint32_t w = img->width, h = img->height;
    sink s = {outfile, 0, img->maximum_file_size, SUCCESS};
    img->metadata_size = 0;
    const int ctype = bpp == 1 ? 0 : bpp == 2 ? 4 : bpp == 3 ? 2 : 6;
    const size_t rb = (size_t)w * bpp;
    unsigned char *zbuf = malloc(IDAT_BYTES);
    if (!zbuf) return OUT_OF_MEMORY_ERROR;

    sink_write(&s, PNG_SIG, 8);
    unsigned char ihdr[13];
    put_be32(ihdr, w);
    put_be32(ihdr + 4, h);
    ihdr[8] = 8; ihdr[9] = (unsigned char)ctype; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    write_chunk(&s, "IHDR", ihdr, 13);
    if (img->output_color == RWPNG_SRGB) {     /* reference src/rwpng.c:501-509 */
        unsigned char g[4], intent = 0;
        put_be32(g, (uint32_t)(img->gamma * 100000.0 + 0.5));
        write_chunk(&s, "gAMA", g, 4);
        write_chunk(&s, "sRGB", &intent, 1);
    }
    write_passthrough(&s, img, RWPNG_AFTER_IHDR);
    write_passthrough(&s, img, RWPNG_AFTER_PLTE);

    const size_t raw_len = (rb + 1) * (size_t)h;
    int wbits = 15;
    while (wbits > 8 && ((size_t)1 << (wbits - 1)) >= raw_len + 262) wbits--;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_BEST_COMPRESSION, Z_DEFLATED, wbits, 9, Z_FILTERED) != Z_OK) {
        free(zbuf);
        return LIBPNG_INIT_ERROR;
    }
    zs.next_out = zbuf;
    zs.avail_out = IDAT_BYTES;
    for (uint32_t y = 0; y < h; y++) {
        zs.next_in = (unsigned char *)next_row(ctx, y);
        zs.avail_in = (uInt)(rb + 1);
        while (zs.avail_in) {
            deflate(&zs, Z_NO_FLUSH);
            if (!zs.avail_out) {
                write_chunk(&s, "IDAT", zbuf, IDAT_BYTES);
                zs.next_out = zbuf;
                zs.avail_out = IDAT_BYTES;
            }
        }
    }
    for (;;) {
        const int zr = deflate(&zs, Z_FINISH);
        if (!zs.avail_out || zr == Z_STREAM_END) {
            if (IDAT_BYTES - zs.avail_out) write_chunk(&s, "IDAT", zbuf, IDAT_BYTES - zs.avail_out);
            zs.next_out = zbuf;
            zs.avail_out = IDAT_BYTES;
        }
        if (zr == Z_STREAM_END) break;
        if (zr != Z_OK && zr != Z_BUF_ERROR) { s.rc = LIBPNG_FATAL_ERROR; break; }
    }
    deflateEnd(&zs);
    write_passthrough(&s, img, RWPNG_AFTER_IDAT);
    write_chunk(&s, "IEND", NULL, 0);
    free(zbuf);

    if (s.rc == SUCCESS || s.rc == TOO_LARGE_FILE) img->file_size = s.written;
    return s.rc;
}

/* rows narrowed and filtered on the CPU */
struct cpu_rows {
    png24_image *img;
    unsigned char *row_filters;
    unsigned bpp;
    size_t rb;
    unsigned char *cur, *prev, *filt;
};

static const unsigned char *cpu_next_row(void *ctx, uint32_t y) {
    struct cpu_rows *c = ctx;
    const unsigned bpp = c->bpp;
    if (y) { unsigned char *t = c->prev; c->prev = c->cur; c->cur = t; }
    const unsigned char *p = c->img->row_pointers[y];
    for (uint32_t x = 0; x < c->img->width; x++, p += 4) {   /* narrow to the colour type (G is luminance) */
        unsigned char *q = c->cur + (size_t)x * bpp;
        switch (bpp) {
        case 1: q[0] = p[1]; break;
        case 2: q[0] = p[1]; q[1] = p[3]; break;
        case 3: q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; break;
        default: memcpy(q, p, 4); break;
        }
    }
    /* row 0 is always chosen by the heuristic (reference src/rwpng.c:488-495); later rows use the
     * caller's explicit filter */
    int type;
    if (c->row_filters && y > 0) {
        const unsigned m = c->row_filters[y];
        type = m == 0x10 ? 1 : m == 0x20 ? 2 : m == 0x40 ? 3 : m == 0x80 ? 4 : m == 0x08 ? 0 : -1;
        if (type < 0) type = rwpng_heuristic_filter(c->prev, c->cur, c->rb, bpp);
    } else {
        type = rwpng_heuristic_filter(y ? c->prev : NULL, c->cur, c->rb, bpp);
    }
    c->filt[0] = (unsigned char)type;
    filter_row(type, c->filt + 1, c->cur, y ? c->prev : NULL, c->rb, bpp);
    return c->filt;
}

pngloss_error rwpng_write_image24(FILE *outfile, png24_image *img, unsigned char *row_filters) {
    const uint32_t w = img->width, h = img->height;
    /* autodetect grayscale and alpha on the pixels being written (reference src/rwpng.c:557-573) */
    bool gray = true, opaque = true;
    for (uint32_t y = 0; y < h && (gray || opaque); y++) {
        const unsigned char *p = img->row_pointers[y];
        for (uint32_t x = 0; x < w; x++, p += 4) {
            if (p[0] != p[1] || p[1] != p[2]) gray = false;
            if (p[3] < 255) opaque = false;
        }
    }
    struct cpu_rows c = {img, row_filters, gray ? (opaque ? 1 : 2) : (opaque ? 3 : 4), 0, NULL, NULL, NULL};
    c.rb = (size_t)w * c.bpp;
    c.cur = malloc(c.rb ? c.rb : 1);
    c.prev = malloc(c.rb ? c.rb : 1);
    c.filt = malloc(c.rb + 1);
    pngloss_error rc = OUT_OF_MEMORY_ERROR;
    if (c.cur && c.prev && c.filt) rc = write_png_stream(outfile, img, c.bpp, cpu_next_row, &c);
    free(c.cur); free(c.prev); free(c.filt);
    return rc;
}

/* rows narrowed and filtered on the GPU (pngloss_b200_image.scanlines) */
struct gpu_rows {
    const unsigned char *scan;
    size_t stride;
};

static const unsigned char *gpu_next_row(void *ctx, uint32_t y) {
    const struct gpu_rows *g = ctx;
    return g->scan + (size_t)y * g->stride;
}

pngloss_error rwpng_write_scanlines(FILE *outfile, png24_image *img, unsigned bytes_per_pixel,
                                    const unsigned char *scanlines) {
    if (bytes_per_pixel < 1 || bytes_per_pixel > 4 || !scanlines) return INVALID_ARGUMENT;
    struct gpu_rows g = {scanlines, 1 + (size_t)img->width * bytes_per_pixel};
    return write_png_stream(outfile, img, bytes_per_pixel, gpu_next_row, &g);
}
