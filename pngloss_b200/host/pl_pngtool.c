/* Small command line around pl_png.c, used by the tests (and handy for debugging):
 *   pl_pngtool decode in.png out.rgba     -> 8 bytes (width, height, little endian) + RGBA8 pixels
 *   pl_pngtool encode in.rgba out.png [filters.bin]   filters.bin = height libpng masks, optional
 *   pl_pngtool copy in.png out.png        -> decode + re-encode with the heuristic filters (chunks pass through)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pl_png.h"

static int fail(const char *msg) {
    fprintf(stderr, "pl_pngtool: %s\n", msg);
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 4) return fail("usage: decode in.png out.rgba | encode in.rgba out.png [filters.bin] | copy in.png out.png");
    if (!strcmp(argv[1], "decode") || !strcmp(argv[1], "copy")) {
        FILE *in = fopen(argv[2], "rb");
        if (!in) return fail("cannot open input");
        png24_image img;
        pngloss_error rc = rwpng_read_image24(in, &img, false, true);
        fclose(in);
        if (rc) { fprintf(stderr, "pl_pngtool: read error %d\n", rc); return (int)rc; }
        FILE *out = fopen(argv[3], "wb");
        if (!out) return fail("cannot open output");
        if (!strcmp(argv[1], "decode")) {
            unsigned dims[2] = {img.width, img.height};
            fwrite(dims, sizeof dims, 1, out);
            fwrite(img.rgba_data, 4, (size_t)img.width * img.height, out);
        } else {
            rc = rwpng_write_image24(out, &img, NULL);
        }
        fclose(out);
        fprintf(stderr, "%ux%u input_color=%d gamma=%.5f file_size=%zu\n", img.width, img.height, img.input_color,
                img.gamma, img.file_size);
        rwpng_free_image24(&img);
        return (int)rc;
    }
    if (!strcmp(argv[1], "encode")) {
        FILE *in = fopen(argv[2], "rb");
        if (!in) return fail("cannot open input");
        unsigned dims[2];
        if (fread(dims, sizeof dims, 1, in) != 1) return fail("short input");
        png24_image img;
        memset(&img, 0, sizeof img);
        img.width = dims[0];
        img.height = dims[1];
        img.gamma = 0.45455;
        img.output_color = RWPNG_GAMA_ONLY;
        img.rgba_data = malloc((size_t)dims[0] * dims[1] * 4);
        img.row_pointers = malloc(sizeof(unsigned char *) * dims[1]);
        if (fread(img.rgba_data, 4, (size_t)dims[0] * dims[1], in) != (size_t)dims[0] * dims[1]) return fail("short input");
        fclose(in);
        for (unsigned y = 0; y < dims[1]; y++) img.row_pointers[y] = img.rgba_data + (size_t)y * dims[0] * 4;
        unsigned char *filters = NULL;
        if (argc > 4) {
            FILE *ff = fopen(argv[4], "rb");
            if (!ff) return fail("cannot open filters");
            filters = malloc(dims[1]);
            if (fread(filters, 1, dims[1], ff) != dims[1]) return fail("short filters");
            fclose(ff);
        }
        FILE *out = fopen(argv[3], "wb");
        if (!out) return fail("cannot open output");
        pngloss_error rc = rwpng_write_image24(out, &img, filters);
        fclose(out);
        free(filters);
        rwpng_free_image24(&img);
        return (int)rc;
    }
    return fail("unknown command");
}
