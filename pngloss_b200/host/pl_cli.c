/* pngloss - command line driver on top of libpngloss_b200.
 *
 * Keeps the reference's command line surface (src/pngloss.c:28-160, src/pngloss_opts.c:22-136): the same
 * options, defaults (strength 19, bleed 2, "-loss.png"), validation messages, stdin/stdout mode, output
 * naming, temp-file + rename, verbose messages and exit codes.  What changes is the shape of the work:
 * the reference decodes, optimises and encodes one file at a time (src/pngloss.c:173-205); one image is
 * five busy warps on a B200, so this driver decodes every input first, hands all images to the GPU in
 * one pngloss_b200_optimize_batch() call and then encodes.  Per-file results are identical.
 */
#include <getopt.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/pngloss_b200.h"
#include "pl_png.h"

static const char *USAGE = "\
usage:  pngloss [options] -- pngfile [pngfile ...]\n\
        pngloss [options] - >stdout <stdin\n\n\
options:\n\
  -s, --strength 19 how much quality to sacrifice, from 0 to 100 (default 19)\n\
  -b, --bleed 2     bleed divider, from 1 (full dithering) to 32767 (none)\n\
  -f, --force       overwrite existing output files\n\
  -o, --output file destination file path to use instead of --ext\n\
  -v, --verbose     print status messages\n\
  -q, --quiet       don't print status messages (default, overrides -v)\n\
  -V, --version     print version number\n\
  --skip-if-larger  only save converted files if they're smaller than original\n\
  --ext new.png     set custom suffix/extension for output filenames\n\
  --strip           remove optional metadata\n\
\n\
Lossily compresses PNGs by using more compressible colors that are close\n\
enough to the original color values (strength sets what is close enough).\n\
All input files are optimized in one batch on the GPU.  The output filename\n\
is the input name with \"-loss.png\" (or your --ext) unless the input is stdin,\n\
in which case the compressed image goes to stdout; the output path \"-\" with\n\
a single input file also writes to stdout.  Existing outputs are skipped\n\
unless --force is given.\n";

static const char *VERSION = "1.0.1-b200";

struct options {
    const char *extension, *output_file_path;
    char *const *files;
    unsigned long strength, bleed_divider;
    unsigned int num_files;
    bool using_stdin, using_stdout, force, skip_if_larger, strip, print_help, print_version, missing_arguments,
        verbose;
};

enum { arg_ext = 1000, arg_no_force, arg_skip_larger, arg_strip };

static const struct option long_options[] = {
    {"verbose", no_argument, NULL, 'v'},       {"quiet", no_argument, NULL, 'q'},
    {"force", no_argument, NULL, 'f'},         {"no-force", no_argument, NULL, arg_no_force},
    {"ext", required_argument, NULL, arg_ext}, {"skip-if-larger", no_argument, NULL, arg_skip_larger},
    {"output", required_argument, NULL, 'o'},  {"strip", no_argument, NULL, arg_strip},
    {"version", no_argument, NULL, 'V'},       {"help", no_argument, NULL, 'h'},
    {"strength", required_argument, NULL, 's'}, {"bleed", required_argument, NULL, 'b'},
    {NULL, 0, NULL, 0},
};

static bool parse_number(const char *arg, unsigned long *out) {
    char *end;
    unsigned long v = strtoul(arg, &end, 10);
    if (end == arg || *end) return false;
    *out = v;
    return true;
}

/* reference src/pngloss_opts.c:38-136 */
static pngloss_error parse_options(int argc, char *argv[], struct options *o) {
    int opt;
    while ((opt = getopt_long(argc, argv, "vqfo:Vhs:b:", long_options, NULL)) != -1) {
        switch (opt) {
        case 'v': o->verbose = true; break;
        case 'q': o->verbose = false; break;
        case 'f': o->force = true; break;
        case arg_no_force: o->force = false; break;
        case arg_ext: o->extension = optarg; break;
        case 'o':
            if (o->output_file_path) {
                fputs("--output option can be used only once\n", stderr);
                return INVALID_ARGUMENT;
            }
            if (strcmp(optarg, "-") == 0) o->using_stdout = true;
            else o->output_file_path = optarg;
            break;
        case arg_skip_larger: o->skip_if_larger = true; break;
        case arg_strip: o->strip = true; break;
        case 'h': o->print_help = true; break;
        case 'V': o->print_version = true; break;
        case 's':
            if (!parse_number(optarg, &o->strength)) {
                fputs("-s, --strength requires a numeric argument\n", stderr);
                return INVALID_ARGUMENT;
            }
            break;
        case 'b':
            if (!parse_number(optarg, &o->bleed_divider)) {
                fputs("-b, --bleed requires a numeric argument\n", stderr);
                return INVALID_ARGUMENT;
            }
            break;
        default: return INVALID_ARGUMENT;
        }
    }
    int argn = optind;
    if (argn < argc) {
        if (argn == argc - 1 && 0 == strcmp(argv[argn], "-")) {
            o->using_stdin = true;
            o->using_stdout = !o->output_file_path;
            argn = argc - 1;
        }
        o->num_files = (unsigned)(argc - argn);
        o->files = argv + argn;
    } else if (argn <= 1) {
        o->missing_arguments = true;
    }
    return SUCCESS;
}

static void print_full_version(FILE *fd) {
    fprintf(fd, "pngloss, %s, B200 build of the quantise + filter-search path.\n", VERSION);
    rwpng_version_info(fd);
    fprintf(fd, "   %d CUDA device(s) visible.\n\n", pngloss_b200_device_count());
}

/* ---- file name helpers (reference src/pngloss.c:305-373) ---------------------------------------------- */
static bool file_exists(const char *name) {
    FILE *f = fopen(name, "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

static char *add_filename_extension(const char *filename, const char *newext) {
    size_t x = strlen(filename);
    char *out = malloc(x + 4 + strlen(newext) + 1);
    if (!out) return NULL;
    strcpy(out, filename);
    if (x > 4 && (strncmp(out + x - 4, ".png", 4) == 0 || strncmp(out + x - 4, ".PNG", 4) == 0))
        strcpy(out + x - 4, newext);
    else
        strcpy(out + x, newext);
    return out;
}

static const char *filename_part(const char *path) {
    const char *slash = strrchr(path, '/');
    return slash ? slash + 1 : path;
}

/* ---- one job per input file -------------------------------------------------------------------------------- */
struct job {
    const char *filename;
    char *outname, *outname_free;
    png24_image input, output;
    unsigned char *row_filters;
    pngloss_error rc;
    bool loaded;
};

static pngloss_error read_image(const char *filename, bool using_stdin, png24_image *img, bool strip, bool verbose) {
    FILE *in = using_stdin ? stdin : fopen(filename, "rb");
    if (!in) {
        fprintf(stderr, "  error: cannot open %s for reading\n", filename);
        return READ_ERROR;
    }
    pngloss_error rc = rwpng_read_image24(in, img, strip, verbose);
    if (!using_stdin) fclose(in);
    if (rc) fprintf(stderr, "  error: cannot decode image %s\n", using_stdin ? "from stdin" : filename_part(filename));
    return rc;
}

/* reference src/pngloss.c:460-484 */
static pngloss_error prepare_output_image(const png24_image *in, png24_image *out) {
    memset(out, 0, sizeof *out);
    out->width = in->width;
    out->height = in->height;
    out->gamma = in->gamma;
    out->output_color = in->output_color;
    out->rgba_data = malloc((size_t)in->height * in->width * 4);
    out->row_pointers = malloc((size_t)in->height * sizeof out->row_pointers[0]);
    if (!out->rgba_data || !out->row_pointers) return OUT_OF_MEMORY_ERROR;
    for (size_t y = 0; y < in->height; y++) {
        out->row_pointers[y] = out->rgba_data + y * in->width * 4;
        memcpy(out->row_pointers[y], in->row_pointers[y], (size_t)in->width * 4);
    }
    return SUCCESS;
}

/* reference src/pngloss.c:375-431: temp file + rename so that a failed write never damages the target */
static pngloss_error write_image(png24_image *img, unsigned char *row_filters, const char *outname,
                                 const struct options *o) {
    FILE *out;
    char *tempname = NULL;
    if (o->using_stdout) {
        out = stdout;
        if (o->verbose) fprintf(stderr, "  writing compressed image to stdout\n");
    } else {
        tempname = malloc(strlen(outname) + 5);
        if (!tempname) return OUT_OF_MEMORY_ERROR;
        sprintf(tempname, "%s.tmp", outname);
        if (!(out = fopen(tempname, "wb"))) {
            fprintf(stderr, "  error: cannot open '%s' for writing\n", tempname);
            free(tempname);
            return CANT_WRITE_ERROR;
        }
        if (o->verbose) fprintf(stderr, "  writing compressed image as %s\n", filename_part(outname));
    }
    pngloss_error rc = rwpng_write_image24(out, img, row_filters);
    if (!o->using_stdout) {
        fclose(out);
        if (rc == SUCCESS && rename(tempname, outname) != 0) rc = CANT_WRITE_ERROR;
        if (rc) unlink(tempname);
    } else {
        fflush(out);
    }
    free(tempname);
    if (rc && rc != TOO_LARGE_FILE)
        fprintf(stderr, "  error: failed writing image to %s (%d)\n", o->using_stdout ? "stdout" : outname, rc);
    return rc;
}

int main(int argc, char *argv[]) {
    struct options o;
    memset(&o, 0, sizeof o);
    o.strength = 19;
    o.bleed_divider = 2;
    pngloss_error rc = parse_options(argc, argv, &o);
    if (rc != SUCCESS) return rc;

    if (o.print_version) { puts(VERSION); return SUCCESS; }
    if (o.missing_arguments) { print_full_version(stderr); fputs(USAGE, stderr); return MISSING_ARGUMENT; }
    if (o.print_help) { print_full_version(stdout); fputs(USAGE, stdout); return SUCCESS; }
    if (o.strength > 255) { fputs("Must specify a strength in the range 0-255.\n", stderr); return INVALID_ARGUMENT; }
    if (o.bleed_divider < 1 || o.bleed_divider > 32767) {
        fputs("Must specify a bleed divider in the range 1-32767.\n", stderr);
        return INVALID_ARGUMENT;
    }
    if (o.extension && o.output_file_path) {
        fputs("--ext and --output options can't be used at the same time\n", stderr);
        return INVALID_ARGUMENT;
    }
    if (!o.extension) o.extension = "-loss.png";
    if (o.output_file_path && o.num_files != 1) {
        fputs("  error: Only one input file is allowed when --output is used. This error also happens when "
              "filenames with spaces are not in quotes.\n", stderr);
        return INVALID_ARGUMENT;
    }
    if (o.using_stdout && !o.using_stdin && o.num_files != 1) {
        fputs("  error: Only one input file is allowed when using the special output path \"-\" to write to "
              "stdout. This error also happens when filenames with spaces are not in quotes.\n", stderr);
        return INVALID_ARGUMENT;
    }
    if (!o.num_files && !o.using_stdin) {
        fputs("No input files specified.\n", stderr);
        if (o.verbose) print_full_version(stderr);
        fputs(USAGE, stderr);
        return MISSING_ARGUMENT;
    }

    const unsigned n = o.num_files;
    struct job *jobs = calloc(n, sizeof *jobs);
    pngloss_b200_image *gpu_jobs = calloc(n, sizeof *gpu_jobs);
    unsigned *gpu_index = calloc(n, sizeof *gpu_index);
    if (!jobs || !gpu_jobs || !gpu_index) return OUT_OF_MEMORY_ERROR;

    /* ---- 1. names, overwrite check, decode (reference src/pngloss.c:173-259, per file) ---------------- */
    unsigned n_gpu = 0;
    for (unsigned i = 0; i < n; i++) {
        struct job *j = &jobs[i];
        j->filename = o.using_stdin ? "stdin" : o.files[i];
        j->outname = (char *)o.output_file_path;
        if (!o.using_stdout) {
            if (!j->outname) j->outname = j->outname_free = add_filename_extension(j->filename, o.extension);
            if (!o.force && file_exists(j->outname)) {
                fprintf(stderr, "  error: '%s' exists; not overwriting\n", j->outname);
                j->rc = NOT_OVERWRITING_ERROR;
                continue;
            }
        }
        if (o.verbose) fprintf(stderr, "%s:\n", j->filename);
        j->rc = read_image(j->filename, o.using_stdin, &j->input, o.strip, o.verbose);
        if (j->rc) continue;
        if (o.verbose) {
            fprintf(stderr, "  read %luKB file\n", (unsigned long)(j->input.file_size + 500UL) / 1000UL);
            if (j->input.input_color == RWPNG_SRGB) fprintf(stderr, "  passing sRGB tag from the input\n");
            else if (j->input.gamma != 0.45455)
                fprintf(stderr, "  converted image from gamma %2.1f to gamma 2.2\n", 1.0 / j->input.gamma);
        }
        j->rc = prepare_output_image(&j->input, &j->output);
        j->row_filters = malloc(j->input.height);   /* NULL is a valid value (reference :262-263) */
        if (j->rc) continue;
        j->loaded = true;
        pngloss_b200_image *g = &gpu_jobs[n_gpu];
        g->pixels = j->output.rgba_data;
        g->stride = (size_t)j->output.width * 4;
        g->width = j->output.width;
        g->height = j->output.height;
        g->row_filters = j->row_filters;
        gpu_index[n_gpu++] = i;
    }

    /* ---- 2. one batched call replaces the per-file optimize_with_rows (reference src/pngloss.c:266) ---- */
    if (n_gpu) {
        pngloss_b200_ctx *ctx = NULL;
        int dev = getenv("PNGLOSS_B200_DEVICE") ? atoi(getenv("PNGLOSS_B200_DEVICE")) : 0;
        int grc = pngloss_b200_ctx_create(&ctx, dev, NULL);
        if (grc == 0) {
            grc = pngloss_b200_optimize_batch(ctx, gpu_jobs, n_gpu, (unsigned)o.strength, (long)o.bleed_divider);
            if (grc && grc != PNGLOSS_B200_NO_ACCEPTABLE_ROW)
                fprintf(stderr, "  error: %s\n", pngloss_b200_ctx_error(ctx));
        } else {
            fprintf(stderr, "  error: no usable CUDA device %d (pngloss_b200 has no CPU fallback)\n", dev);
        }
        for (unsigned k = 0; k < n_gpu; k++) {
            struct job *j = &jobs[gpu_index[k]];
            const int st = grc && !gpu_jobs[k].status ? grc : gpu_jobs[k].status;
            if (st == PNGLOSS_B200_NO_ACCEPTABLE_ROW) {
                fprintf(stderr, "\naborting because no good row in %s\n", j->filename);   /* reference abort()s */
                abort();
            }
            if (st) j->rc = st == PNGLOSS_B200_OUT_OF_MEMORY ? OUT_OF_MEMORY_ERROR : PNGLOSS_DEVICE_ERROR;
            else if (o.verbose)
                fprintf(stderr, "%s:\n  compression complete (%u bytes per pixel)\n", j->filename,
                        gpu_jobs[k].bytes_per_pixel);
        }
        if (ctx) pngloss_b200_ctx_destroy(ctx);
    }

    /* ---- 3. encode (reference src/pngloss.c:268-300) -------------------------------------------------- */
    unsigned error_count = 0, skipped_count = 0;
    pngloss_error latest_error = SUCCESS;
    for (unsigned i = 0; i < n; i++) {
        struct job *j = &jobs[i];
        if (j->loaded && j->rc == SUCCESS) {
            if (o.skip_if_larger) j->output.maximum_file_size = j->input.file_size - 1;
            j->output.chunks = j->input.chunks;
            j->input.chunks = NULL;
            j->rc = write_image(&j->output, j->row_filters, j->outname, &o);
            if (o.verbose) {
                if (j->rc == SUCCESS) {
                    fprintf(stderr, "  wrote %luKB file (%.1f%% of original)\n",
                            ((unsigned long)j->output.file_size + 500UL) / 1000UL,
                            100.0f * (float)j->output.file_size / (float)j->input.file_size);
                    if (j->output.metadata_size > 0)
                        fprintf(stderr, "  copied %dKB of additional PNG metadata\n",
                                (int)(j->output.metadata_size + 500) / 1000);
                } else if (j->rc == TOO_LARGE_FILE) {
                    fprintf(stderr, "  file exceeded maximum size of %luKB\n",
                            ((unsigned long)j->output.maximum_file_size + 500UL) / 1000UL);
                }
            }
            /* on stdout an empty result would be nasty: send the original instead (reference :286-293) */
            if (o.using_stdout && j->rc == TOO_LARGE_FILE) {
                pngloss_error wrc = write_image(&j->input, NULL, j->outname, &o);
                if (wrc) j->rc = wrc;
            }
        }
        if (j->rc) {
            latest_error = j->rc;
            if (j->rc == TOO_LOW_QUALITY || j->rc == TOO_LARGE_FILE) skipped_count++;
            else error_count++;
        }
        rwpng_free_image24(&j->input);
        rwpng_free_image24(&j->output);
        free(j->row_filters);
        free(j->outname_free);
    }
    if (o.verbose) {
        if (error_count)
            fprintf(stderr, "There were errors compressing %d file%s out of a total of %d file%s.\n", error_count,
                    error_count == 1 ? "" : "s", n, n == 1 ? "" : "s");
        if (skipped_count)
            fprintf(stderr, "Skipped %d file%s out of a total of %d file%s.\n", skipped_count,
                    skipped_count == 1 ? "" : "s", n, n == 1 ? "" : "s");
        if (!skipped_count && !error_count) fprintf(stderr, "Compressed %d image%s.\n", n, n == 1 ? "" : "s");
    }
    free(jobs);
    free(gpu_jobs);
    free(gpu_index);
    return latest_error;
}
