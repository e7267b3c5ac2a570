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
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
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
  --gpus N          shard the files over N GPUs (default 1, 0 = all visible) [new]\n\
  --jobs N          CPU threads for PNG decoding / encoding (default: all cores) [new]\n\
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
    unsigned long gpus, cpu_jobs;   /* additions of this build; everything else is the reference's */
    bool using_stdin, using_stdout, force, skip_if_larger, strip, print_help, print_version, missing_arguments,
        verbose;
};

enum { arg_ext = 1000, arg_no_force, arg_skip_larger, arg_strip, arg_gpus, arg_jobs };

static const struct option long_options[] = {
    {"verbose", no_argument, NULL, 'v'},       {"quiet", no_argument, NULL, 'q'},
    {"force", no_argument, NULL, 'f'},         {"no-force", no_argument, NULL, arg_no_force},
    {"ext", required_argument, NULL, arg_ext}, {"skip-if-larger", no_argument, NULL, arg_skip_larger},
    {"output", required_argument, NULL, 'o'},  {"strip", no_argument, NULL, arg_strip},
    {"version", no_argument, NULL, 'V'},       {"help", no_argument, NULL, 'h'},
    {"strength", required_argument, NULL, 's'}, {"bleed", required_argument, NULL, 'b'},
    {"gpus", required_argument, NULL, arg_gpus}, {"jobs", required_argument, NULL, arg_jobs},
    {NULL, 0, NULL, 0},
};

/* What an option does when it is seen.  The option letters, long names and messages are the reference's
 * command line surface (src/pngloss_opts.c:22-47); the parser itself is a table walk. */
enum action { SET_TRUE, SET_FALSE, TAKE_STRING, TAKE_NUMBER, TAKE_OUTPUT };
struct option_rule {
    int id;
    enum action action;
    size_t field;              /* offset of the struct options member it touches */
    const char *number_error;  /* TAKE_NUMBER: message for a non-numeric argument */
};
#define FIELD(m) offsetof(struct options, m)
static const struct option_rule rules[] = {
    {'v', SET_TRUE, FIELD(verbose), NULL},
    {'q', SET_FALSE, FIELD(verbose), NULL},
    {'f', SET_TRUE, FIELD(force), NULL},
    {arg_no_force, SET_FALSE, FIELD(force), NULL},
    {arg_skip_larger, SET_TRUE, FIELD(skip_if_larger), NULL},
    {arg_strip, SET_TRUE, FIELD(strip), NULL},
    {'h', SET_TRUE, FIELD(print_help), NULL},
    {'V', SET_TRUE, FIELD(print_version), NULL},
    {arg_ext, TAKE_STRING, FIELD(extension), NULL},
    {'o', TAKE_OUTPUT, FIELD(output_file_path), NULL},
    {'s', TAKE_NUMBER, FIELD(strength), "-s, --strength requires a numeric argument\n"},
    {'b', TAKE_NUMBER, FIELD(bleed_divider), "-b, --bleed requires a numeric argument\n"},
    {arg_gpus, TAKE_NUMBER, FIELD(gpus), "--gpus requires a numeric argument\n"},
    {arg_jobs, TAKE_NUMBER, FIELD(cpu_jobs), "--jobs requires a positive numeric argument\n"},
};

static bool whole_number(const char *text, unsigned long *value) {
    if (!*text) return false;
    char *rest = NULL;
    const unsigned long v = strtoul(text, &rest, 10);
    if (*rest) return false;
    *value = v;
    return true;
}

/* Behaviour of reference src/pngloss_opts.c:38-136: options in any order; "-o -" means stdout; a lone trailing
 * "-" means stdin (and stdout, unless -o named a file); no arguments at all is "missing arguments". */
static pngloss_error parse_options(int argc, char *argv[], struct options *o) {
    for (int id; (id = getopt_long(argc, argv, "vqfo:Vhs:b:", long_options, NULL)) != -1;) {
        const struct option_rule *rule = NULL;
        for (size_t k = 0; k < sizeof rules / sizeof rules[0] && !rule; k++)
            if (rules[k].id == id) rule = &rules[k];
        if (!rule) return INVALID_ARGUMENT;            /* getopt already complained */
        void *member = (char *)o + rule->field;
        switch (rule->action) {
        case SET_TRUE: *(bool *)member = true; break;
        case SET_FALSE: *(bool *)member = false; break;
        case TAKE_STRING: *(const char **)member = optarg; break;
        case TAKE_NUMBER:
            if (!whole_number(optarg, (unsigned long *)member) || (id == arg_jobs && !o->cpu_jobs)) {
                fputs(rule->number_error, stderr);
                return INVALID_ARGUMENT;
            }
            break;
        case TAKE_OUTPUT:
            if (o->output_file_path) {
                fputs("--output option can be used only once\n", stderr);
                return INVALID_ARGUMENT;
            }
            if (optarg[0] == '-' && !optarg[1]) o->using_stdout = true;
            else o->output_file_path = optarg;
            break;
        }
    }
    const int positional = argc - optind;
    if (positional <= 0) {
        o->missing_arguments = optind <= 1;            /* bare "pngloss" */
        return SUCCESS;
    }
    o->files = argv + optind;
    o->num_files = (unsigned)positional;
    if (positional == 1 && strcmp(argv[optind], "-") == 0) {
        o->using_stdin = true;
        o->using_stdout = o->output_file_path == NULL;
    }
    return SUCCESS;
}

static void print_full_version(FILE *fd) {
    fprintf(fd, "pngloss, %s, B200 build of the quantise + filter-search path.\n", VERSION);
    rwpng_version_info(fd);
    fprintf(fd, "   %d CUDA device(s) visible.\n\n", pngloss_b200_device_count());
}

/* ---- file name helpers (behaviour of reference src/pngloss.c:305-373) ------------------------------------ */
static bool file_exists(const char *name) { return access(name, F_OK) == 0; }

/* "x.png" / "x.PNG" -> "x" + newext; any other name gets newext appended */
static char *add_filename_extension(const char *filename, const char *newext) {
    size_t stem = strlen(filename);
    if (stem > 4) {
        const char *suffix = filename + stem - 4;
        if (!strcmp(suffix, ".png") || !strcmp(suffix, ".PNG")) stem -= 4;
    }
    const size_t total = stem + strlen(newext) + 1;
    char *out = malloc(total);
    if (out) snprintf(out, total, "%.*s%s", (int)stem, filename, newext);
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
    /* the result as filtered scanlines from the GPU (K4); NULL with PNGLOSS_CPU_FILTER=1, then the writer narrows
     * and filters the quantised pixels itself like the reference's libpng does */
    unsigned char *scanlines;
    unsigned scan_bpp;
    pngloss_error rc;
    bool loaded;
    char *log;          /* this file's messages, printed in file order once a phase is over */
    size_t log_len;
    FILE *logf;
};

struct run {
    const struct options *o;
    struct job *jobs;
    unsigned n;
};

static pngloss_error read_image(const char *filename, bool using_stdin, png24_image *img, bool strip, bool verbose,
                                FILE *log) {
    FILE *in = using_stdin ? stdin : fopen(filename, "rb");
    if (!in) {
        fprintf(log, "  error: cannot open %s for reading\n", filename);
        return READ_ERROR;
    }
    pngloss_error rc = rwpng_read_image24(in, img, strip, verbose);
    if (!using_stdin) fclose(in);
    if (rc) fprintf(log, "  error: cannot decode image %s\n", using_stdin ? "from stdin" : filename_part(filename));
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
static pngloss_error write_image(png24_image *img, unsigned char *row_filters, const unsigned char *scanlines,
                                 unsigned scan_bpp, const char *outname, const struct options *o, FILE *log) {
    FILE *out;
    char *tempname = NULL;
    char *membuf = NULL;      /* stdout: the file is built in memory and only a complete, accepted one is sent - */
    size_t memlen = 0;        /* a writer that stops at the --skip-if-larger cap must not leave half a PNG there */
    if (o->using_stdout) {
        out = open_memstream(&membuf, &memlen);
        if (!out) return OUT_OF_MEMORY_ERROR;
        if (o->verbose) fprintf(log, "  writing compressed image to stdout\n");
    } else {
        tempname = malloc(strlen(outname) + 5);
        if (!tempname) return OUT_OF_MEMORY_ERROR;
        sprintf(tempname, "%s.tmp", outname);
        if (!(out = fopen(tempname, "wb"))) {
            fprintf(log, "  error: cannot open '%s' for writing\n", tempname);
            free(tempname);
            return CANT_WRITE_ERROR;
        }
        if (o->verbose) fprintf(log, "  writing compressed image as %s\n", filename_part(outname));
    }
    pngloss_error rc = scanlines ? rwpng_write_scanlines(out, img, scan_bpp, scanlines)
                                 : rwpng_write_image24(out, img, row_filters);
    if (!o->using_stdout) {
        fclose(out);
        if (rc == SUCCESS && rename(tempname, outname) != 0) rc = CANT_WRITE_ERROR;
        if (rc) unlink(tempname);
    } else {
        fclose(out);
        if (rc == SUCCESS && (fwrite(membuf, 1, memlen, stdout) != memlen || fflush(stdout) != 0)) rc = CANT_WRITE_ERROR;
        free(membuf);
    }
    free(tempname);
    if (rc && rc != TOO_LARGE_FILE)
        fprintf(log, "  error: failed writing image to %s (%d)\n", o->using_stdout ? "stdout" : outname, rc);
    return rc;
}

/* ---- a small parallel-for over the files (decode and encode are independent per file) ------------------- */
struct pfor {
    void (*fn)(struct run *, unsigned);
    struct run *r;
    atomic_uint next;
};

static void *pfor_worker(void *arg) {
    struct pfor *p = arg;
    for (;;) {
        unsigned i = atomic_fetch_add(&p->next, 1);
        if (i >= p->r->n) break;
        p->fn(p->r, i);
    }
    return NULL;
}

static void parallel_for(struct run *r, void (*fn)(struct run *, unsigned)) {
    unsigned threads = (unsigned)r->o->cpu_jobs;
    if (!threads) {
        long cores = sysconf(_SC_NPROCESSORS_ONLN);
        threads = cores > 0 ? (unsigned)cores : 1;
    }
    if (threads > r->n) threads = r->n;
    struct pfor p = {fn, r, 0};
    if (threads <= 1 || r->o->using_stdin) {
        pfor_worker(&p);
        return;
    }
    pthread_t *tid = calloc(threads, sizeof *tid);
    unsigned started = 0;
    for (; tid && started < threads; started++)
        if (pthread_create(&tid[started], NULL, pfor_worker, &p)) break;
    if (!started) pfor_worker(&p);
    for (unsigned k = 0; k < started; k++) pthread_join(tid[k], NULL);
    free(tid);
}

static void flush_logs(struct run *r) {
    for (unsigned i = 0; i < r->n; i++) {
        struct job *j = &r->jobs[i];
        if (j->logf) {
            fclose(j->logf);
            j->logf = NULL;
        }
        if (j->log && j->log_len) fwrite(j->log, 1, j->log_len, stderr);
        free(j->log);
        j->log = NULL;
        j->log_len = 0;
        j->logf = open_memstream(&j->log, &j->log_len);
        if (!j->logf) j->logf = stderr;
    }
}

/* phase 1 for one file: name, overwrite check, decode (reference src/pngloss.c:173-259) */
static void decode_job(struct run *r, unsigned i) {
    const struct options *o = r->o;
    struct job *j = &r->jobs[i];
    j->outname = (char *)o->output_file_path;
    if (!o->using_stdout) {
        if (!j->outname) j->outname = j->outname_free = add_filename_extension(j->filename, o->extension);
        if (!o->force && file_exists(j->outname)) {
            fprintf(j->logf, "  error: '%s' exists; not overwriting\n", j->outname);
            j->rc = NOT_OVERWRITING_ERROR;
            return;
        }
    }
    if (o->verbose) fprintf(j->logf, "%s:\n", j->filename);
    j->rc = read_image(j->filename, o->using_stdin, &j->input, o->strip, o->verbose, j->logf);
    if (j->rc) return;
    if (o->verbose) {
        fprintf(j->logf, "  read %luKB file\n", (unsigned long)(j->input.file_size + 500UL) / 1000UL);
        if (j->input.input_color == RWPNG_SRGB) fprintf(j->logf, "  passing sRGB tag from the input\n");
        else if (j->input.gamma != 0.45455)
            fprintf(j->logf, "  converted image from gamma %2.1f to gamma 2.2\n", 1.0 / j->input.gamma);
    }
    j->rc = prepare_output_image(&j->input, &j->output);
    if (!o->using_stdout) {   /* the decoded input is only needed again for the stdout fall-back of --skip-if-larger */
        free(j->input.rgba_data);
        free(j->input.row_pointers);
        j->input.rgba_data = NULL;
        j->input.row_pointers = NULL;
    }
    j->row_filters = malloc(j->input.height);   /* NULL is a valid value (reference :262-263) */
    if (!getenv("PNGLOSS_CPU_FILTER"))          /* NULL is a valid value here too: CPU filtering */
        j->scanlines = malloc((size_t)j->input.height * (1 + 4 * (size_t)j->input.width));
    if (!j->rc) j->loaded = true;
}

/* phase 3 for one file: encode (reference src/pngloss.c:268-300) */
static void encode_job(struct run *r, unsigned i) {
    const struct options *o = r->o;
    struct job *j = &r->jobs[i];
    if (!j->loaded || j->rc != SUCCESS) return;
    if (o->skip_if_larger) j->output.maximum_file_size = j->input.file_size - 1;
    j->output.chunks = j->input.chunks;
    j->input.chunks = NULL;
    j->rc = write_image(&j->output, j->row_filters, j->scanlines, j->scan_bpp, j->outname, o, j->logf);
    if (o->verbose) {
        if (j->rc == SUCCESS) {
            fprintf(j->logf, "  wrote %luKB file (%.1f%% of original)\n",
                    ((unsigned long)j->output.file_size + 500UL) / 1000UL,
                    100.0f * (float)j->output.file_size / (float)j->input.file_size);
            if (j->output.metadata_size > 0)
                fprintf(j->logf, "  copied %dKB of additional PNG metadata\n",
                        (int)(j->output.metadata_size + 500) / 1000);
        } else if (j->rc == TOO_LARGE_FILE) {
            fprintf(j->logf, "  file exceeded maximum size of %luKB\n",
                    ((unsigned long)j->output.maximum_file_size + 500UL) / 1000UL);
        }
    }
    /* on stdout an empty result would be nasty: send the original instead (reference :286-293) */
    if (o->using_stdout && j->rc == TOO_LARGE_FILE) {
        pngloss_error wrc = write_image(&j->input, NULL, NULL, 0, j->outname, o, j->logf);
        if (wrc) j->rc = wrc;
    }
}

/* ---- phase 2: the GPU(s).  One host thread and one context per GPU, images assigned longest first ------ */
struct gpu_shard {
    int device;
    pngloss_b200_ctx *ctx;             /* created by run_gpus (all contexts share one NCCL communicator) */
    int reduce;                        /* sum the symbol histogram over the GPUs (NCCL) */
    uint64_t hist[256];                /* symbol counts of the whole batch (of this GPU's share without NCCL) */
    int hist_ok;
    pngloss_b200_image *images;
    unsigned *job_index;
    unsigned n;
    unsigned long long pixels;
    unsigned strength;
    long bleed;
    int rc;
    char err[256];
};

static void *gpu_worker(void *arg) {
    struct gpu_shard *s = arg;
    pngloss_b200_ctx *ctx = s->ctx;
    if (!ctx) {
        s->rc = PNGLOSS_B200_DEVICE_ERROR;
        snprintf(s->err, sizeof s->err, "no usable CUDA device %d (pngloss_b200 has no CPU fallback)", s->device);
        return NULL;
    }
    if (!s->n) return NULL;
    s->rc = pngloss_b200_optimize_batch(ctx, s->images, s->n, s->strength, s->bleed);
    if (s->rc && s->rc != PNGLOSS_B200_NO_ACCEPTABLE_ROW) snprintf(s->err, sizeof s->err, "%s", pngloss_b200_ctx_error(ctx));
    return NULL;
}

static int by_pixels_desc(const void *a, const void *b) {
    const pngloss_b200_image *x = a, *y = b;
    const unsigned long long px = (unsigned long long)x->width * x->height, py = (unsigned long long)y->width * y->height;
    return px < py ? 1 : px > py ? -1 : 0;
}

/* The GPUs of a run: one context per GPU, created with the first chunk of files and kept until the end; with
 * several GPUs the contexts share one NCCL communicator. */
static struct {
    pngloss_b200_ctx *ctx[64];
    unsigned n;
    int nccl, ready;
    unsigned images;
} gpus;

static void open_gpus(const struct options *o, unsigned n_images) {
    if (gpus.ready) return;
    gpus.ready = 1;
    int visible = pngloss_b200_device_count();
    int first_dev = getenv("PNGLOSS_B200_DEVICE") ? atoi(getenv("PNGLOSS_B200_DEVICE")) : 0;
    unsigned ngpu = o->gpus ? (unsigned)o->gpus : (visible > 0 ? (unsigned)visible : 1);
    if (visible > 0 && ngpu > (unsigned)visible) ngpu = (unsigned)visible;
    if (ngpu > n_images) ngpu = n_images;
    if (ngpu > 64) ngpu = 64;
    if (!ngpu) ngpu = 1;
    gpus.n = ngpu;
    int all = 1;
    for (unsigned g = 0; g < ngpu; g++) {
        if (pngloss_b200_ctx_create(&gpus.ctx[g], first_dev + (int)g, NULL)) gpus.ctx[g] = NULL;
        all &= gpus.ctx[g] != NULL;
    }
    if (ngpu > 1 && all) {
        gpus.nccl = pngloss_b200_comm_init_all(gpus.ctx, (int)ngpu) == 0;
        if (!gpus.nccl && o->verbose)
            fprintf(stderr, "  note: no NCCL communicator (%s); symbol counts are added up on the host\n",
                    pngloss_b200_ctx_error(gpus.ctx[0]));
    }
}

/* Batch-level form of the reference's per-image "used N unique symbols" (src/pngloss_image.c:315-325): the
 * symbol histogram of everything the run quantised, summed over the GPUs by the library's NCCL all-reduce. */
static void *histogram_worker(void *arg) {
    struct gpu_shard *s = arg;
    s->hist_ok = s->ctx && pngloss_b200_ctx_symbol_histogram(s->ctx, s->hist, s->reduce) == 0;
    return NULL;
}

static void close_gpus(const struct options *o) {
    if (!gpus.ready) return;
    if (o->verbose && gpus.images) {
        struct gpu_shard *sh = calloc(gpus.n, sizeof *sh);
        pthread_t *tid = calloc(gpus.n, sizeof *tid);
        for (unsigned g = 0; sh && tid && g < gpus.n; g++) {
            sh[g].ctx = gpus.ctx[g];
            sh[g].reduce = gpus.nccl;
            if (gpus.n == 1 || pthread_create(&tid[g], NULL, histogram_worker, &sh[g])) {
                histogram_worker(&sh[g]);
                tid[g] = 0;
            }
        }
        uint64_t total[256] = {0}, bytes = 0;
        unsigned unique = 0;
        int ok = sh && tid;
        for (unsigned g = 0; ok && g < gpus.n; g++)
            if (gpus.n > 1 && tid[g]) pthread_join(tid[g], NULL);
        for (unsigned g = 0; ok && g < (gpus.nccl ? 1u : gpus.n); g++) {
            ok &= sh[g].hist_ok;
            for (int k = 0; k < 256; k++) total[k] += sh[g].hist[k];
        }
        for (int k = 0; k < 256; k++) {
            unique += total[k] != 0;
            bytes += total[k];
        }
        if (ok)
            fprintf(stderr, "batch of %u image%s on %u GPU%s: used %u unique symbols in %llu bytes%s\n", gpus.images,
                    gpus.images == 1 ? "" : "s", gpus.n, gpus.n == 1 ? "" : "s", unique, (unsigned long long)bytes,
                    gpus.nccl ? " (histogram summed over the GPUs by NCCL)" : "");
        free(sh);
        free(tid);
    }
    for (unsigned g = 0; g < gpus.n; g++) pngloss_b200_ctx_destroy(gpus.ctx[g]);
    gpus.ready = 0;
}

static void run_gpus(struct run *r) {
    const struct options *o = r->o;
    unsigned n_loaded = 0;
    for (unsigned i = 0; i < r->n; i++) n_loaded += r->jobs[i].loaded;
    if (!n_loaded) return;
    open_gpus(o, n_loaded);
    gpus.images += n_loaded;
    const unsigned ngpu = gpus.n;

    /* all loaded images, largest first, each to the GPU with the fewest pixels so far */
    struct tagged { pngloss_b200_image im; unsigned job; } *all = calloc(n_loaded, sizeof *all);
    struct gpu_shard *shards = calloc(ngpu, sizeof *shards);
    unsigned k = 0;
    for (unsigned i = 0; i < r->n; i++) {
        struct job *j = &r->jobs[i];
        if (!j->loaded) continue;
        all[k].im.pixels = j->output.rgba_data;
        all[k].im.stride = (size_t)j->output.width * 4;
        all[k].im.width = j->output.width;
        all[k].im.height = j->output.height;
        all[k].im.row_filters = j->row_filters;
        all[k].im.scanlines = j->scanlines;
        all[k].im.flags = j->scanlines ? PNGLOSS_B200_IMAGE_NO_PIXELS : 0;   /* the encoder needs nothing else */
        all[k].job = i;
        k++;
    }
    qsort(all, n_loaded, sizeof *all, by_pixels_desc);   /* struct starts with the image: same comparator */
    for (unsigned g = 0; g < ngpu; g++) {
        shards[g].device = (int)g;
        shards[g].ctx = gpus.ctx[g];
        shards[g].images = calloc(n_loaded, sizeof(pngloss_b200_image));
        shards[g].job_index = calloc(n_loaded, sizeof(unsigned));
        shards[g].strength = (unsigned)o->strength;
        shards[g].bleed = (long)o->bleed_divider;
    }
    for (unsigned i = 0; i < n_loaded; i++) {
        unsigned best = 0;
        for (unsigned g = 1; g < ngpu; g++)
            if (shards[g].pixels < shards[best].pixels) best = g;
        struct gpu_shard *s = &shards[best];
        s->images[s->n] = all[i].im;
        s->job_index[s->n++] = all[i].job;
        s->pixels += (unsigned long long)all[i].im.width * all[i].im.height;
    }
    pthread_t *tid = calloc(ngpu, sizeof *tid);
    for (unsigned g = 0; g < ngpu; g++)
        if (ngpu == 1 || pthread_create(&tid[g], NULL, gpu_worker, &shards[g])) {
            gpu_worker(&shards[g]);
            tid[g] = 0;
        }
    for (unsigned g = 0; g < ngpu; g++)
        if (ngpu > 1 && tid[g]) pthread_join(tid[g], NULL);
    for (unsigned g = 0; g < ngpu; g++) {
        struct gpu_shard *s = &shards[g];
        if (s->rc && s->err[0]) fprintf(stderr, "  error: %s\n", s->err);
        for (unsigned q = 0; q < s->n; q++) {
            struct job *j = &r->jobs[s->job_index[q]];
            const int st = s->rc && !s->images[q].status ? s->rc : s->images[q].status;
            if (st == PNGLOSS_B200_NO_ACCEPTABLE_ROW) {
                fprintf(stderr, "\naborting because no good row in %s\n", j->filename);   /* reference abort()s */
                abort();
            }
            j->scan_bpp = s->images[q].scan_bytes_per_pixel;
            if (st) j->rc = st == PNGLOSS_B200_OUT_OF_MEMORY ? OUT_OF_MEMORY_ERROR : PNGLOSS_DEVICE_ERROR;
            else if (o->verbose)
                fprintf(j->logf, "%s:\n  compression complete (%u bytes per pixel, GPU %d)\n", j->filename,
                        s->images[q].bytes_per_pixel, s->device);
        }
        free(s->images);
        free(s->job_index);
    }
    free(tid);
    free(shards);
    free(all);
}

/* width * height of a PNG from its IHDR (0 if the file cannot be read: it will fail in decode_job with a message) */
static unsigned long long decoded_bytes_estimate(const char *filename) {
    unsigned char hdr[24];
    FILE *f = fopen(filename, "rb");
    if (!f) return 0;
    const size_t got = fread(hdr, 1, sizeof hdr, f);
    fclose(f);
    if (got != sizeof hdr || memcmp(hdr + 12, "IHDR", 4)) return 0;
    const unsigned long long w = ((unsigned long long)hdr[16] << 24) | (hdr[17] << 16) | (hdr[18] << 8) | hdr[19];
    const unsigned long long h = ((unsigned long long)hdr[20] << 24) | (hdr[21] << 16) | (hdr[22] << 8) | hdr[23];
    return w * h * 13ull + h;
}

static unsigned long long host_budget_bytes(void) {
    const char *env = getenv("PNGLOSS_HOST_BUDGET_MB");
    if (env && atoll(env) > 0) return (unsigned long long)atoll(env) << 20;
    unsigned long long avail_kb = 0;
    FILE *f = fopen("/proc/meminfo", "r");
    if (f) {
        char line[128];
        while (fgets(line, sizeof line, f))
            if (sscanf(line, "MemAvailable: %llu kB", &avail_kb) == 1) break;
        fclose(f);
    }
    return avail_kb ? avail_kb * 1024ull / 2 : 8ull << 30;
}

int main(int argc, char *argv[]) {
    struct options o;
    memset(&o, 0, sizeof o);
    o.strength = 19;
    o.bleed_divider = 2;
    o.gpus = 1;
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

    struct run r = {&o, calloc(o.num_files, sizeof(struct job)), o.num_files};
    if (!r.jobs) return OUT_OF_MEMORY_ERROR;
    for (unsigned i = 0; i < r.n; i++) r.jobs[i].filename = o.using_stdin ? "stdin" : o.files[i];

    /* The reference handles one file at a time (src/pngloss.c:173-205); the GPU wants many.  Files are taken in
     * chunks whose decoded size (input + output pixels + scanlines, about 13 bytes per pixel) fits a host-memory
     * budget - PNGLOSS_HOST_BUDGET_MB, default half of the available memory - so that a directory of thousands
     * of large PNGs neither exhausts the host nor starves the GPU: decode chunk, one batched call per GPU, encode,
     * free, next chunk. */
    const unsigned long long budget = host_budget_bytes();
    for (unsigned start = 0; start < r.n;) {
        unsigned end = start;
        unsigned long long need = 0;
        while (end < r.n) {
            const unsigned long long est = o.using_stdin ? 0 : decoded_bytes_estimate(r.jobs[end].filename);
            if (end > start && need + est > budget) break;
            need += est;
            end++;
        }
        struct run chunk = {&o, r.jobs + start, end - start};
        flush_logs(&chunk);                    /* opens the per-file message buffers */
        parallel_for(&chunk, decode_job);      /* 1. decode the chunk's inputs (CPU threads) */
        flush_logs(&chunk);
        run_gpus(&chunk);                      /* 2. one batched call per GPU replaces the per-file
                                                     optimize_with_rows (reference src/pngloss.c:266) */
        flush_logs(&chunk);
        parallel_for(&chunk, encode_job);      /* 3. encode the chunk's outputs (CPU threads) */
        flush_logs(&chunk);
        for (unsigned i = start; i < end; i++) {   /* pixels go now; results and messages stay for the summary */
            struct job *j = &r.jobs[i];
            rwpng_free_image24(&j->input);
            rwpng_free_image24(&j->output);
            free(j->row_filters);
            free(j->scanlines);
            j->row_filters = j->scanlines = NULL;
        }
        start = end;
    }
    close_gpus(&o);

    unsigned error_count = 0, skipped_count = 0;
    pngloss_error latest_error = SUCCESS;
    for (unsigned i = 0; i < r.n; i++) {
        struct job *j = &r.jobs[i];
        if (j->rc) {
            latest_error = j->rc;
            if (j->rc == TOO_LOW_QUALITY || j->rc == TOO_LARGE_FILE) skipped_count++;
            else error_count++;
        }
        free(j->outname_free);
        if (j->logf && j->logf != stderr) fclose(j->logf);
        free(j->log);
    }
    const unsigned n = r.n;
    if (o.verbose) {
        if (error_count)
            fprintf(stderr, "There were errors compressing %d file%s out of a total of %d file%s.\n", error_count,
                    error_count == 1 ? "" : "s", n, n == 1 ? "" : "s");
        if (skipped_count)
            fprintf(stderr, "Skipped %d file%s out of a total of %d file%s.\n", skipped_count,
                    skipped_count == 1 ? "" : "s", n, n == 1 ? "" : "s");
        if (!skipped_count && !error_count) fprintf(stderr, "Compressed %d image%s.\n", n, n == 1 ? "" : "s");
    }
    free(r.jobs);
    return latest_error;
}
