// Host shim: the C ABI of include/pngloss_b200.h on top of the kernels in pl_kernels.cuh.
//
// There is deliberately no CPU implementation of the path in this file or anywhere in the library:
// when CUDA is unavailable every entry point fails (PNGLOSS_B200_DEVICE_ERROR).
#include "../../include/pngloss_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <new>
#include <vector>

#include "pl_kernels.cuh"

struct pngloss_b200_job;
struct pngloss_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int lpc = 0;
    int bm = -1;   // bucket-maxima variant of K2: -1 choose from the strength, 0 off, 1 on
    int sm_count = 0;
    int lean = -1; // lean kernel (pl_k2_lean) where it applies: -1 yes (default), 0 never, 1 yes
    int solo = -1; // latency kernel (pl_k2_solo) for one-image-per-CTA grids: -1 where it applies (default), 0 never,
                   // 1 one chain warp, 2 five chain warps
    char err[512] = {0};
    // job API (pngloss_b200_submit / _wait): copy streams, the device batches it recycles, jobs in flight
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<pngloss_b200_batch *> pool;
    std::vector<pngloss_b200_job *> inflight;   // in submission order
    size_t max_jobs = 2;                        // device batches the job API keeps (pngloss_b200_ctx_set_pipeline)
    bool job_streams = false;                   // every pooled batch computes on a stream of its own
    // multi-GPU (pl_comm.cuh): NCCL communicator of this context's device, scratch for small reductions
    void *comm = nullptr;
    int comm_ranks = 0, comm_rank = 0;
    unsigned long long *comm_scratch = nullptr;
    unsigned long long symbols[256] = {0};      // symbol counts of every image the host-buffer calls finished
    unsigned char *scrub = nullptr;             // pngloss_b200_ctx_flush_l2
    size_t scrub_bytes = 0;
};

struct pngloss_b200_batch {
    pngloss_b200_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;       // the context's stream, or the batch's own (job pipeline)
    bool own_stream = false;
    bool pooled = false;                 // owned by the job API's pool
    size_t n = 0;
    std::vector<uint32_t> w, h;
    std::vector<PlImageDev> himgs;
    std::vector<int> order;              // image indices sorted by (w, h): equal sizes share a CTA
    unsigned char *slab = nullptr;
    size_t slab_bytes = 0;
    // views into the slab
    PlImageDev *dimgs = nullptr;
    int *dslots = nullptr;
    unsigned char *zero_begin = nullptr;  // region cleared before every run
    size_t zero_bytes = 0;
    uint32_t *chan_hist = nullptr, *flags = nullptr, *status = nullptr, *final_hist = nullptr;
    unsigned long long *batch_hist = nullptr;
    std::vector<int> hslots;
    uint32_t *hstatus = nullptr;         // pinned: [n][4] status words, copied back asynchronously
    unsigned long long *hbhist = nullptr; // pinned (tail of the same allocation): the batch histogram, job API
    // row filters of all images, contiguous on the device so that the job API fetches them with one copy
    unsigned char *dfilters = nullptr, *hfilters = nullptr;   // hfilters: pinned staging, allocated on demand
    std::vector<size_t> filt_off;
    size_t filt_bytes = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ran = false;
    uint32_t info[4] = {0, 0, 0, 0};
    bool desc_dirty = true;
    bool in_place = false;               // PNGLOSS_B200_BATCH_IN_PLACE: out aliases in
    // K4 (filtered scanlines): second slab, allocated by the first pngloss_b200_batch_scanlines call
    unsigned char *scan_slab = nullptr;
    PlScanDev *dscan = nullptr;
    uint32_t *oflags = nullptr, *hoflags = nullptr;   // [n][4]; hoflags pinned
    std::vector<size_t> scan_off;
    cudaEvent_t ev_scan[3] = {nullptr, nullptr, nullptr};
    bool scan_ran = false;
    // job API
    bool busy = false;
    cudaEvent_t ev_up = nullptr, ev_done = nullptr, ev_down = nullptr;
};

struct pngloss_b200_job {
    pngloss_b200_ctx *ctx = nullptr;
    pngloss_b200_batch *batch = nullptr;
    pngloss_b200_image *images = nullptr;
    size_t n = 0;
    bool finalized = false;
    bool scanlines = false;   // some image asked for filtered scanlines (K4 ran)
    int rc = 0;
};

static int set_err(pngloss_b200_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define PL_CUDA(ctx, call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return set_err((ctx), e_ == cudaErrorMemoryAllocation ? PNGLOSS_B200_OUT_OF_MEMORY \
                                                                  : PNGLOSS_B200_DEVICE_ERROR, \
                           "%s failed: %s", #call, cudaGetErrorString(e_));                    \
    } while (0)

extern "C" int pngloss_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int pngloss_b200_ctx_create(pngloss_b200_ctx **out, int device, void *cuda_stream) {
    if (!out) return PNGLOSS_B200_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return PNGLOSS_B200_DEVICE_ERROR;
    pngloss_b200_ctx *ctx = new (std::nothrow) pngloss_b200_ctx();
    if (!ctx) return PNGLOSS_B200_OUT_OF_MEMORY;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PNGLOSS_B200_DEVICE_ERROR; }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return PNGLOSS_B200_DEVICE_ERROR;
        }
        ctx->own_stream = true;
    }
    if (cudaEventCreate(&ctx->t0) != cudaSuccess || cudaEventCreate(&ctx->t1) != cudaSuccess) {
        delete ctx;
        return PNGLOSS_B200_DEVICE_ERROR;
    }
    *out = ctx;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" void pngloss_b200_comm_destroy(pngloss_b200_ctx *ctx);

extern "C" void pngloss_b200_ctx_destroy(pngloss_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    pngloss_b200_comm_destroy(ctx);
    if (ctx->comm_scratch) cudaFree(ctx->comm_scratch);
    if (ctx->scrub) cudaFree(ctx->scrub);
    while (!ctx->inflight.empty()) pngloss_b200_wait(ctx->inflight.front());
    for (pngloss_b200_batch *b : ctx->pool) pngloss_b200_batch_destroy(b);
    ctx->pool.clear();
    if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *pngloss_b200_ctx_error(const pngloss_b200_ctx *ctx) { return ctx ? ctx->err : ""; }

extern "C" int pngloss_b200_ctx_set_lanes(pngloss_b200_ctx *ctx, int lpc) {
    if (!ctx || !(lpc == 0 || lpc == 1 || lpc == 2 || lpc == 4 || lpc == 8))
        return PNGLOSS_B200_INVALID_ARGUMENT;
    ctx->lpc = lpc;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_set_lean(pngloss_b200_ctx *ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1) return PNGLOSS_B200_INVALID_ARGUMENT;
    ctx->lean = mode;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_set_solo(pngloss_b200_ctx *ctx, int mode) {
    if (!ctx || mode < -1 || mode > 3) return PNGLOSS_B200_INVALID_ARGUMENT;
    ctx->solo = mode;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_set_pipeline(pngloss_b200_ctx *ctx, int jobs_in_flight) {
    if (!ctx || jobs_in_flight < 1 || jobs_in_flight > 64) return PNGLOSS_B200_INVALID_ARGUMENT;
    ctx->max_jobs = (size_t)jobs_in_flight;
    ctx->job_streams = jobs_in_flight > 2;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_set_bucket_maxima(pngloss_b200_ctx *ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1) return PNGLOSS_B200_INVALID_ARGUMENT;
    ctx->bm = mode;
    return PNGLOSS_B200_SUCCESS;
}

// The stopwatch events sit on the compute stream; the copy streams of the job API are joined into it
// first, so that the interval covers everything enqueued on any of the three.
static std::vector<cudaStream_t> side_streams(pngloss_b200_ctx *ctx) {
    std::vector<cudaStream_t> v;
    if (ctx->h2d) v.push_back(ctx->h2d);
    if (ctx->d2h) v.push_back(ctx->d2h);
    for (pngloss_b200_batch *b : ctx->pool)
        if (b->own_stream) v.push_back(b->stream);
    return v;
}
static int join_copy_streams(pngloss_b200_ctx *ctx, cudaEvent_t scratch) {
    for (cudaStream_t s : side_streams(ctx)) {
        PL_CUDA(ctx, cudaEventRecord(scratch, s));
        PL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, scratch, 0));
    }
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_timer_start(pngloss_b200_ctx *ctx) {
    if (!ctx) return PNGLOSS_B200_INVALID_ARGUMENT;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = join_copy_streams(ctx, ctx->t0)) return rc;
    PL_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    // later work on the copy / job streams must not start before the clock does
    for (cudaStream_t s : side_streams(ctx)) PL_CUDA(ctx, cudaStreamWaitEvent(s, ctx->t0, 0));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_timer_stop(pngloss_b200_ctx *ctx, float *ms) {
    if (!ctx || !ms) return PNGLOSS_B200_INVALID_ARGUMENT;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = join_copy_streams(ctx, ctx->t1)) return rc;
    PL_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
    PL_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    PL_CUDA(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_ctx_sync(pngloss_b200_ctx *ctx) {
    if (!ctx) return PNGLOSS_B200_INVALID_ARGUMENT;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->h2d) PL_CUDA(ctx, cudaStreamSynchronize(ctx->h2d));
    PL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (pngloss_b200_batch *b : ctx->pool)
        if (b->own_stream) PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    if (ctx->d2h) PL_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
    return PNGLOSS_B200_SUCCESS;
}

// Overwrites a buffer larger than the L2 cache on the context's stream, so that the next kernels find their
// inputs in HBM (benchmarks of workloads smaller than L2).
extern "C" int pngloss_b200_ctx_flush_l2(pngloss_b200_ctx *ctx) {
    if (!ctx) return PNGLOSS_B200_INVALID_ARGUMENT;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->scrub) {
        ctx->scrub_bytes = (size_t)512 << 20;
        PL_CUDA(ctx, cudaMalloc((void **)&ctx->scrub, ctx->scrub_bytes));
    }
    PL_CUDA(ctx, cudaMemsetAsync(ctx->scrub, 0x5a, ctx->scrub_bytes, ctx->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" void *pngloss_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void pngloss_b200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---- batch ---------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int pngloss_b200_batch_create(pngloss_b200_ctx *ctx, size_t n, const uint32_t *widths,
                                         const uint32_t *heights, pngloss_b200_batch **out) {
    return pngloss_b200_batch_create_ex(ctx, n, widths, heights, 0u, out);
}

extern "C" int pngloss_b200_batch_create_ex(pngloss_b200_ctx *ctx, size_t n, const uint32_t *widths,
                                            const uint32_t *heights, unsigned flags,
                                            pngloss_b200_batch **out) {
    if (!ctx || !out || !n || !widths || !heights || (flags & ~PNGLOSS_B200_BATCH_IN_PLACE))
        return PNGLOSS_B200_INVALID_ARGUMENT;
    *out = nullptr;
    for (size_t i = 0; i < n; i++) {
        // the reference reader caps rowbytes * height at INT_MAX (src/rwpng.c:286-290)
        if (!widths[i] || !heights[i] || (uint64_t)widths[i] * heights[i] * 4 > 0x7fffffffull)
            return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "image %zu: bad size %ux%u", i,
                           widths[i], heights[i]);
    }
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    pngloss_b200_batch *b = new (std::nothrow) pngloss_b200_batch();
    if (!b) return PNGLOSS_B200_OUT_OF_MEMORY;
    b->ctx = ctx;
    b->stream = ctx->stream;
    b->in_place = (flags & PNGLOSS_B200_BATCH_IN_PLACE) != 0;
    b->n = n;
    b->w.assign(widths, widths + n);
    b->h.assign(heights, heights + n);
    b->himgs.resize(n);
    if (cudaMallocHost((void **)&b->hstatus, n * 4 * sizeof(uint32_t) + 256 * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        delete b;
        return set_err(ctx, PNGLOSS_B200_OUT_OF_MEMORY, "cudaMallocHost(status) failed");
    }
    memset(b->hstatus, 0, n * 4 * sizeof(uint32_t) + 256 * sizeof(unsigned long long));
    b->hbhist = (unsigned long long *)(b->hstatus + n * 4);   // n * 16 bytes in: 8-byte aligned
    b->order.resize(n);
    for (size_t i = 0; i < n; i++) b->order[i] = (int)i;
    std::stable_sort(b->order.begin(), b->order.end(), [&](int a, int c) {
        return b->w[a] != b->w[c] ? b->w[a] < b->w[c] : b->h[a] < b->h[c];
    });

    // slab layout
    size_t off = 0;
    auto take = [&](size_t bytes, size_t al) { off = align_up(off, al); size_t o = off; off += bytes; return o; };
    const size_t o_imgs = take(n * sizeof(PlImageDev), 256);
    const size_t o_slots = take(n * 8 * sizeof(int), 256);   // worst case: one image per CTA of 8 slots
    const size_t o_zero = take(0, 256);
    const size_t o_chan = take(n * PL_FILTERS * 4 * 256 * sizeof(uint32_t), 256);
    const size_t o_flags = take(n * 2 * sizeof(uint32_t), 16);
    const size_t o_status = take(n * 4 * sizeof(uint32_t), 16);
    const size_t o_bhist = take(256 * sizeof(unsigned long long), 16);
    const size_t o_zero_end = take(0, 256);
    const size_t o_final = take(n * 256 * sizeof(uint32_t), 256);
    std::vector<size_t> o_in(n), o_out(n), o_filt(n), o_err(n), o_cand(n), o_oprev(n);
    b->filt_off.resize(n);
    for (size_t i = 0; i < n; i++) {
        b->filt_off[i] = b->filt_bytes;
        b->filt_bytes += align_up(heights[i], 16);
    }
    const size_t o_filters = take(b->filt_bytes, 256);
    // pixel buffers of consecutive images lie back to back (when their size is a multiple of 256 bytes), so
    // that a caller whose host images are contiguous too is served by one copy per direction
    for (size_t i = 0; i < n; i++) o_in[i] = take((size_t)widths[i] * heights[i] * 4, 256);
    for (size_t i = 0; i < n; i++)
        o_out[i] = b->in_place ? o_in[i] : take((size_t)widths[i] * heights[i] * 4, 256);
    for (size_t i = 0; i < n; i++) {
        o_oprev[i] = take((size_t)widths[i] * 4, 256);
        o_filt[i] = o_filters + b->filt_off[i];
        o_err[i] = take((size_t)2 * PL_FILTERS * 2 * (widths[i] + PL_ERR_PAD) * sizeof(short4), 256);
        o_cand[i] = take((size_t)PL_FILTERS * widths[i] * 4, 256);
    }
    b->slab_bytes = align_up(off, 256);
    cudaError_t e = cudaMalloc((void **)&b->slab, b->slab_bytes);
    if (e != cudaSuccess) {
        const size_t want = b->slab_bytes;
        cudaFreeHost(b->hstatus);
        delete b;
        cudaGetLastError();
        return set_err(ctx, PNGLOSS_B200_OUT_OF_MEMORY, "cudaMalloc(%zu bytes) failed: %s", want,
                       cudaGetErrorString(e));
    }
    b->dimgs = (PlImageDev *)(b->slab + o_imgs);
    b->dslots = (int *)(b->slab + o_slots);
    b->zero_begin = b->slab + o_zero;
    b->zero_bytes = o_zero_end - o_zero;
    b->chan_hist = (uint32_t *)(b->slab + o_chan);
    b->flags = (uint32_t *)(b->slab + o_flags);
    b->status = (uint32_t *)(b->slab + o_status);
    b->batch_hist = (unsigned long long *)(b->slab + o_bhist);
    b->dfilters = b->slab + o_filters;
    b->final_hist = (uint32_t *)(b->slab + o_final);
    for (size_t i = 0; i < n; i++) {
        PlImageDev &d = b->himgs[i];
        d.in = (const uchar4 *)(b->slab + o_in[i]);
        d.out = (uchar4 *)(b->slab + o_out[i]);
        d.filters = b->slab + o_filt[i];
        d.chan_hist = b->chan_hist + i * PL_FILTERS * 4 * 256;
        d.flags = b->flags + i * 2;
        d.final_hist = b->final_hist + i * 256;
        d.err = (short4 *)(b->slab + o_err[i]);
        d.cand = (uchar4 *)(b->slab + o_cand[i]);
        d.oprev = (uchar4 *)(b->slab + o_oprev[i]);
        d.status = b->status + i * 4;
        d.width = widths[i];
        d.height = heights[i];
        d.adaptive_all = 0;
        d.force_mode = 0;
    }
    for (int k = 0; k < 4; k++) {
        if (cudaEventCreate(&b->ev[k]) != cudaSuccess) {
            pngloss_b200_batch_destroy(b);
            return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "cudaEventCreate failed");
        }
    }
    *out = b;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" void pngloss_b200_batch_destroy(pngloss_b200_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    // an idle batch of the job pool has nothing in flight (its job's last event was waited for): no need to
    // drain the compute stream it shares with the jobs that are still running
    if (!b->pooled || b->busy) cudaStreamSynchronize(b->stream);
    for (int k = 0; k < 4; k++)
        if (b->ev[k]) cudaEventDestroy(b->ev[k]);
    if (b->ev_up) cudaEventDestroy(b->ev_up);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    if (b->ev_down) cudaEventDestroy(b->ev_down);
    if (b->slab) cudaFree(b->slab);
    if (b->hstatus) cudaFreeHost(b->hstatus);
    if (b->hfilters) cudaFreeHost(b->hfilters);
    if (b->scan_slab) cudaFree(b->scan_slab);
    if (b->own_stream) cudaStreamDestroy(b->stream);
    if (b->hoflags) cudaFreeHost(b->hoflags);
    for (int k = 0; k < 3; k++)
        if (b->ev_scan[k]) cudaEventDestroy(b->ev_scan[k]);
    delete b;
}

extern "C" int pngloss_b200_batch_set_mode(pngloss_b200_batch *b, size_t i, int adaptive_all,
                                           uint32_t force_bpp) {
    if (!b || i >= b->n || force_bpp > 4) return PNGLOSS_B200_INVALID_ARGUMENT;
    b->himgs[i].adaptive_all = adaptive_all ? 1u : 0u;
    b->himgs[i].force_mode = force_bpp;
    b->desc_dirty = true;
    return PNGLOSS_B200_SUCCESS;
}

static int upload_on(pngloss_b200_batch *b, size_t i, const unsigned char *pixels, size_t stride,
                     cudaStream_t stream) {
    if (!b || i >= b->n || !pixels) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t rowbytes = (size_t)b->w[i] * 4;
    if (stride < rowbytes) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "stride < width*4");
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpy2DAsync((void *)b->himgs[i].in, rowbytes, pixels, stride, rowbytes, b->h[i],
                                   cudaMemcpyHostToDevice, stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_upload(pngloss_b200_batch *b, size_t i, const unsigned char *pixels,
                                         size_t stride) {
    return upload_on(b, i, pixels, stride, b ? b->stream : nullptr);
}

// rows[] as the reference hands them over: usually equally spaced (src/rwpng.c lays rgba_data out
// contiguously), but the contract allows arbitrary pointers.
static bool constant_stride(unsigned char *const *rows, uint32_t h, size_t rowbytes, size_t *stride) {
    if (h == 1) { *stride = rowbytes; return true; }
    if (rows[1] < rows[0]) return false;
    const size_t s = (size_t)(rows[1] - rows[0]);
    if (s < rowbytes) return false;
    for (uint32_t y = 2; y < h; y++)
        if (rows[y] != rows[0] + (size_t)y * s) return false;
    *stride = s;
    return true;
}

extern "C" int pngloss_b200_batch_upload_rows(pngloss_b200_batch *b, size_t i,
                                              unsigned char *const *rows) {
    if (!b || i >= b->n || !rows) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t rowbytes = (size_t)b->w[i] * 4;
    size_t stride = 0;
    if (constant_stride(rows, b->h[i], rowbytes, &stride))
        return pngloss_b200_batch_upload(b, i, rows[0], stride);
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t y = 0; y < b->h[i]; y++)
        PL_CUDA(ctx, cudaMemcpyAsync((unsigned char *)b->himgs[i].in + (size_t)y * rowbytes, rows[y],
                                     rowbytes, cudaMemcpyHostToDevice, b->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_synth(pngloss_b200_batch *b, size_t i, uint64_t seed) {
    if (!b || i >= b->n) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t px = (size_t)b->w[i] * b->h[i];
    const unsigned blocks = (unsigned)std::min<size_t>((px + 255) / 256, 148 * 16);
    pl_k_synth<<<blocks, 256, 0, b->stream>>>((uchar4 *)b->himgs[i].in, b->w[i], b->h[i], seed);
    PL_CUDA(ctx, cudaGetLastError());
    return PNGLOSS_B200_SUCCESS;
}

template <int LPC, bool BM>
static int launch_k2(pngloss_b200_batch *b, int nblocks, unsigned strength, long bleed) {
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t smem = sizeof(PlCtaSmem<LPC, BM>) + PL_K2_SMEM_ALIGN;   // slack for the in-kernel alignment
    // per device and cheap: set on every launch rather than cached in a static shared by the per-GPU threads
    PL_CUDA(ctx, cudaFuncSetAttribute(pl_k2_quantize<LPC, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    pl_k2_quantize<LPC, BM><<<nblocks, PL_K2_THREADS, smem, b->stream>>>(b->dimgs, b->dslots, (int)strength,
                                                                          (int)bleed);
    PL_CUDA(ctx, cudaGetLastError());
    b->info[0] = (uint32_t)nblocks;
    b->info[1] = (uint32_t)PlCfg<LPC>::CPW;
    b->info[2] = (uint32_t)smem;
    return PNGLOSS_B200_SUCCESS;
}
// the lean kernel: 8 images per CTA, three CTAs per SM (pl_k2_lean.cuh)
static int launch_k2_lean(pngloss_b200_batch *b, int nblocks, unsigned strength, long bleed) {
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t smem = sizeof(PlLeanSmem) + PL_L_SMEM_ALIGN;
    PL_CUDA(ctx, cudaFuncSetAttribute(pl_k2_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PL_CUDA(ctx, cudaFuncSetAttribute(pl_k2_lean, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
    pl_k2_lean<<<nblocks, PL_K2_THREADS, smem, b->stream>>>(b->dimgs, b->dslots, (int)strength, (int)bleed);
    PL_CUDA(ctx, cudaGetLastError());
    b->info[0] = (uint32_t)nblocks;
    b->info[1] = (uint32_t)PL_L_CPW;
    b->info[2] = (uint32_t)smem;
    return PNGLOSS_B200_SUCCESS;
}

// the latency kernel: one image per CTA, chain / producer / post warps (pl_k2_solo.cuh)
template <int FPW, bool COMPACT>
static int launch_k2_solo(pngloss_b200_batch *b, int nblocks, unsigned strength, long bleed) {
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t smem = sizeof(PlSoloSmem) + 16;
    PL_CUDA(ctx, cudaFuncSetAttribute(pl_k2_solo<FPW, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PL_CUDA(ctx, cudaFuncSetAttribute(pl_k2_solo<FPW, COMPACT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
    pl_k2_solo<FPW, COMPACT><<<nblocks, PlSoloCfg<FPW, COMPACT>::THREADS, smem, b->stream>>>(
        b->dimgs, b->dslots, (int)strength, (int)bleed, (unsigned)(ctx->sm_count > 0 ? ctx->sm_count : 148));
    PL_CUDA(ctx, cudaGetLastError());
    b->info[0] = (uint32_t)nblocks;
    b->info[1] = 1;
    b->info[2] = (uint32_t)smem;
    return PNGLOSS_B200_SUCCESS;
}

template <int LPC>
static int launch_k2(pngloss_b200_batch *b, int nblocks, unsigned strength, long bleed, bool bm) {
    return bm ? launch_k2<LPC, true>(b, nblocks, strength, bleed)
              : launch_k2<LPC, false>(b, nblocks, strength, bleed);
}

// Lane mapping by batch size, from the B200 sweeps in profiles/ (r1_sweep_lanes_bm.txt, r2_sweep_lean.txt).  A
// chain's step costs 0.51 / 0.58 / 0.65 / 0.67 us with 8 / 4 / 2 / 1 lanes per channel when a CTA has an SM to
// itself, and 1.3x / 1.7x / 2.1x that with 2 / 3 / 4 CTAs per SM - so: the widest lane groups (fewest images per
// CTA) that still give every CTA its own SM, and 8 images per CTA once even that needs more CTAs than SMs.
static int choose_lpc(const pngloss_b200_batch *b) {
    if (b->ctx->lpc) return b->ctx->lpc;
    const size_t sms = b->ctx->sm_count > 0 ? (size_t)b->ctx->sm_count : 148;
    const size_t per_cta = (b->n + sms - 1) / sms;   // images a CTA must take for one CTA per SM
    return per_cta <= 1 ? 8 : per_cta <= 2 ? 4 : per_cta <= 4 ? 2 : 1;
}

// The latency kernel (pl_k2_solo.cuh: 3.0 against 2.0 Mpx/s per image, profiles/r2_sweep_solo.txt) takes a batch when
// every image can have a CTA of its own with at most four CTAs per SM (eight-warp CTAs up to two per SM, four-warp
// CTAs beyond: 592 images run at 1224 Mpx/s against 883 for the generic kernel at two lanes per channel), and
// the strength is at most 126 (where its winner tables exist).  0 = no, 5 / 1 = filter candidates per chain warp.  An explicit lane mapping
// (pngloss_b200_ctx_set_lanes) keeps the generic kernel unless the latency kernel is asked for explicitly.
#ifndef PL_SOLO_MAX_CTAS_PER_SM
#define PL_SOLO_MAX_CTAS_PER_SM 4   /* four-warp CTAs: 122 registers x 128 threads, 55 KB of shared memory */
#endif
static int use_solo(const pngloss_b200_batch *b, unsigned strength) {
    const pngloss_b200_ctx *ctx = b->ctx;
    uint32_t wmax = 0;
    for (size_t i = 0; i < b->n; i++) wmax = std::max(wmax, b->w[i]);
    // (its tables have room for the one-symbol buckets of the smallest strengths, unlike the other kernels')
    const bool table = strength + 1 <= PL_BM_MAX_STEP && wmax < PL_BM_MAX_WIDTH;
    if (!table || ctx->solo == 0) return 0;
    if (ctx->solo > 0) return (ctx->lpc == 0 || ctx->lpc == 8) ? (ctx->solo == 2 ? 1 : 5) : 0;
    const size_t sms = ctx->sm_count > 0 ? (size_t)ctx->sm_count : 148;
    return (ctx->lpc == 0 && b->n <= PL_SOLO_MAX_CTAS_PER_SM * sms) ? 5 : 0;
}

// Host-side part of a run: CTA packing, descriptors and the cleared accumulators, enqueued on `stream`
// (the batch's own stream, or the job API's upload stream).
static int prepare_run(pngloss_b200_batch *b, unsigned strength, long bleed, cudaStream_t stream, int *lpc_out,
                       int *nblocks_out) {
    pngloss_b200_ctx *ctx = b->ctx;
    // same ranges as the CLI checks (reference src/pngloss.c:123,128)
    if (strength > 255 || bleed < 1 || bleed > 32767)
        return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "strength 0..255, bleed 1..32767");
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int lpc = use_solo(b, strength) ? 8 : choose_lpc(b);
    const int cpw = 8 / lpc;
    // pack equally sized images into CTAs of cpw slots
    b->hslots.clear();
    size_t k = 0;
    while (k < b->n) {
        size_t e = k;
        while (e < b->n && e - k < (size_t)cpw && b->w[b->order[e]] == b->w[b->order[k]] &&
               b->h[b->order[e]] == b->h[b->order[k]])
            e++;
        for (size_t s = 0; s < (size_t)cpw; s++) b->hslots.push_back(k + s < e ? b->order[k + s] : -1);
        k = e;
    }
    *lpc_out = lpc;
    *nblocks_out = (int)(b->hslots.size() / cpw);
    if (b->desc_dirty) {
        PL_CUDA(ctx, cudaMemcpyAsync(b->dimgs, b->himgs.data(), b->n * sizeof(PlImageDev),
                                     cudaMemcpyHostToDevice, stream));
        b->desc_dirty = false;
    }
    PL_CUDA(ctx, cudaMemcpyAsync(b->dslots, b->hslots.data(), b->hslots.size() * sizeof(int),
                                 cudaMemcpyHostToDevice, stream));
    return PNGLOSS_B200_SUCCESS;
}

static int launch_run(pngloss_b200_batch *b, unsigned strength, long bleed, int lpc, int nblocks);

extern "C" int pngloss_b200_batch_run(pngloss_b200_batch *b, unsigned strength, long bleed) {
    if (!b) return PNGLOSS_B200_INVALID_ARGUMENT;
    int lpc = 0, nblocks = 0;
    const int rc = prepare_run(b, strength, bleed, b->stream, &lpc, &nblocks);
    return rc ? rc : launch_run(b, strength, bleed, lpc, nblocks);
}

// The three kernels of a run on the batch's (compute) stream.
static int launch_run(pngloss_b200_batch *b, unsigned strength, long bleed, int lpc, int nblocks) {
    pngloss_b200_ctx *ctx = b->ctx;
    // The accumulators are cleared here, on the compute stream, and not with the uploads: a memset is a
    // kernel, and while another job's K2 is resident (its two CTAs take all of an SM's shared memory) no
    // other kernel gets onto the SMs - on the upload stream it would hold the next job's pixels back until
    // that K2 has finished (profiles/r1_e2e_job_timeline.txt).
    PL_CUDA(ctx, cudaMemsetAsync(b->zero_begin, 0, b->zero_bytes, b->stream));
    // K1: enough row slices per image to fill the machine, capped by the image height
    uint32_t hmin = b->h[0];
    for (size_t i = 1; i < b->n; i++) hmin = std::min(hmin, b->h[i]);
    size_t want = (4 * 148 + b->n - 1) / b->n;
    unsigned slices = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, 256), hmin));
    PL_CUDA(ctx, cudaEventRecord(b->ev[0], b->stream));
    pl_k1_orig_hist<<<(unsigned)(b->n * slices), PL_K1_THREADS, 0, b->stream>>>(b->dimgs, slices);
    PL_CUDA(ctx, cudaGetLastError());
    PL_CUDA(ctx, cudaEventRecord(b->ev[1], b->stream));
    // Bucket maxima replace the per-byte candidate scan by a table look-up; the table only exists for
    // PL_BM_MIN_STEP <= strength + 1 <= PL_BM_MAX_STEP (below, the scan is short anyway).
    uint32_t wmax = 0;
    for (size_t i = 0; i < b->n; i++) wmax = std::max(wmax, b->w[i]);
    // ... and it only pays where a lane scans many candidates itself (one or two lanes per channel; measured,
    // profiles/r1_sweep_lanes_bm.txt)
    const bool bm = ctx->bm >= 0 ? ctx->bm != 0
                                 : (strength + 1 >= PL_BM_MIN_STEP && strength + 1 <= PL_BM_MAX_STEP &&
                                    wmax < PL_BM_MAX_WIDTH && lpc <= 2);
    // the lean kernel needs 16-byte aligned rows for its bulk copies
    bool w4 = true;
    for (size_t i = 0; i < b->n; i++) w4 = w4 && (b->w[i] & 3u) == 0;
    // ... and it only wins where its third CTA per SM is used (measured, profiles/r2_sweep_lean.txt: 3164 against
    // 2551 Mpx/s at 444 CTAs, 2407 against 2978 at 296): by default only for grids beyond two CTAs per SM
    const bool lean = lpc == 1 && bm && w4 && wmax < PL_BM_MAX_WIDTH &&
                      (ctx->lean > 0 || (ctx->lean < 0 && nblocks > 2 * (ctx->sm_count > 0 ? ctx->sm_count : 148)));
    const int solo = lpc == 8 ? use_solo(b, strength) : 0;
    int rc;
    // (one chain warp: eight warps per CTA up to two CTAs per SM, the four-warp layout beyond)
    const bool compact = ctx->solo == 3 || (ctx->solo < 0 && nblocks > 2 * (ctx->sm_count > 0 ? ctx->sm_count : 148));
    if (solo) rc = solo == 1 ? launch_k2_solo<1, false>(b, nblocks, strength, bleed)
                   : compact ? launch_k2_solo<5, true>(b, nblocks, strength, bleed)
                             : launch_k2_solo<5, false>(b, nblocks, strength, bleed);
    else if (lean) rc = launch_k2_lean(b, nblocks, strength, bleed);
    else
    switch (lpc) {
    case 8: rc = launch_k2<8>(b, nblocks, strength, bleed, bm); break;
    case 4: rc = launch_k2<4>(b, nblocks, strength, bleed, bm); break;
    case 2: rc = launch_k2<2>(b, nblocks, strength, bleed, bm); break;
    default: rc = launch_k2<1>(b, nblocks, strength, bleed, bm); break;
    }
    if (rc) return rc;
    PL_CUDA(ctx, cudaEventRecord(b->ev[2], b->stream));
    pl_k3_batch_hist<<<(unsigned)std::min<size_t>(b->n, 64), 256, 0, b->stream>>>(b->dimgs, (int)b->n,
                                                                                    b->batch_hist);
    PL_CUDA(ctx, cudaGetLastError());
    PL_CUDA(ctx, cudaEventRecord(b->ev[3], b->stream));
    b->info[3] = 3 | ((bm || solo) ? 0x100u : 0u) | ((lean && !solo) ? 0x200u : 0u) | (solo ? 0x400u : 0u);
    b->ran = true;
    return PNGLOSS_B200_SUCCESS;
}

static int download_on(pngloss_b200_batch *b, size_t i, unsigned char *pixels, size_t stride,
                       unsigned char *row_filters, cudaStream_t stream) {
    if (!b || i >= b->n) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t rowbytes = (size_t)b->w[i] * 4;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (pixels) {
        if (stride < rowbytes) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "stride < width*4");
        PL_CUDA(ctx, cudaMemcpy2DAsync(pixels, stride, b->himgs[i].out, rowbytes, rowbytes, b->h[i],
                                       cudaMemcpyDeviceToHost, stream));
    }
    if (row_filters)
        PL_CUDA(ctx, cudaMemcpyAsync(row_filters, b->himgs[i].filters, b->h[i], cudaMemcpyDeviceToHost,
                                     stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_download(pngloss_b200_batch *b, size_t i, unsigned char *pixels,
                                           size_t stride, unsigned char *row_filters) {
    return download_on(b, i, pixels, stride, row_filters, b ? b->stream : nullptr);
}

extern "C" int pngloss_b200_batch_download_rows(pngloss_b200_batch *b, size_t i,
                                                unsigned char *const *rows, unsigned char *row_filters) {
    if (!b || i >= b->n || !rows) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t rowbytes = (size_t)b->w[i] * 4;
    size_t stride = 0;
    if (constant_stride(rows, b->h[i], rowbytes, &stride))
        return pngloss_b200_batch_download(b, i, rows[0], stride, row_filters);
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t y = 0; y < b->h[i]; y++)
        PL_CUDA(ctx, cudaMemcpyAsync(rows[y], (const unsigned char *)b->himgs[i].out + (size_t)y * rowbytes,
                                     rowbytes, cudaMemcpyDeviceToHost, b->stream));
    if (row_filters)
        PL_CUDA(ctx, cudaMemcpyAsync(row_filters, b->himgs[i].filters, b->h[i], cudaMemcpyDeviceToHost,
                                     b->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_download_input(pngloss_b200_batch *b, size_t i, unsigned char *pixels,
                                                 size_t stride) {
    if (!b || i >= b->n || !pixels) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    const size_t rowbytes = (size_t)b->w[i] * 4;
    if (stride < rowbytes) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "stride < width*4");
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpy2DAsync(pixels, stride, b->himgs[i].in, rowbytes, rowbytes, b->h[i],
                                   cudaMemcpyDeviceToHost, b->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_finish(pngloss_b200_batch *b, int *status, uint32_t *bpp,
                                         uint32_t *retried) {
    if (!b) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (b->ran)
        PL_CUDA(ctx, cudaMemcpyAsync(b->hstatus, b->status, b->n * 4 * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, b->stream));
    PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    int first = PNGLOSS_B200_SUCCESS;
    for (size_t i = 0; i < b->n; i++) {
        const int st = !b->ran ? PNGLOSS_B200_SUCCESS
                       : b->hstatus[i * 4 + 0] == PL_ST_OK ? PNGLOSS_B200_SUCCESS
                                                           : PNGLOSS_B200_NO_ACCEPTABLE_ROW;
        if (status) status[i] = st;
        if (bpp) bpp[i] = b->hstatus[i * 4 + 1];
        if (retried) retried[i] = b->hstatus[i * 4 + 2];
        if (st && !first) {
            first = st;
            set_err(ctx, st, "image %zu: no acceptable row even at strength 0", i);
        }
    }
    return first;
}

extern "C" int pngloss_b200_batch_image_histogram(pngloss_b200_batch *b, size_t i, uint32_t *out256) {
    if (!b || i >= b->n || !out256) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpyAsync(out256, b->himgs[i].final_hist, 256 * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, b->stream));
    PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    return PNGLOSS_B200_SUCCESS;
}

// K1's output for one image: original_frequency per (filter, RGBA channel), 5 x 4 x 256 counts.  Summed over the
// channels the image's colour mode uses it is the reference's optimize_state_init table (src/optimize_state.c:66-83).
extern "C" int pngloss_b200_batch_image_original_histogram(pngloss_b200_batch *b, size_t i, uint32_t *out5x4x256) {
    if (!b || i >= b->n || !out5x4x256) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpyAsync(out5x4x256, b->himgs[i].chan_hist, PL_FILTERS * 4 * 256 * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, b->stream));
    PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_histogram(pngloss_b200_batch *b, uint64_t *out256) {
    if (!b || !out256) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpyAsync(out256, b->batch_hist, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 b->stream));
    PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" void *pngloss_b200_batch_histogram_device(pngloss_b200_batch *b) {
    return b ? (void *)b->batch_hist : nullptr;
}

extern "C" int pngloss_b200_batch_timings(pngloss_b200_batch *b, float ms[4]) {
    if (!b || !ms || !b->ran) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaEventSynchronize(b->ev[3]));
    PL_CUDA(ctx, cudaEventElapsedTime(&ms[0], b->ev[0], b->ev[1]));
    PL_CUDA(ctx, cudaEventElapsedTime(&ms[1], b->ev[1], b->ev[2]));
    PL_CUDA(ctx, cudaEventElapsedTime(&ms[2], b->ev[2], b->ev[3]));
    PL_CUDA(ctx, cudaEventElapsedTime(&ms[3], b->ev[0], b->ev[3]));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_launch_info(pngloss_b200_batch *b, uint32_t info[4]) {
    if (!b || !info) return PNGLOSS_B200_INVALID_ARGUMENT;
    memcpy(info, b->info, sizeof b->info);
    return PNGLOSS_B200_SUCCESS;
}

// ---- K4: filtered scanlines ---------------------------------------------------------------------------------
extern "C" int pngloss_b200_batch_scanlines(pngloss_b200_batch *b) {
    if (!b) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    if (!b->ran) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "scanlines: the batch has not been run");
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = b->n;
    if (!b->scan_slab) {
        // worst case 4 bytes per pixel; the colour type is only known once the output has been scanned
        size_t off = 0;
        auto take = [&](size_t bytes, size_t al) { off = align_up(off, al); size_t o = off; off += bytes; return o; };
        const size_t o_desc = take(n * sizeof(PlScanDev), 256);
        const size_t o_flags = take(n * 4 * sizeof(uint32_t), 256);
        b->scan_off.resize(n);
        for (size_t i = 0; i < n; i++) b->scan_off[i] = take((size_t)b->h[i] * (1 + (size_t)b->w[i] * 4), 256);
        cudaError_t e = cudaMalloc((void **)&b->scan_slab, align_up(off, 256));
        if (e != cudaSuccess) {
            cudaGetLastError();
            b->scan_slab = nullptr;
            return set_err(ctx, PNGLOSS_B200_OUT_OF_MEMORY, "cudaMalloc(%zu bytes of scanlines) failed", off);
        }
        b->dscan = (PlScanDev *)(b->scan_slab + o_desc);
        b->oflags = (uint32_t *)(b->scan_slab + o_flags);
        // any failure below undoes the whole block, so that the next call starts over instead of launching K4 on
        // a descriptor table that was never uploaded
        auto undo = [&]() {
            cudaGetLastError();
            cudaFree(b->scan_slab);
            b->scan_slab = nullptr;
            if (b->hoflags) cudaFreeHost(b->hoflags);
            b->hoflags = nullptr;
            for (int k = 0; k < 3; k++) {
                if (b->ev_scan[k]) cudaEventDestroy(b->ev_scan[k]);
                b->ev_scan[k] = nullptr;
            }
        };
        if (cudaMallocHost((void **)&b->hoflags, n * 4 * sizeof(uint32_t)) != cudaSuccess ||
            cudaEventCreate(&b->ev_scan[0]) != cudaSuccess || cudaEventCreate(&b->ev_scan[1]) != cudaSuccess ||
            cudaEventCreate(&b->ev_scan[2]) != cudaSuccess) {
            undo();
            return set_err(ctx, PNGLOSS_B200_OUT_OF_MEMORY, "scanlines: host-side allocation failed");
        }
        std::vector<PlScanDev> h(n);
        for (size_t i = 0; i < n; i++) {
            h[i].px = b->himgs[i].out;
            h[i].filters = b->himgs[i].filters;
            h[i].scan = b->scan_slab + b->scan_off[i];
            h[i].oflags = b->oflags + i * 4;
            h[i].width = b->w[i];
            h[i].height = b->h[i];
        }
        if (cudaMemcpyAsync(b->dscan, h.data(), n * sizeof(PlScanDev), cudaMemcpyHostToDevice, b->stream) != cudaSuccess ||
            cudaStreamSynchronize(b->stream) != cudaSuccess) {   // (h goes out of scope)
            undo();
            return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "scanlines: descriptor upload failed");
        }
    }
    PL_CUDA(ctx, cudaMemsetAsync(b->oflags, 0, n * 4 * sizeof(uint32_t), b->stream));
    uint32_t hmin = b->h[0];
    for (size_t i = 1; i < n; i++) hmin = std::min(hmin, b->h[i]);
    const size_t want = (8 * 148 + n - 1) / n;   // ~8 CTAs of 256 threads per SM
    const unsigned slices = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, 1024), hmin));
    PL_CUDA(ctx, cudaEventRecord(b->ev_scan[0], b->stream));
    pl_k4_scan_output<<<(unsigned)(n * slices), PL_K4_THREADS, 0, b->stream>>>(b->dscan, slices);
    PL_CUDA(ctx, cudaGetLastError());
    PL_CUDA(ctx, cudaEventRecord(b->ev_scan[1], b->stream));
    pl_k4_scanlines<<<(unsigned)(n * slices), PL_K4_THREADS, 0, b->stream>>>(b->dscan, slices);
    PL_CUDA(ctx, cudaGetLastError());
    PL_CUDA(ctx, cudaEventRecord(b->ev_scan[2], b->stream));
    PL_CUDA(ctx, cudaMemcpyAsync(b->hoflags, b->oflags, n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
    b->scan_ran = true;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_scanline_info(pngloss_b200_batch *b, size_t i, uint32_t *bytes_per_pixel,
                                                uint32_t *row0_filter, size_t *bytes, float milliseconds[2]) {
    if (!b || i >= b->n || !b->scan_ran) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaStreamSynchronize(b->stream));
    const uint32_t *f = b->hoflags + i * 4;
    const uint32_t bpp = f[0] ? (f[1] ? 4u : 3u) : (f[1] ? 2u : 1u);
    if (bytes_per_pixel) *bytes_per_pixel = bpp;
    if (row0_filter) *row0_filter = f[2];
    if (bytes) *bytes = (size_t)b->h[i] * (1 + (size_t)b->w[i] * bpp);
    if (milliseconds) {
        PL_CUDA(ctx, cudaEventElapsedTime(&milliseconds[0], b->ev_scan[0], b->ev_scan[1]));
        PL_CUDA(ctx, cudaEventElapsedTime(&milliseconds[1], b->ev_scan[1], b->ev_scan[2]));
    }
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_batch_download_scanlines(pngloss_b200_batch *b, size_t i, unsigned char *dst,
                                                     size_t capacity) {
    if (!b || i >= b->n || !dst || !b->scan_ran) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    size_t bytes = 0;
    int rc = pngloss_b200_batch_scanline_info(b, i, nullptr, nullptr, &bytes, nullptr);
    if (rc) return rc;
    if (capacity < bytes) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "scanlines: buffer too small");
    PL_CUDA(ctx, cudaMemcpyAsync(dst, b->scan_slab + b->scan_off[i], bytes, cudaMemcpyDeviceToHost, b->stream));
    return PNGLOSS_B200_SUCCESS;
}

// Upper bound of what an in-place batch allocates for one image (see the slab layout in batch_create_ex).
static size_t device_bytes_per_image(uint32_t w, uint32_t h, bool scanlines) {
    const size_t px = (size_t)w * h;
    return (scanlines ? align_up((size_t)h * (1 + (size_t)w * 4), 256) + sizeof(PlScanDev) + 16 : 0) +
           align_up(px * 4, 256) + align_up((size_t)w * 4, 256) + align_up(h, 16) +
           align_up((size_t)2 * PL_FILTERS * 2 * (w + PL_ERR_PAD) * sizeof(short4), 256) +
           align_up((size_t)PL_FILTERS * w * 4, 256) + PL_FILTERS * 4 * 256 * sizeof(uint32_t) +
           256 * sizeof(uint32_t) + sizeof(PlImageDev) + 8 * sizeof(int) + 64 + 1024;
}

// ---- host-buffer jobs -----------------------------------------------------------------------------------
// A job is one batch of host images on its way through the device: uploads on the context's upload
// stream, the three kernels on its compute stream, downloads on its download stream, chained by events.
// Kernels of consecutive jobs run back to back on the one compute stream (two K2 grids sharing the SMs
// was measured and rejected); what overlaps are the PCIe copies of one job with the kernels of
// another.  Device batches are in-place (the quantised rows overwrite the uploaded ones), so two jobs of
// the bench's size fit next to each other.
static int finalize_job(pngloss_b200_job *job) {
    if (job->finalized) return job->rc;
    pngloss_b200_ctx *ctx = job->ctx;
    pngloss_b200_batch *b = job->batch;
    job->finalized = true;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaEventSynchronize(b->ev_down);
    int first = 0;
    if (e != cudaSuccess) {
        first = set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "job failed: %s", cudaGetErrorString(e));
        for (size_t i = 0; i < job->n; i++) job->images[i].status = first;
    } else {
        for (size_t i = 0; i < job->n; i++) {
            const int st = b->hstatus[i * 4 + 0] == PL_ST_OK ? PNGLOSS_B200_SUCCESS : PNGLOSS_B200_NO_ACCEPTABLE_ROW;
            job->images[i].status = st;
            job->images[i].bytes_per_pixel = b->hstatus[i * 4 + 1];
            job->images[i].retried_rows = b->hstatus[i * 4 + 2];
            if (job->images[i].row_filters)
                memcpy(job->images[i].row_filters, b->hfilters + b->filt_off[i], b->h[i]);
            if (st && !first)
                first = set_err(ctx, st, "image %zu: no acceptable row even at strength 0", i);
        }
        for (int k = 0; k < 256; k++) ctx->symbols[k] += b->hbhist[k];
        if (job->scanlines) {
            // the colour types are known now: fetch exactly the bytes each image's scanlines take
            for (size_t i = 0; i < job->n && e == cudaSuccess; i++) {
                pngloss_b200_image &im = job->images[i];
                if (!im.scanlines) continue;
                const uint32_t *f = b->hoflags + i * 4;
                im.scan_bytes_per_pixel = f[0] ? (f[1] ? 4u : 3u) : (f[1] ? 2u : 1u);
                im.scan_row0_filter = f[2];
                im.scan_bytes = (size_t)b->h[i] * (1 + (size_t)b->w[i] * im.scan_bytes_per_pixel);
                e = cudaMemcpyAsync(im.scanlines, b->scan_slab + b->scan_off[i], im.scan_bytes,
                                    cudaMemcpyDeviceToHost, ctx->d2h);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->d2h);
            if (e != cudaSuccess) {
                first = set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "scanline download failed: %s", cudaGetErrorString(e));
                for (size_t i = 0; i < job->n; i++) job->images[i].status = first;
            }
        }
    }
    b->busy = false;
    job->rc = first;
    return first;
}

static void trim_pool(pngloss_b200_ctx *ctx, size_t keep) {
    // drop idle batches (oldest first) until at most `keep` batches remain
    for (size_t k = 0; k < ctx->pool.size() && ctx->pool.size() > keep;) {
        if (!ctx->pool[k]->busy) {
            pngloss_b200_batch_destroy(ctx->pool[k]);
            ctx->pool.erase(ctx->pool.begin() + (long)k);
        } else {
            k++;
        }
    }
}

// An idle in-place batch of exactly these shapes, recycled (a CLI or service feeding equal batches) or new.
static int acquire_batch(pngloss_b200_ctx *ctx, const std::vector<uint32_t> &w, const std::vector<uint32_t> &h,
                         pngloss_b200_batch **out) {
    for (;;) {
        for (pngloss_b200_batch *b : ctx->pool)
            if (!b->busy && b->w == w && b->h == h) { *out = b; return 0; }
        // idle batches of other shapes only hold memory
        for (size_t k = 0; k < ctx->pool.size();) {
            if (!ctx->pool[k]->busy) {
                pngloss_b200_batch_destroy(ctx->pool[k]);
                ctx->pool.erase(ctx->pool.begin() + (long)k);
            } else {
                k++;
            }
        }
        pngloss_b200_batch *b = nullptr;
        int rc = ctx->pool.size() >= ctx->max_jobs ? PNGLOSS_B200_OUT_OF_MEMORY
                                       : pngloss_b200_batch_create_ex(ctx, w.size(), w.data(), h.data(),
                                                                      PNGLOSS_B200_BATCH_IN_PLACE, &b);
        if (!rc) {
            if (cudaEventCreateWithFlags(&b->ev_up, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b->ev_down, cudaEventDisableTiming) != cudaSuccess) {
                pngloss_b200_batch_destroy(b);
                return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "cudaEventCreate failed");
            }
            if (cudaMallocHost((void **)&b->hfilters, b->filt_bytes) != cudaSuccess) {
                cudaGetLastError();
                pngloss_b200_batch_destroy(b);
                return set_err(ctx, PNGLOSS_B200_OUT_OF_MEMORY, "cudaMallocHost(row filters) failed");
            }
            if (ctx->job_streams) {
                // pipeline of many small jobs: their kernels run side by side (a job's CTAs take the places another
                // job's CTAs leave), so every batch computes on a stream of its own
                cudaStream_t st = nullptr;
                if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
                    pngloss_b200_batch_destroy(b);
                    return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "cudaStreamCreate failed");
                }
                b->stream = st;
                b->own_stream = true;
            }
            b->pooled = true;
            ctx->pool.push_back(b);
            *out = b;
            return 0;
        }
        // no room (or the pipeline is full): let the oldest job in flight finish and take its place
        if (rc != PNGLOSS_B200_OUT_OF_MEMORY || ctx->inflight.empty()) return rc;
        pngloss_b200_job *oldest = nullptr;
        for (pngloss_b200_job *j : ctx->inflight)
            if (!j->finalized) { oldest = j; break; }
        if (!oldest) return rc;
        finalize_job(oldest);
    }
}

extern "C" int pngloss_b200_submit(pngloss_b200_ctx *ctx, pngloss_b200_image *images, size_t n,
                                   unsigned strength, long bleed, pngloss_b200_job **out) {
    if (!ctx || !images || !n || !out) return PNGLOSS_B200_INVALID_ARGUMENT;
    *out = nullptr;
    for (size_t i = 0; i < n; i++)
        if (!images[i].pixels) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "image %zu: NULL pixels", i);
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->h2d) PL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
    if (!ctx->d2h) PL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
    std::vector<uint32_t> w(n), h(n);
    for (size_t i = 0; i < n; i++) {
        w[i] = images[i].width;
        h[i] = images[i].height;
    }
    pngloss_b200_batch *b = nullptr;
    int rc = acquire_batch(ctx, w, h, &b);
    if (rc) return rc;
    pngloss_b200_job *job = new (std::nothrow) pngloss_b200_job();
    if (!job) return PNGLOSS_B200_OUT_OF_MEMORY;
    job->ctx = ctx;
    job->batch = b;
    job->images = images;
    job->n = n;

    // upload stream: descriptors (small copies from pageable memory, which wait for what the stream already
    // holds - so they go first), then the pixels
    int lpc = 0, nblocks = 0;
    for (size_t i = 0; i < n && !rc; i++)
        rc = pngloss_b200_batch_set_mode(b, i, images[i].row_filters == nullptr, images[i].force_bytes_per_pixel);
    if (!rc) rc = prepare_run(b, strength, bleed, ctx->h2d, &lpc, &nblocks);
    // (images that are contiguous on the host and on the device travel as one copy: thousands of queued
    // copies would fill the driver's queue and block this call until the kernels have run)
    for (size_t i = 0; i < n && !rc;) {
        size_t e2 = i + 1, bytes = (size_t)b->w[i] * b->h[i] * 4;
        if (images[i].stride == (size_t)b->w[i] * 4) {
            while (e2 < n && images[e2].stride == (size_t)b->w[e2] * 4 &&
                   images[e2].pixels == images[i].pixels + bytes &&
                   (const unsigned char *)b->himgs[e2].in == (const unsigned char *)b->himgs[i].in + bytes) {
                bytes += (size_t)b->w[e2] * b->h[e2] * 4;
                e2++;
            }
        }
        if (e2 > i + 1) {
            if (cudaMemcpyAsync((void *)b->himgs[i].in, images[i].pixels, bytes, cudaMemcpyHostToDevice, ctx->h2d) !=
                cudaSuccess)
                rc = set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            rc = upload_on(b, i, images[i].pixels, images[i].stride, ctx->h2d);
        }
        i = e2;
    }
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaEventRecord(b->ev_up, ctx->h2d);
    // compute stream: kernels, then the per-image status words
    if (!rc && e == cudaSuccess) e = cudaStreamWaitEvent(b->stream, b->ev_up, 0);
    if (!rc && e == cudaSuccess) rc = launch_run(b, strength, bleed, lpc, nblocks);
    for (size_t i = 0; i < n; i++) job->scanlines |= images[i].scanlines != nullptr;
    if (!rc && e == cudaSuccess && job->scanlines) {
        b->ran = true;
        rc = pngloss_b200_batch_scanlines(b);   // K4 on the compute stream; flags come back with it
    }
    if (!rc && e == cudaSuccess)
        e = cudaMemcpyAsync(b->hstatus, b->status, b->n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            b->stream);
    if (!rc && e == cudaSuccess)
        e = cudaMemcpyAsync(b->hbhist, b->batch_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            b->stream);
    if (!rc && e == cudaSuccess) e = cudaEventRecord(b->ev_done, b->stream);
    // download stream: pixels (to out_pixels, or back over the input) and row filters
    if (!rc && e == cudaSuccess) e = cudaStreamWaitEvent(ctx->d2h, b->ev_done, 0);
    // (a copy into pageable memory would block this call until the kernels are done, so the row filters
    // travel through a pinned staging buffer and reach the caller's arrays in pngloss_b200_wait)
    auto dst_of = [&](size_t i) { return images[i].out_pixels ? images[i].out_pixels : images[i].pixels; };
    auto dstride_of = [&](size_t i) { return images[i].out_pixels ? images[i].out_stride : images[i].stride; };
    auto skip_px = [&](size_t i) { return images[i].scanlines && (images[i].flags & PNGLOSS_B200_IMAGE_NO_PIXELS); };
    for (size_t i = 0; i < n && !rc && e == cudaSuccess;) {
        if (skip_px(i)) { i++; continue; }
        size_t e2 = i + 1, bytes = (size_t)b->w[i] * b->h[i] * 4;
        if (dstride_of(i) == (size_t)b->w[i] * 4) {
            while (e2 < n && !skip_px(e2) && dstride_of(e2) == (size_t)b->w[e2] * 4 && dst_of(e2) == dst_of(i) + bytes &&
                   (const unsigned char *)b->himgs[e2].out == (const unsigned char *)b->himgs[i].out + bytes) {
                bytes += (size_t)b->w[e2] * b->h[e2] * 4;
                e2++;
            }
        }
        if (e2 > i + 1)
            e = cudaMemcpyAsync(dst_of(i), b->himgs[i].out, bytes, cudaMemcpyDeviceToHost, ctx->d2h);
        else
            rc = download_on(b, i, dst_of(i), dstride_of(i), nullptr, ctx->d2h);
        i = e2;
    }
    if (!rc && e == cudaSuccess)
        e = cudaMemcpyAsync(b->hfilters, b->dfilters, b->filt_bytes, cudaMemcpyDeviceToHost, ctx->d2h);
    if (!rc && e == cudaSuccess) e = cudaEventRecord(b->ev_down, ctx->d2h);
    if (!rc && e != cudaSuccess)
        rc = set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "enqueue failed: %s", cudaGetErrorString(e));
    if (rc) {
        // nothing of this job may be left running on a batch we hand back
        cudaStreamSynchronize(ctx->h2d);
        cudaStreamSynchronize(b->stream);
        cudaStreamSynchronize(ctx->d2h);
        for (size_t i = 0; i < n; i++) images[i].status = rc;
        delete job;
        return rc;
    }
    b->busy = true;
    b->ran = true;
    ctx->inflight.push_back(job);
    *out = job;
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_wait(pngloss_b200_job *job) {
    if (!job) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = job->ctx;
    const int rc = finalize_job(job);
    ctx->inflight.erase(std::remove(ctx->inflight.begin(), ctx->inflight.end(), job), ctx->inflight.end());
    delete job;
    return rc;
}

extern "C" int pngloss_b200_optimize_batch(pngloss_b200_ctx *ctx, pngloss_b200_image *images, size_t n,
                                           unsigned strength, long bleed) {
    if (!ctx || !images || !n) return PNGLOSS_B200_INVALID_ARGUMENT;
    for (size_t i = 0; i < n; i++)
        if (!images[i].pixels) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "image %zu: NULL pixels", i);
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    // A batch larger than the free device memory is processed as consecutive groups that fit (the
    // reference has no such limit: it holds one image at a time, src/pngloss.c:173-205).
    size_t free_b = 0, total_b = 0, pooled = 0;
    PL_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    for (pngloss_b200_batch *b : ctx->pool)
        if (!b->busy) pooled += b->slab_bytes;
    size_t budget = (size_t)(0.92 * (double)(free_b + pooled));
    if (const char *e = getenv("PNGLOSS_B200_MEM_BUDGET_MB"))   // tests: force the grouping on small inputs
        budget = std::min(budget, (size_t)strtoull(e, nullptr, 10) << 20);
    size_t need_all = 0;
    for (size_t i = 0; i < n; i++)
        need_all += device_bytes_per_image(images[i].width, images[i].height, images[i].scanlines != nullptr);
    // everything at once if it fits; otherwise groups of half the budget, two in flight, so that the
    // copies of one group hide behind the kernels of the other
    const bool grouped = need_all > budget;
    if (grouped) budget /= 2;
    int first = 0;
    pngloss_b200_job *prev = nullptr;
    for (size_t begin = 0; begin < n;) {
        size_t end = begin, need = 0;
        while (end < n) {
            const size_t one = device_bytes_per_image(images[end].width, images[end].height,
                                                      images[end].scanlines != nullptr);
            if (end > begin && need + one > budget) break;
            need += one;
            end++;
        }
        pngloss_b200_job *job = nullptr;
        int rc = pngloss_b200_submit(ctx, images + begin, end - begin, strength, bleed, &job);
        if (prev) {
            const int prc = pngloss_b200_wait(prev);
            if (prc && !first) first = prc;
            prev = nullptr;
        }
        if (rc) {
            if (!first) first = rc;
        } else {
            prev = job;
        }
        begin = end;
    }
    if (prev) {
        const int prc = pngloss_b200_wait(prev);
        if (prc && !first) first = prc;
    }
    if (!grouped) trim_pool(ctx, 1);   // keep one allocation for a caller that repeats the batch
    return first;
}

// ---- drop-in entry points (reference src/pngloss_image.h) ----------------------------------------------
// The reference's entry points carry no context argument, so the library keeps one per calling thread: a context
// (its own stream - calls from different threads run side by side on the GPU, there is no global lock) and the
// device batch of the last call, which is reused when the next image has the same size (a command line that
// walks over a directory of equally sized files, the website's scaled previews) instead of allocating and
// freeing 2 x 4 bytes per pixel of device memory per call.
struct PlThreadDefault {
    pngloss_b200_ctx *ctx = nullptr;
    pngloss_b200_batch *batch = nullptr;
    uint32_t w = 0, h = 0;
    bool failed = false;
    ~PlThreadDefault() {
        // the main thread's copy dies at process exit, possibly after the CUDA runtime: then there is nothing to free
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) return;
        if (batch) pngloss_b200_batch_destroy(batch);
        if (ctx) pngloss_b200_ctx_destroy(ctx);
    }
};
static thread_local PlThreadDefault t_default;

static pngloss_b200_ctx *default_ctx() {
    PlThreadDefault &d = t_default;
    if (!d.ctx && !d.failed) {
        int dev = 0;
        if (const char *e = getenv("PNGLOSS_B200_DEVICE")) dev = atoi(e);
        if (pngloss_b200_ctx_create(&d.ctx, dev, nullptr) != PNGLOSS_B200_SUCCESS) {
            fprintf(stderr, "pngloss_b200: no usable CUDA device %d (there is no CPU fallback)\n", dev);
            d.ctx = nullptr;
            d.failed = true;
        }
    }
    return d.ctx;
}

static int optimize_rows_impl(unsigned char *const *rows, uint32_t width, uint32_t height,
                              unsigned char *row_filters, bool verbose, unsigned strength, long bleed,
                              uint32_t force_bpp = 0) {
    pngloss_b200_ctx *ctx = default_ctx();
    if (!ctx) return PNGLOSS_B200_DEVICE_ERROR;
    PlThreadDefault &d = t_default;
    if (d.batch && (d.w != width || d.h != height)) {
        pngloss_b200_batch_destroy(d.batch);
        d.batch = nullptr;
    }
    int rc = 0;
    if (!d.batch) {
        rc = pngloss_b200_batch_create(ctx, 1, &width, &height, &d.batch);
        if (rc) { d.batch = nullptr; return rc; }
        d.w = width;
        d.h = height;
    }
    pngloss_b200_batch *b = d.batch;
    int st = 0;
    rc = pngloss_b200_batch_set_mode(b, 0, row_filters == nullptr, force_bpp);
    if (!rc) rc = pngloss_b200_batch_upload_rows(b, 0, rows);
    if (!rc) rc = pngloss_b200_batch_run(b, strength, bleed);
    if (!rc) rc = pngloss_b200_batch_finish(b, &st, nullptr, nullptr);
    if (!rc) rc = pngloss_b200_batch_download_rows(b, 0, rows, row_filters);
    if (!rc) rc = pngloss_b200_ctx_sync(ctx);
    if (!rc && verbose) {
        // same two closing lines as the reference (src/pngloss_image.c:311-325), without the spinner
        uint32_t hist[256];
        unsigned used = 0;
        if (!pngloss_b200_batch_image_histogram(b, 0, hist))
            for (int i = 0; i < 256; i++) used += hist[i] != 0;
        fputs("  compression complete\n", stderr);
        fprintf(stderr, "  used %u unique symbols\n", used);
    }
    if (rc && rc != PNGLOSS_B200_OUT_OF_MEMORY)
        fprintf(stderr, "pngloss_b200: %s\n", pngloss_b200_ctx_error(ctx));
    if (rc) {   // do not keep a batch whose last run failed
        pngloss_b200_batch_destroy(b);
        d.batch = nullptr;
    }
    return rc;
}

extern "C" int optimize_with_rows(unsigned char **rows, uint32_t width, uint32_t height,
                                  unsigned char *row_filters, bool verbose,
                                  uint_fast8_t quantization_strength, int_fast16_t bleed_divider) {
    int rc = optimize_rows_impl(rows, width, height, row_filters, verbose, quantization_strength,
                                bleed_divider);
    if (rc == PNGLOSS_B200_NO_ACCEPTABLE_ROW) {
        // the reference prints and abort()s (src/pngloss_image.c:268-271)
        fprintf(stderr, "\naborting because no good row\n");
        abort();
    }
    return rc;
}

// reference src/pngloss_image.c:159.  The packed 1/2/3-byte pixels are widened to the RGBA layout the
// kernels work on and narrowed back, the mirror image of what the reference's optimize_with_rows does
// around this call (src/pngloss_image.c:105-147); the colour mode is forced instead of detected.
extern "C" int optimize_image(pngloss_image *image, unsigned char *row_filters, bool verbose,
                              uint_fast8_t quantization_strength, int_fast16_t bleed_divider) {
    if (!image || !image->rows || image->bytes_per_pixel < 1 || image->bytes_per_pixel > 4)
        return PNGLOSS_B200_INVALID_ARGUMENT;
    const uint32_t w = image->width, h = image->height, bpp = image->bytes_per_pixel;
    if (bpp == 4)
        return optimize_rows_impl(image->rows, w, h, row_filters, verbose, quantization_strength,
                                  bleed_divider, 4);
    std::vector<unsigned char> rgba;
    std::vector<unsigned char *> rows(h);
    try {
        rgba.resize((size_t)w * h * 4);
    } catch (const std::bad_alloc &) {
        return PNGLOSS_B200_OUT_OF_MEMORY;
    }
    for (uint32_t y = 0; y < h; y++) {
        rows[y] = rgba.data() + (size_t)y * w * 4;
        const unsigned char *src = image->rows[y];
        for (uint32_t x = 0; x < w; x++) {
            unsigned char *p = rows[y] + (size_t)x * 4;
            const unsigned char *q = src + (size_t)x * bpp;
            if (bpp == 3) { p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; p[3] = 255; }
            else { p[0] = p[1] = p[2] = q[0]; p[3] = bpp == 2 ? q[1] : 255; }
        }
    }
    int rc = optimize_rows_impl(rows.data(), w, h, row_filters, verbose, quantization_strength,
                                bleed_divider, bpp);
    if (rc == PNGLOSS_B200_NO_ACCEPTABLE_ROW) {
        fprintf(stderr, "\naborting because no good row\n");
        abort();
    }
    if (rc) return rc;
    for (uint32_t y = 0; y < h; y++) {
        unsigned char *dst = image->rows[y];
        for (uint32_t x = 0; x < w; x++) {
            const unsigned char *p = rows[y] + (size_t)x * 4;
            unsigned char *q = dst + (size_t)x * bpp;
            if (bpp == 3) { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; }
            else { q[0] = p[1]; if (bpp == 2) q[1] = p[3]; }
        }
    }
    return rc;
}

extern "C" void optimize_with_stride(unsigned char *pixels, uint32_t width, uint32_t height,
                                     uint32_t stride, bool verbose, uint_fast8_t quantization_strength,
                                     int_fast16_t bleed_divider) {
    std::vector<unsigned char *> rows(height);
    for (uint32_t i = 0; i < height; i++) rows[i] = pixels + (size_t)i * stride;
    optimize_with_rows(rows.data(), width, height, nullptr, verbose, quantization_strength, bleed_divider);
}

extern "C" void optimizeForAverageFilter(unsigned char pixels[], int width, int height, int quantization) {
    // "propagating half the color error is good middle ground" - bleed fixed at 2 (src/pngloss_image.c:35)
    optimize_with_stride(pixels, (uint32_t)width, (uint32_t)height, (uint32_t)width * 4u, false,
                         (uint_fast8_t)quantization, 2);
}

#include "pl_comm.cuh"
