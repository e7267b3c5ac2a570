// Multi-GPU part of the C ABI: the one collective the path has - the sum of the batch symbol histograms over
// all GPUs (256 x u64; the reference's per-image analogue is the "used N unique symbols" line of
// src/pngloss_image.c:315-325) - as an NCCL all-reduce issued by the library itself on the context's stream.
// NCCL is loaded at run time (dlopen "libnccl.so.2"; PNGLOSS_B200_NCCL_LIB overrides the path), so that the
// single-GPU product has no dependency on it.  Included by pl_api.cu.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct PlNccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

static PlNccl *pl_nccl() {
    static std::mutex mu;
    static PlNccl api;
    static bool tried = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        const char *path = getenv("PNGLOSS_B200_NCCL_LIB");
        void *h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) return nullptr;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
        if (api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce &&
            api.GetErrorString)
            api.handle = h;
    }
    return api.handle ? &api : nullptr;
}

#define PL_NCCL(ctx, api, call)                                                                   \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess)                                                                    \
            return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "%s failed: %s", #call, (api)->GetErrorString(r_)); \
    } while (0)

extern "C" int pngloss_b200_comm_unique_id(unsigned char id[PNGLOSS_B200_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == PNGLOSS_B200_COMM_ID_BYTES, "ncclUniqueId size");
    PlNccl *api = pl_nccl();
    if (!api || !id) return PNGLOSS_B200_DEVICE_ERROR;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return PNGLOSS_B200_DEVICE_ERROR;
    memcpy(id, &u, sizeof u);
    return PNGLOSS_B200_SUCCESS;
}

static int comm_scratch(pngloss_b200_ctx *ctx) {
    if (!ctx->comm_scratch) PL_CUDA(ctx, cudaMalloc((void **)&ctx->comm_scratch, 256 * sizeof(unsigned long long)));
    return PNGLOSS_B200_SUCCESS;
}

extern "C" int pngloss_b200_comm_init_rank(pngloss_b200_ctx *ctx, int nranks, int rank,
                                           const unsigned char id[PNGLOSS_B200_COMM_ID_BYTES]) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks || ctx->comm) return PNGLOSS_B200_INVALID_ARGUMENT;
    PlNccl *api = pl_nccl();
    if (!api) return set_err(ctx, PNGLOSS_B200_DEVICE_ERROR, "libnccl.so.2 not found (PNGLOSS_B200_NCCL_LIB)");
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    PL_NCCL(ctx, api, api->CommInitRank(&comm, nranks, u, rank));
    ctx->comm = comm;
    ctx->comm_ranks = nranks;
    ctx->comm_rank = rank;
    return comm_scratch(ctx);
}

extern "C" int pngloss_b200_comm_init_all(pngloss_b200_ctx **ctxs, int n) {
    if (!ctxs || n < 1 || n > 64) return PNGLOSS_B200_INVALID_ARGUMENT;
    for (int i = 0; i < n; i++)
        if (!ctxs[i] || ctxs[i]->comm) return PNGLOSS_B200_INVALID_ARGUMENT;
    PlNccl *api = pl_nccl();
    if (!api) return set_err(ctxs[0], PNGLOSS_B200_DEVICE_ERROR, "libnccl.so.2 not found (PNGLOSS_B200_NCCL_LIB)");
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    PL_NCCL(ctxs[0], api, api->CommInitAll(comms, n, devs));
    for (int i = 0; i < n; i++) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->comm_ranks = n;
        ctxs[i]->comm_rank = i;
        PL_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        if (int rc = comm_scratch(ctxs[i])) return rc;
    }
    return PNGLOSS_B200_SUCCESS;
}

extern "C" void pngloss_b200_comm_destroy(pngloss_b200_ctx *ctx) {
    if (!ctx || !ctx->comm) return;
    PlNccl *api = pl_nccl();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (api) api->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_ranks = 0;
}

extern "C" int pngloss_b200_comm_size(const pngloss_b200_ctx *ctx) { return ctx && ctx->comm ? ctx->comm_ranks : 0; }

// The collective: the batch's 256 x u64 symbol counts summed over every rank, in place on the device,
// enqueued behind the batch's kernels on the context's stream (asynchronous; read the result with
// pngloss_b200_batch_histogram after pngloss_b200_batch_finish).
extern "C" int pngloss_b200_batch_allreduce_histogram(pngloss_b200_batch *b) {
    if (!b) return PNGLOSS_B200_INVALID_ARGUMENT;
    pngloss_b200_ctx *ctx = b->ctx;
    if (!ctx->comm) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "no communicator on this context");
    if (!b->ran) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "allreduce_histogram before run");
    PlNccl *api = pl_nccl();
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_NCCL(ctx, api, api->AllReduce(b->batch_hist, b->batch_hist, 256, ncclUint64, ncclSum, (ncclComm_t)ctx->comm,
                                     b->stream));
    return PNGLOSS_B200_SUCCESS;
}

// Small host-side values reduced over the ranks (what a launcher needs around the path: agreeing on a batch
// size, the slowest rank's time, a barrier).  Blocking.  op: 0 sum, 1 max, 2 min.
extern "C" int pngloss_b200_comm_allreduce_u64(pngloss_b200_ctx *ctx, uint64_t *values, size_t n, int op) {
    if (!ctx || !values || n == 0 || n > 256 || op < 0 || op > 2) return PNGLOSS_B200_INVALID_ARGUMENT;
    if (!ctx->comm) return set_err(ctx, PNGLOSS_B200_INVALID_ARGUMENT, "no communicator on this context");
    PlNccl *api = pl_nccl();
    PL_CUDA(ctx, cudaSetDevice(ctx->device));
    PL_CUDA(ctx, cudaMemcpyAsync(ctx->comm_scratch, values, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    PL_NCCL(ctx, api,
            api->AllReduce(ctx->comm_scratch, ctx->comm_scratch, n, ncclUint64, op == 0 ? ncclSum : op == 1 ? ncclMax : ncclMin,
                           (ncclComm_t)ctx->comm, ctx->stream));
    PL_CUDA(ctx, cudaMemcpyAsync(values, ctx->comm_scratch, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PNGLOSS_B200_SUCCESS;
}

// Symbol counts (256 x u64) of every image this context's host-buffer calls (pngloss_b200_optimize_batch,
// pngloss_b200_submit / _wait) have finished so far - the batch-level form of the reference's "used N unique
// symbols" report (src/pngloss_image.c:315-325).  across_ranks != 0: summed over all ranks of the context's
// communicator by the NCCL all-reduce (blocking; every rank must call).
extern "C" int pngloss_b200_ctx_symbol_histogram(pngloss_b200_ctx *ctx, uint64_t out256[256], int across_ranks) {
    if (!ctx || !out256) return PNGLOSS_B200_INVALID_ARGUMENT;
    for (int k = 0; k < 256; k++) out256[k] = ctx->symbols[k];
    if (across_ranks && ctx->comm) return pngloss_b200_comm_allreduce_u64(ctx, out256, 256, 0);
    return PNGLOSS_B200_SUCCESS;
}
