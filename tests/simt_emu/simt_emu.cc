// TEST INFRASTRUCTURE - see simt_emu.h.
#include "simt_emu.h"

namespace simt {

Fiber *cur = nullptr;
ucontext_t sched_ctx;
BlockState blk;
dim3 g_blockIdx, g_blockDim, g_gridDim;
unsigned char *dyn_smem = nullptr;
unsigned long long n_collectives = 0;
unsigned long long progress = 0;

static const std::function<void()> *g_body = nullptr;

static void fiber_entry() {
    (*g_body)();
    cp_async_wait_all();
    cur->done = true;
    blk.exited++;
    // a thread that exits must not hold up a CTA barrier the others already reached
    if (blk.bar_arrived && blk.bar_arrived + blk.exited == blk.nthreads) {
        blk.bar_arrived = 0;
        blk.bar_gen++;
    }
    swapcontext(&cur->ctx, &sched_ctx);
}

void launch(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem_bytes) {
    const unsigned nthreads = block.x * block.y * block.z;
    static const size_t kStack = 256 * 1024;
    std::vector<unsigned char> smem(smem_bytes + 64);
    std::vector<Fiber> fibers(nthreads);
    for (auto &f : fibers) f.stack.resize(kStack);
    g_body = &body;
    g_blockDim = block;
    g_gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = dim3(bx, by, bz);
                memset(smem.data(), 0xA5, smem.size());  // poison: kernels must not assume zeroed smem
                dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
                blk = BlockState();
                blk.nthreads = nthreads;
                blk.warps.assign((nthreads + 31) / 32, WarpState());
                for (unsigned t = 0; t < nthreads; t++) {
                    Fiber &f = fibers[t];
                    f.done = false;
                    f.pending.clear();
                    f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack.data();
                    f.ctx.uc_stack.ss_size = f.stack.size();
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
                }
                unsigned live = nthreads;
                while (live) {
                    const unsigned long long before = progress;
                    for (unsigned t = 0; t < nthreads; t++) {
                        Fiber &f = fibers[t];
                        if (f.done) continue;
                        cur = &f;
                        swapcontext(&sched_ctx, &f.ctx);
                        if (f.done) { live--; progress++; }
                    }
                    if (progress == before) {
                        fprintf(stderr, "simt_emu: deadlock in block (%u,%u): %u threads blocked "
                                "(divergent collective or barrier)\n", bx, by, live);
                        abort();
                    }
                }
            }
    cur = nullptr;
    g_body = nullptr;
}

}  // namespace simt
