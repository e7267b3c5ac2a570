// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT, never linked into libpngloss_b200.so.
//
// A tiny SIMT emulator: compiles the *same* kernel source that nvcc compiles for sm_100a with plain
// g++ and executes it on CPU fibers (one ucontext per CUDA thread, one CTA at a time), so that the
// kernel logic (lane mapping, warp collectives, barriers, tile/window bookkeeping) can be checked
// against the oracle in the GPU-less dev container before GPU minutes are spent.  It models:
//   * lock-step-free independent threads that only meet at collectives (worst case for missing
//     __syncwarp / __syncthreads: a fiber runs until it blocks, so unsynchronised cross-lane smem
//     traffic shows up as wrong results),
//   * cp.async as copies that land only at the matching wait (worst case for a missing wait),
//   * full and partial-mask warp shuffles / votes / redux, CTA barriers, smem and global atomics.
// It does not model timing, bank conflicts, or the memory model beyond the above.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

struct uchar4 { unsigned char x, y, z, w; };
struct short4 { short x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return uchar4{a, b, c, d}; }
static inline short4 make_short4(short a, short b, short c, short d) { return short4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

namespace simt {

struct Fiber;
// One rendezvous slot per member mask, so that disjoint lane groups of one warp can sit in different
// collectives at the same time (e.g. redux.sync over 8-lane groups with four different masks).
struct Slot {
    unsigned arrived = 0;
    unsigned gen = 0;
    unsigned long long val[2][32];
};
struct WarpState {
    std::map<unsigned, Slot> slots;
};
struct MbarCopy { void *dst; const void *src; unsigned n; };
struct MbarState {
    unsigned count = 0;      // arrivals per phase
    unsigned pending = 0;    // arrivals still missing in the current phase
    unsigned long long tx = 0;        // bytes announced by expect_tx in the current phase
    unsigned long long queued = 0;    // bytes of the bulk copies issued against the current phase
    unsigned phase = 0;
    std::vector<MbarCopy> copies;
};
struct BlockState {
    std::map<const void *, MbarState> mbars;
    unsigned nthreads = 0;
    unsigned bar_arrived = 0;
    unsigned bar_gen = 0;
    unsigned exited = 0;
    std::vector<WarpState> warps;
};
struct PendingCopy { void *dst; const void *src; unsigned n; };
struct Fiber {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    dim3 tid;
    bool done = false;
    std::vector<PendingCopy> pending;
};

extern Fiber *cur;
extern ucontext_t sched_ctx;
extern BlockState blk;
extern dim3 g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char *dyn_smem;
extern unsigned long long n_collectives;
extern unsigned long long progress;  // bumped on every arrival/completion; a pass without any = deadlock

static inline void yield() { swapcontext(&cur->ctx, &sched_ctx); }
static inline unsigned lane_id() { return cur->tid.x & 31u; }
static inline WarpState &my_warp() { return blk.warps[cur->tid.x >> 5]; }

// Generic warp rendezvous: every lane in `mask` deposits v, then all proceed with a snapshot.
static inline const unsigned long long *rendezvous(unsigned mask, unsigned long long v) {
    Slot &w = my_warp().slots[mask];
    unsigned lane = lane_id();
    if (!(mask >> lane & 1u)) { fprintf(stderr, "simt_emu: lane %u not in mask %08x\n", lane, mask); abort(); }
    unsigned g = w.gen;
    progress++;
    w.val[g & 1][lane] = v;
    w.arrived |= 1u << lane;
    if (w.arrived == mask) {
        w.arrived = 0;
        w.gen = g + 1;
        n_collectives++;
    } else {
        while (w.gen == g) yield();
    }
    return w.val[g & 1];
}

void launch(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem_bytes);

}  // namespace simt

#define threadIdx (simt::cur->tid)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
static const int warpSize = 32;

// ---- warp collectives ---------------------------------------------------------------------------
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl type");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned long long *all = simt::rendezvous(mask, raw);
    unsigned lane = simt::lane_id();
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = base + ((unsigned)src & (unsigned)(width - 1));
    T out;
    memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned long long *all = simt::rendezvous(mask, raw);
    unsigned lane = simt::lane_id();
    unsigned s = lane ^ (unsigned)lanemask;
    if ((s & ~(unsigned)(width - 1)) != (lane & ~(unsigned)(width - 1))) s = lane;
    T out;
    memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned long long *all = simt::rendezvous(mask, raw);
    unsigned lane = simt::lane_id();
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = (lane - base >= delta) ? lane - delta : lane;
    T out;
    memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned long long *all = simt::rendezvous(mask, raw);
    unsigned lane = simt::lane_id();
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = (lane - base + delta < (unsigned)width) ? lane + delta : lane;
    T out;
    memcpy(&out, &all[s], sizeof(T));
    return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned long long *all = simt::rendezvous(mask, pred ? 1ull : 0ull);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((mask >> i & 1u) && all[i]) r |= 1u << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::rendezvous(mask, 0); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    const unsigned long long *all = simt::rendezvous(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (mask >> i & 1u) r = (unsigned)all[i] > r ? (unsigned)all[i] : r;
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    const unsigned long long *all = simt::rendezvous(mask, v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++)
        if (mask >> i & 1u) r = (unsigned)all[i] < r ? (unsigned)all[i] : r;
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    const unsigned long long *all = simt::rendezvous(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (mask >> i & 1u) r += (unsigned)all[i];
    return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

static inline void __syncthreads() {
    simt::BlockState &b = simt::blk;
    unsigned g = b.bar_gen;
    simt::progress++;
    b.bar_arrived++;
    if (b.bar_arrived + b.exited == b.nthreads) {
        b.bar_arrived = 0;
        b.bar_gen = g + 1;
    } else {
        while (b.bar_gen == g) simt::yield();
    }
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- atomics (fibers are cooperative, so plain RMW is atomic) -------------------------------------
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; *p = o < v ? o : v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }

// ---- scalar intrinsics ----------------------------------------------------------------------------
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 7u;
        unsigned byte = (unsigned)((src >> (8 * sel)) & 0xffu);
        if ((s >> (4 * i)) & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;   // sign-replicate mode
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

// ---- cp.async model: copies land at the wait ---------------------------------------------------------
namespace simt {
static inline void cp_async(void *dst, const void *src, unsigned n) { cur->pending.push_back({dst, src, n}); }
static inline void cp_async_wait_all() {
    for (auto &c : cur->pending) memcpy(c.dst, c.src, c.n);
    cur->pending.clear();
}
}  // namespace simt

// ---- mbarrier + bulk copy model ---------------------------------------------------------------------------
// Default: the bytes of a bulk copy land at the last possible moment - when a waiter finds the phase
// complete (worst case for a consumer that reads before it waited).  With PL_EMU_BULK_EARLY=1 in the
// environment they land the moment the copy is issued (worst case for a producer that refills a buffer
// somebody still reads).
namespace simt {
static inline bool bulk_early() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("PL_EMU_BULK_EARLY"); v = e && *e == '1'; }
    return v != 0;
}
static inline MbarState &mbar_of(const void *bar) {
    auto it = blk.mbars.find(bar);
    if (it == blk.mbars.end()) { fprintf(stderr, "simt_emu: mbarrier %p used before init\n", bar); abort(); }
    return it->second;
}
static inline void mbar_init(const void *bar, unsigned count) {
    MbarState st;
    st.count = st.pending = count;
    blk.mbars[bar] = st;
    progress++;
}
static inline void mbar_arrive(const void *bar, unsigned tx_bytes) {
    MbarState &m = mbar_of(bar);
    if (!m.pending) { fprintf(stderr, "simt_emu: too many arrivals on mbarrier %p\n", bar); abort(); }
    m.pending--;
    m.tx += tx_bytes;
    progress++;
    // a phase without bulk copies completes with its last arrival, as on the hardware (phases that wait for
    // bytes complete when a waiter looks: the bytes land as late as possible)
    if (m.pending == 0 && m.tx == 0 && m.queued == 0) {
        m.pending = m.count;
        m.phase ^= 1u;
    }
}
static inline void bulk_copy(void *dst, const void *src, unsigned n, const void *bar) {
    if (n == 0 || (n & 15u) || ((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) {
        fprintf(stderr, "simt_emu: bulk copy needs 16-byte aligned dst/src/size (%p, %p, %u)\n", dst, src, n);
        abort();
    }
    MbarState &m = mbar_of(bar);
    m.queued += n;
    if (bulk_early()) memcpy(dst, src, n);
    else m.copies.push_back({dst, src, n});
    progress++;
}
// true once the phase of parity `parity` has completed
static inline bool mbar_test(const void *bar, unsigned parity) {
    MbarState &m = mbar_of(bar);
    if (m.pending == 0 && m.queued == m.tx) {
        for (auto &c : m.copies) memcpy(c.dst, c.src, c.n);
        m.copies.clear();
        m.pending = m.count;
        m.tx = m.queued = 0;
        m.phase ^= 1u;
        progress++;
    } else if (m.pending == 0 && m.queued > m.tx) {
        fprintf(stderr, "simt_emu: mbarrier %p got more bytes than expect_tx announced\n", bar);
        abort();
    }
    return (m.phase & 1u) != (parity & 1u);
}
}  // namespace simt
