// Device-side portability shim.
//
// The kernels in pl_kernels.cuh are written once.  nvcc compiles them for sm_100a (the product);
// tests/simt_emu compiles the very same text with g++ (-DPL_SIMT_EMU) to run kernel *logic* on CPU
// fibers in the GPU-less dev container.  Everything that needs inline PTX or a CUDA-only builtin
// is wrapped here so that the kernel bodies stay identical in both builds.
#pragma once

#ifdef PL_SIMT_EMU
#include "simt_emu.h"
#define PL_DYN_SMEM(name) unsigned char *name = simt::dyn_smem
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define PL_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define PL_FULL 0xffffffffu

// Path-coverage counters, only in the emulator build (tests assert that both the fast and the general
// variants of a code path were exercised); they compile to nothing on the device.
#ifdef PL_SIMT_EMU
extern unsigned long long pl_emu_counters[12];
#define PL_EMU_COUNT(slot) (pl_emu_counters[slot]++)
#else
#define PL_EMU_COUNT(slot) ((void)0)
#endif
enum {
    PL_CNT_TAPS_TABLE = 0, PL_CNT_TAPS_COMPUTED = 1, PL_CNT_FIXUP_REPLAY = 2, PL_CNT_FIXUP_SKIPPED = 3,
    // bucket-maxima variant: bytes answered by the look-up / by the scan, table updates by +1 / by max
    PL_CNT_BM_LOOKUP = 4, PL_CNT_BM_SCAN = 5, PL_CNT_BM_FASTUPD = 6, PL_CNT_BM_GENERAL = 7,
    // latency kernel: warp-pixels answered by the fast path / redone on the general path
    PL_CNT_SOLO_FAST = 8, PL_CNT_SOLO_GENERAL = 9
};

// ---- cp.async (LDGSTS): global -> shared without a register round trip ---------------------------
// 4- and 8-byte forms only exist as .ca; the sources are rows this CTA itself wrote (same SM, so
// L1 is coherent for them after the CTA barrier) or read-only input.
__device__ __forceinline__ void pl_cp_async4(void *smem_dst, const void *gmem_src) {
#ifdef PL_SIMT_EMU
    simt::cp_async(smem_dst, gmem_src, 4);
#else
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void pl_cp_async8(void *smem_dst, const void *gmem_src) {
#ifdef PL_SIMT_EMU
    simt::cp_async(smem_dst, gmem_src, 8);
#else
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void pl_cp_async_wait_all() {
#ifdef PL_SIMT_EMU
    simt::cp_async_wait_all();
#else
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// round a dynamic shared memory pointer up to `align` bytes (power of two) in the shared address space
__device__ __forceinline__ unsigned char *pl_align_shared(unsigned char *p, unsigned align) {
#ifdef PL_SIMT_EMU
    return (unsigned char *)(((uintptr_t)p + align - 1) & ~(uintptr_t)(align - 1));
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(p);
    return p + ((align - (s & (align - 1))) & (align - 1));
#endif
}

// ---- 2 KB-aligned shared-memory table of 256 64-bit entries -------------------------------------------
// The candidate scan of K2 reads entry ((first + k * step) mod 256) many times per pixel.  With the table
// 2 KB aligned the address is base | (byte_offset & 0x7f8): one LOP3 instead of a multiply-add chain.
#ifdef PL_SIMT_EMU
struct PlHkTable { const unsigned long long *p; };
__device__ __forceinline__ PlHkTable pl_hk_table(const unsigned long long *tab) { return PlHkTable{tab}; }
__device__ __forceinline__ unsigned long long pl_hk_load(PlHkTable t, unsigned byteoff) {
    return t.p[(byteoff & 0x7f8u) >> 3];
}
// caller guarantees byteoff < 2048 (no wrap-around): the address is base + offset, which ptxas folds
// into the load's immediate field when the offset is base + constant
__device__ __forceinline__ unsigned long long pl_hk_load_nowrap(PlHkTable t, unsigned byteoff) {
    if (byteoff >= 2048u) { fprintf(stderr, "simt_emu: histogram table overrun\n"); abort(); }
    return t.p[byteoff >> 3];
}
#else
struct PlHkTable { unsigned saddr; };
__device__ __forceinline__ PlHkTable pl_hk_table(const unsigned long long *tab) {
    PlHkTable t;
    t.saddr = (unsigned)__cvta_generic_to_shared(tab);
    return t;
}
__device__ __forceinline__ unsigned long long pl_hk_load(PlHkTable t, unsigned byteoff) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(t.saddr | (byteoff & 0x7f8u)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pl_hk_load_nowrap(PlHkTable t, unsigned byteoff) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(t.saddr + byteoff) : "memory");
    return v;
}
#endif

// ---- exact small-integer division by a runtime constant -------------------------------------------
// n / d for 0 <= n < 2^17, 1 <= d < 2^15, with magic = floor(2^32 / d) + 1 precomputed (d >= 2).
// Error bound: n * (magic*d - 2^32) <= n * d < 2^32, so the high word is exactly floor(n / d).
// d == 1 has no 32-bit magic; callers pass magic = 0 for it and we return n.
__device__ __forceinline__ unsigned pl_udiv_magic(unsigned n, unsigned magic) {
    return magic ? __umulhi(n, magic) : n;
}
__host__ __device__ __forceinline__ unsigned pl_make_magic(unsigned d) {
    return d <= 1 ? 0u : (unsigned)(0x100000000ull / d) + 1u;
}
// C-style signed division truncating toward zero, |n| < 2^17.
__device__ __forceinline__ int pl_sdiv_magic(int n, unsigned magic) {
    unsigned a = (unsigned)(n < 0 ? -n : n);
    int q = (int)pl_udiv_magic(a, magic);
    return n < 0 ? -q : q;
}
// n / 2^k truncating toward zero (C semantics for signed operands)
__device__ __forceinline__ int pl_sdiv_pow2(int n, int k) {
    return (n + ((n >> 31) & ((1 << k) - 1))) >> k;
}
// (2 * n) / 9 truncating toward zero for |n| < 2^16
__device__ __forceinline__ int pl_two_ninths(int n) {
    int m = 2 * n;
    unsigned a = (unsigned)(m < 0 ? -m : m);
    int q = (int)__umulhi(a, 477218589u);  // floor(2^32/9)+1
    return m < 0 ? -q : q;
}
__device__ __forceinline__ int pl_sext16(int v) { return (int)(short)v; }
// sign extension of the low byte
__device__ __forceinline__ int pl_sext8(int v) { return (int)(((unsigned)v & 255u) ^ 128u) - 128; }

// ---- mbarrier + bulk asynchronous copies (TMA unit, 1-D form: cp.async.bulk, SASS UBLKCP / SYNCS) ------
// K2's lean variant stages long row segments with these instead of per-lane cp.async: one instruction
// moves a whole segment global -> shared and reports its bytes to an mbarrier in shared memory.
// Source, destination and size must be multiples of 16 bytes.
#ifdef PL_SIMT_EMU
__device__ __forceinline__ void pl_mbar_init(unsigned long long *bar, unsigned count) { simt::mbar_init(bar, count); }
__device__ __forceinline__ void pl_mbar_arrive(unsigned long long *bar) { simt::mbar_arrive(bar, 0); }
__device__ __forceinline__ void pl_mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    simt::mbar_arrive(bar, bytes);
}
__device__ __forceinline__ void pl_mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!simt::mbar_test(bar, parity)) simt::yield();
}
__device__ __forceinline__ void pl_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes,
                                            unsigned long long *bar) {
    simt::bulk_copy(smem_dst, gmem_src, bytes, bar);
}
__device__ __forceinline__ void pl_fence_proxy_async() {}
__device__ __forceinline__ void pl_fence_mbar_init() {}
#else
__device__ __forceinline__ void pl_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void pl_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void pl_mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
// blocks until the phase of the given parity has completed (the first phase of a barrier has parity 0)
__device__ __forceinline__ void pl_mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PL_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"   /* suspend-time hint (ns): sleep, do not spin */
        "@p bra PL_MBAR_DONE_%=;\n"
        "bra PL_MBAR_WAIT_%=;\n"
        "PL_MBAR_DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void pl_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes,
                                            unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// orders this thread's earlier generic-proxy accesses (st.global / st.shared) before later accesses of the
// async proxy (bulk copies) to the same memory
__device__ __forceinline__ void pl_fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void pl_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
#endif

// ---- explicit shared-memory accesses by 32-bit shared address ---------------------------------------------------
// The latency kernel (pl_k2_solo.cuh) addresses its tables through these: a pointer that ptxas cannot prove to be
// shared costs a generic-to-shared conversion (S2R SR_CgaCtaId + LEA, ~25 cycles) on every use inside the loop.
#ifdef PL_SIMT_EMU
typedef unsigned char *PlSh;
__device__ __forceinline__ PlSh pl_sh(void *p) { return (unsigned char *)p; }
__device__ __forceinline__ PlSh pl_sh_opaque(PlSh a) { return a; }
__device__ __forceinline__ PlSh pl_sh_select(PlSh a, PlSh b, unsigned mask) { return mask ? a : b; }
__device__ __forceinline__ uint32_t pl_lds32(PlSh a) { return *(volatile uint32_t *)a; }
__device__ __forceinline__ unsigned long long pl_lds64(PlSh a) { return *(volatile unsigned long long *)a; }
__device__ __forceinline__ uint4 pl_lds128(PlSh a) {
    const volatile uint32_t *p = (const volatile uint32_t *)a;
    return make_uint4(p[0], p[1], p[2], p[3]);
}
__device__ __forceinline__ void pl_sts32(PlSh a, uint32_t v) { *(volatile uint32_t *)a = v; }
__device__ __forceinline__ void pl_atoms_inc32(PlSh a) { (*(volatile uint32_t *)a)++; }
__device__ __forceinline__ void pl_atoms_max32(PlSh a, uint32_t v) {
    if (v > *(volatile uint32_t *)a) *(volatile uint32_t *)a = v;
}
__device__ __forceinline__ void pl_atoms_add32(PlSh a, uint32_t v) { *(volatile uint32_t *)a += v; }
__device__ __forceinline__ uint32_t pl_atoms_add32_ret(PlSh a, uint32_t v) {
    const uint32_t o = *(volatile uint32_t *)a;
    *(volatile uint32_t *)a = o + v;
    return o;
}
#else
typedef unsigned PlSh;
__device__ __forceinline__ PlSh pl_sh(void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// hides where an address came from, so that it stays in its register instead of being re-derived where it is used
__device__ __forceinline__ PlSh pl_sh_opaque(PlSh a) {
    PlSh r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t pl_lds32(PlSh a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pl_lds64(PlSh a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 pl_lds128(PlSh a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void pl_sts32(PlSh a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void pl_atoms_inc32(PlSh a) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
}
__device__ __forceinline__ void pl_atoms_max32(PlSh a, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// a where mask is all ones, b where it is zero: one LOP3, no predicate
__device__ __forceinline__ PlSh pl_sh_select(PlSh a, PlSh b, unsigned mask) { return (a & mask) | (b & ~mask); }
// unconditional forms for the chain warp (a predicate would become a branch around the instruction): add 0 / max
// with 0 are no-ops
__device__ __forceinline__ void pl_atoms_add32(PlSh a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// ... returning the value before the addition
__device__ __forceinline__ uint32_t pl_atoms_add32_ret(PlSh a, uint32_t v) {
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}
#endif

// Polling wait for warps that are NOT on the critical path: test, then sleep.  (mbarrier.try_wait's suspend hint
// still re-issues ~90 times per 500 cycles - profiles/r2_solo_first_ncu.txt - and the arbiter prefers the higher
// warp id, so a spinning helper warp takes issue slots from the chain warp it shares a scheduler with.)
__device__ __forceinline__ void pl_mbar_wait_relaxed(unsigned long long *bar, unsigned parity) {
#ifdef PL_SIMT_EMU
    while (!simt::mbar_test(bar, parity)) simt::yield();
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(400);
    }
#endif
}
