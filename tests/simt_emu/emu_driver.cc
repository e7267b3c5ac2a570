// TEST INFRASTRUCTURE - runs the product kernels (pngloss_b200/csrc/pl_kernels.cuh, unmodified text)
// on the CPU SIMT emulator.  Host memory plays the role of device memory.
#define PL_SIMT_EMU 1
#include "../../pngloss_b200/csrc/pl_kernels.cuh"

#include <vector>
unsigned long long pl_t_stats[16];
extern "C" void emu_tstats(unsigned long long *o){ for(int i=0;i<16;i++){o[i]=pl_t_stats[i]; pl_t_stats[i]=0;} }

template <int LPC>
static void run_k2(const PlImageDev *imgs, const int *slots, int nblocks, int strength, int bleed,
                   bool bm) {
    if (bm)
        simt::launch([&] { pl_k2_quantize<LPC, true>(imgs, slots, strength, bleed); },
                     dim3(nblocks), dim3(PL_K2_THREADS), sizeof(PlCtaSmem<LPC, true>) + PL_K2_SMEM_ALIGN);
    else
        simt::launch([&] { pl_k2_quantize<LPC, false>(imgs, slots, strength, bleed); },
                     dim3(nblocks), dim3(PL_K2_THREADS), sizeof(PlCtaSmem<LPC, false>) + PL_K2_SMEM_ALIGN);
}

extern "C" int emu_optimize(unsigned char *rgba, int n, uint32_t w, uint32_t h,
                            unsigned char *filters_out, int strength, int bleed, int adaptive_all,
                            int lpc, uint32_t *final_hist, uint32_t *status,
                            unsigned long long *batch_hist, uint32_t *chan_hist_out) {
    // lpc: lanes per channel (8, 4, 2, 1), + 16 for the bucket-maxima variant of K2, + 32 for an
    // in-place batch (the output buffer is the input buffer), + 64 for the lean kernel (pl_k2_lean, lpc 1)
    // + 128 for the latency kernel (pl_k2_solo<5>: one chain warp), + 256 for pl_k2_solo<1> (five chain warps)
    const bool bm = (lpc & 16) != 0, in_place = (lpc & 32) != 0, lean = (lpc & 64) != 0;
    const int solo = (lpc & 128) ? 5 : (lpc & 256) ? 1 : (lpc & 512) ? 4 : 0;   // + 512: the four-warp layout
    lpc &= 15;
    if (solo) lpc = 8;   // one image per CTA
    if (lean && (lpc != 1 || (w & 3))) return -2;
    const size_t npx = (size_t)w * h;
    const size_t ew = (size_t)w + PL_ERR_PAD;
    std::vector<uchar4> out(npx * n);
    std::vector<uint32_t> chan(5 * 4 * 256 * (size_t)n, 0), flags(2 * (size_t)n, 0);
    std::vector<short4> err(2 * 5 * 2 * ew * n);
    std::vector<uchar4> cand(5 * (size_t)w * n), oprev((size_t)w * n);
    memset(err.data(), 0x5A, err.size() * sizeof(short4));   // poison: kernel must not rely on zeros
    memset(cand.data(), 0x5A, cand.size() * sizeof(uchar4));
    memset(out.data(), 0x5A, out.size() * sizeof(uchar4));
    memset(oprev.data(), 0x5A, oprev.size() * sizeof(uchar4));
    std::vector<PlImageDev> imgs(n);
    for (int i = 0; i < n; i++) {
        PlImageDev &d = imgs[i];
        d.in = (const uchar4 *)rgba + npx * i;
        d.out = in_place ? (uchar4 *)rgba + npx * i : out.data() + npx * i;
        d.filters = filters_out + (size_t)h * i;
        d.chan_hist = chan.data() + 5 * 4 * 256 * (size_t)i;
        d.flags = flags.data() + 2 * (size_t)i;
        d.final_hist = final_hist + 256 * (size_t)i;
        d.err = err.data() + 2 * 5 * 2 * ew * i;
        d.cand = cand.data() + 5 * (size_t)w * i;
        d.oprev = oprev.data() + (size_t)w * i;
        d.status = status + 3 * (size_t)i;
        d.width = w;
        d.height = h;
        d.adaptive_all = adaptive_all > 1 ? (unsigned)((adaptive_all >> (2 + i % 8)) & 1) : (unsigned)adaptive_all;
        d.force_mode = 0;
    }
    const PlImageDev *dimgs = imgs.data();
    unsigned k1_slices = h < 4 ? h : 4;
    simt::launch([&] { pl_k1_orig_hist(dimgs, k1_slices); }, dim3(k1_slices * n), dim3(PL_K1_THREADS), 0);
    if (chan_hist_out) memcpy(chan_hist_out, chan.data(), chan.size() * sizeof(uint32_t));

    const int cpw = 8 / lpc;
    const int nblocks = (n + cpw - 1) / cpw;
    std::vector<int> slots((size_t)nblocks * cpw, -1);
    for (int i = 0; i < n; i++) slots[i] = i;
    if (solo) {
        const int *dslots = slots.data();
        if (solo == 5)
            simt::launch([&] { pl_k2_solo<5, false>(dimgs, dslots, strength, bleed, 2u); }, dim3(nblocks),
                         dim3(PlSoloCfg<5, false>::THREADS), sizeof(PlSoloSmem) + 16);
        else if (solo == 4)
            simt::launch([&] { pl_k2_solo<5, true>(dimgs, dslots, strength, bleed, 2u); }, dim3(nblocks),
                         dim3(PlSoloCfg<5, true>::THREADS), sizeof(PlSoloSmem) + 16);
        else
            simt::launch([&] { pl_k2_solo<1, false>(dimgs, dslots, strength, bleed, 2u); }, dim3(nblocks),
                         dim3(PlSoloCfg<1, false>::THREADS), sizeof(PlSoloSmem) + 16);
    } else if (lean) {
        const int *dslots = slots.data();
        simt::launch([&] { pl_k2_lean(dimgs, dslots, strength, bleed); }, dim3(nblocks), dim3(PL_K2_THREADS),
                     sizeof(PlLeanSmem) + PL_L_SMEM_ALIGN);
    } else
    switch (lpc) {
    case 8: run_k2<8>(dimgs, slots.data(), nblocks, strength, bleed, bm); break;
    case 4: run_k2<4>(dimgs, slots.data(), nblocks, strength, bleed, bm); break;
    case 2: run_k2<2>(dimgs, slots.data(), nblocks, strength, bleed, bm); break;
    case 1: run_k2<1>(dimgs, slots.data(), nblocks, strength, bleed, bm); break;
    default: return -1;
    }
    if (batch_hist) {
        memset(batch_hist, 0, 256 * sizeof(unsigned long long));
        simt::launch([&] { pl_k3_batch_hist(dimgs, n, batch_hist); }, dim3(2), dim3(256), 0);
    }
    if (!in_place) memcpy(rgba, out.data(), npx * n * sizeof(uchar4));
    return 0;
}

// K4 on the emulator: pixels = the quantised RGBA images, filters = K2's masks; scan receives, per image,
// height * (1 + 4 * width) bytes of room (the kernel uses height * (1 + bpp * width) of them).
extern "C" int emu_scanlines(const unsigned char *rgba, int n, uint32_t w, uint32_t h, const unsigned char *filters,
                             unsigned char *scan, uint32_t *oflags) {
    const size_t npx = (size_t)w * h, room = (size_t)h * (1 + 4 * (size_t)w);
    std::vector<PlScanDev> d(n);
    memset(oflags, 0, sizeof(uint32_t) * 4 * n);
    for (int i = 0; i < n; i++) {
        d[i].px = (const uchar4 *)rgba + npx * i;
        d[i].filters = filters + (size_t)h * i;
        d[i].scan = scan + room * i;
        d[i].oflags = oflags + 4 * (size_t)i;
        d[i].width = w;
        d[i].height = h;
    }
    const PlScanDev *dd = d.data();
    const unsigned slices = h < 3 ? h : 3;
    simt::launch([&] { pl_k4_scan_output(dd, slices); }, dim3(slices * n), dim3(PL_K4_THREADS), 0);
    simt::launch([&] { pl_k4_scanlines(dd, slices); }, dim3(slices * n), dim3(PL_K4_THREADS), 0);
    return 0;
}

extern "C" void emu_synth(unsigned char *dst, uint32_t w, uint32_t h, unsigned long long seed) {
    simt::launch([&] { pl_k_synth((uchar4 *)dst, w, h, seed); }, dim3(3), dim3(256), 0);
}

extern "C" unsigned long long emu_collectives() { return simt::n_collectives; }

unsigned long long pl_emu_counters[12] = {0};
extern "C" void emu_counters(unsigned long long *out) { memcpy(out, pl_emu_counters, sizeof pl_emu_counters); }
