// pngloss_b200 device kernels (sm_100a).
//
//   K1  pl_k1_orig_hist      per-channel 5-predictor histograms of the ORIGINAL image + the gray /
//                            opaque format scan.         replaces reference
//                            src/optimize_state.c:66-83 (optimize_state_init histogram loop) and
//                            src/pngloss_image.c:64-80 (format scan)
//   K2  pl_k2_quantize<LPC>  the row loop: 5 filter candidates per row, per-byte band quantiser with
//                            running symbol histogram, Sierra error diffusion, row cost, winner
//                            commit.                      replaces reference
//                            src/pngloss_image.c:159-309 (optimize_image) and
//                            src/optimize_state.c:114-361,390-562 (run/row/diffuse/adaptive)
//   K3  pl_k3_batch_hist     sum of the final symbol histograms of a batch (new; feeds the one
//                            cross-GPU all-reduce)
//       pl_k_synth           the synthetic gradient+noise generator of SURVEY 8d (bench inputs)
//
// All arithmetic is integer and must stay bit-exact with the reference; see DESIGN.md for the
// mapping and for why K2 is bound by a serial dependency chain rather than by HBM.
#pragma once
#include "pl_device.h"
#include "pl_types.h"

// --------------------------------------------------------------------------------------------------
// PNG predictors (reference src/optimize_state.c:575-613)
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pl_paeth(int above, int diag, int left) {
    int p = above - diag;
    int q = left - diag;
    int dl = p < 0 ? -p : p;
    int da = q < 0 ? -q : q;
    int dd = (p + q) < 0 ? -(p + q) : (p + q);
    return (dl <= da && dl <= dd) ? left : (da <= dd) ? above : diag;
}
template <int F>
__device__ __forceinline__ int pl_predict(int above, int diag, int left) {
    if (F == 1) return left;
    if (F == 2) return above;
    if (F == 3) return (above + left) >> 1;
    if (F == 4) return pl_paeth(above, diag, left);
    return 0;
}
// Runtime-filter form used by K2: every warp of a CTA runs the same instruction stream (one loop
// body in the instruction cache instead of five); F is warp-uniform so the paeth branch never diverges.
struct PlPredictor {
    int ma, ml, sh;
    bool paeth;
};
__device__ __forceinline__ PlPredictor pl_make_predictor(int F) {
    PlPredictor p;
    p.ma = (F == 2 || F == 3) ? 255 : 0;
    p.ml = (F == 1 || F == 3) ? 255 : 0;
    p.sh = (F == 3) ? 1 : 0;
    p.paeth = (F == 4);
    return p;
}
__device__ __forceinline__ int pl_predict_rt(const PlPredictor &p, int above, int diag, int left) {
    if (p.paeth) return pl_paeth(above, diag, left);
    return ((above & p.ma) + (left & p.ml)) >> p.sh;
}
__device__ __forceinline__ int pl_byte(unsigned v, int c) { return (int)((v >> (8 * c)) & 0xffu); }
__device__ __forceinline__ unsigned pl_u32(uchar4 p) {
    return (unsigned)p.x | ((unsigned)p.y << 8) | ((unsigned)p.z << 16) | ((unsigned)p.w << 24);
}
__device__ __forceinline__ uchar4 pl_uc4(unsigned v) {
    return make_uchar4((unsigned char)v, (unsigned char)(v >> 8), (unsigned char)(v >> 16),
                       (unsigned char)(v >> 24));
}
// |signed residual| as libpng's filter heuristic sums it (reference src/optimize_state.c:510-529)
__device__ __forceinline__ unsigned pl_absres(int here, int pred) {
    unsigned r = (unsigned)(here - pred) & 0xffu;
    return r < 128u ? r : 256u - r;
}

// --------------------------------------------------------------------------------------------------
// K1: original-image histograms + format scan.  grid = images * slices CTAs, 256 threads; CTA b
// takes rows (b % slices), (b % slices) + slices, ... of image b / slices.
// Algorithmic traffic: 4 B/pixel read (neighbours come from L1), 20 KB written per image.
// --------------------------------------------------------------------------------------------------
#define PL_K1_THREADS 256

__global__ void __launch_bounds__(PL_K1_THREADS) pl_k1_orig_hist(const PlImageDev *imgs,
                                                                  unsigned slices) {
    __shared__ uint32_t h[PL_FILTERS * 4 * 256];
    __shared__ uint32_t sflag[2];
    const PlImageDev im = imgs[blockIdx.x / slices];
    const uint32_t W = im.width, H = im.height;
    for (int i = threadIdx.x; i < PL_FILTERS * 4 * 256; i += PL_K1_THREADS) h[i] = 0;
    if (threadIdx.x < 2) sflag[threadIdx.x] = 0;
    __syncthreads();

    unsigned notgray = 0, notopaque = 0;
    for (uint32_t y = blockIdx.x % slices; y < H; y += slices) {
        const uchar4 *row = im.in + (size_t)y * W;
        const uchar4 *up = y ? row - W : row;
        for (uint32_t x = threadIdx.x; x < W; x += PL_K1_THREADS) {
            const unsigned o = pl_u32(row[x]);
            const unsigned l = x ? pl_u32(row[x - 1]) : 0u;
            const unsigned a = y ? pl_u32(up[x]) : 0u;
            const unsigned d = (x && y) ? pl_u32(up[x - 1]) : 0u;
            const int r = pl_byte(o, 0), g = pl_byte(o, 1), b = pl_byte(o, 2);
            notgray |= (unsigned)(r != g || g != b);
            notopaque |= (unsigned)(pl_byte(o, 3) < 255);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int oc = pl_byte(o, c), lc = pl_byte(l, c), ac = pl_byte(a, c), dc = pl_byte(d, c);
                atomicAdd(&h[(0 * 4 + c) * 256 + oc], 1u);
                atomicAdd(&h[(1 * 4 + c) * 256 + ((oc - lc) & 255)], 1u);
                atomicAdd(&h[(2 * 4 + c) * 256 + ((oc - ac) & 255)], 1u);
                atomicAdd(&h[(3 * 4 + c) * 256 + ((oc - ((ac + lc) >> 1)) & 255)], 1u);
                atomicAdd(&h[(4 * 4 + c) * 256 + ((oc - pl_paeth(ac, dc, lc)) & 255)], 1u);
            }
        }
    }
    if (notgray) sflag[0] = 1;
    if (notopaque) sflag[1] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < PL_FILTERS * 4 * 256; i += PL_K1_THREADS)
        if (h[i]) atomicAdd(&im.chan_hist[i], h[i]);
    if (threadIdx.x < 2 && sflag[threadIdx.x]) atomicOr(&im.flags[threadIdx.x], 1u);
}

// --------------------------------------------------------------------------------------------------
// K2: quantise + filter search.
//
// One CTA = 5 warps = the 5 filter candidates of CPW images (CPW = 8 / LPC).  Inside a warp a chain
// (image, filter) owns GROUP = 4*LPC lanes: LPC lanes for each RGBA channel.  The channels of one
// pixel are evaluated together against the histogram as it stood at the start of the pixel and then
// repaired in channel order (see "fix-up" below), which keeps the reference's strictly sequential
// symbol_frequency semantics while exposing 4-way parallelism.
// --------------------------------------------------------------------------------------------------
// Resident CTAs per SM the register allocation of K2 is sized for (5 warps each).  K2 is latency
// bound (a serial dependency chain per warp), so warps in flight matter more than a few spills.
// Shared memory already caps the narrow-lane variants at 3 (LPC 2) and 2 (LPC 1) CTAs per SM, so they
// may use more registers - spent on unrolling the candidate scan for instruction-level parallelism.
#ifndef PL_K2_MIN_BLOCKS
#define PL_K2_MIN_BLOCKS(LPC) ((LPC) >= 4 ? 4 : (LPC) == 2 ? 3 : 2)
#endif

template <int LPC>
struct PlCfg {
    static const int GROUP = 4 * LPC;    // lanes per chain
    static const int CPW = 32 / GROUP;   // chains per warp = images per CTA
    static const int TP = GROUP;         // pixels per tile per chain (one per lane)
    // unroll factor of the candidate scan: ceil((strength + 1) / LPC) candidates per lane, i.e. 3, 6,
    // 11, 21 at the usual strengths 19/20
    static const int UNR = LPC == 8 ? 3 : LPC == 4 ? 3 : 11;
    static const int NACC = LPC == 1 ? 4 : LPC == 2 ? 2 : 1;   // running maxima in the scan (measured best)
    static const int LOG2LPC = LPC == 8 ? 3 : LPC == 4 ? 2 : LPC == 2 ? 1 : 0;
    // narrow lane groups scan many candidates per lane: test the one "exact" candidate separately
    // instead of carrying its flag through every candidate
    static const bool EXSEP = LPC <= 2;
    // the Sierra tap table pays off where the kernel is instruction bound (wide lane groups, many CTAs per
    // SM: +8 %); with one or two warps per sub-partition its load latency costs more than it saves (-4 %)
    static const bool DLUT = LPC >= 4;
    // the narrow mappings have no room for that table (two CTAs of 112 KB per SM) and no taste for its vote;
    // they use a 1 KB table of 6-bit fields for |error| < 128 and fall back per lane
    static const bool DL32 = LPC <= 2;
    // bank stagger between the histograms of the chains of one warp (64-bit entries, rotation)
    static const int HROT = LPC == 2 ? 8 : LPC == 1 ? 4 : 0;
};

template <int LPC>
struct PlWarpSmem {
    // input tiles, double buffered; slot 0 of each tile is the last pixel of the previous tile
    // (one spare slot at the end of each: the pixel loop fetches one pixel ahead without a bound check)
    uint32_t orig[2][PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 2];  // original row y
    uint32_t oa[2][PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 2];    // original row y-1 ("old above")
    uint32_t na[2][PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 2];    // quantised row y-1 ("new above")
    short4 e0[2][PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 1];      // incoming error row 0, cells x+4
    short4 e1[2][PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 1];      // incoming error row 1, cells x+4
    // output staging of the current tile
    uint32_t back[PlCfg<LPC>::CPW][PlCfg<LPC>::TP + 1];     // candidate pixels (slot 0 = carry)
    short4 n0[PlCfg<LPC>::CPW][PlCfg<LPC>::TP];             // finished cells of next error row 0
    short4 n1[PlCfg<LPC>::CPW][PlCfg<LPC>::TP];             // finished cells of next error row 1
};

#define PL_K2_SMEM_ALIGN 2048
#define PL_DLUT_HALF 512   /* the Sierra tap table covers error values -512 .. 511 (see pl_pack_taps) */
#define PL_DL32_HALF 128   /* the small one -128 .. 127 (see pl_pack_taps6) */

// Bucket-maxima variant (BM): number of entries of PlCtaSmem::bmk per (chain, filter).  Buckets exist
// for strength + 1 >= 16 (see pl_bm_counts), which gives at most 128/16 + 1 + 129/16 + 1 = 18 of them.
#define PL_BM_MAX 18
#define PL_BM_MIN_STEP 16
// ... and for strength + 1 <= 127: up to there the buckets of one sign never meet around the 256-bin ring,
// which is what "a bin has at most two buckets" (pl_bm_counts) rests on.
#define PL_BM_MAX_STEP 127
// ... and for images narrower than this (relative counts of a row must fit the 17-bit field)
#define PL_BM_MAX_WIDTH 16384
#define PL_BM_COUNT_SHIFT 15

template <int LPC, bool BM = false>
struct PlCtaSmem {
    // Per chain and symbol one 64-bit entry that is already most of a candidate key: high word = the
    // running symbol_frequency, low word = rank of original_frequency[filter][symbol] << 10.
    // Must stay the first member: every 256-entry table is then 2 KB aligned (see pl_hk_load).
    // Chains that share a half-warp usually look at the same symbols (residuals cluster at 0), so the
    // table of chain ci is stored rotated by ci * HROT entries to land in different banks.
    unsigned long long hk[PlCfg<LPC>::CPW][PL_FILTERS * 256];
    unsigned long long dlut[PlCfg<LPC>::DLUT ? 2 * PL_DLUT_HALF : 1];   // packed Sierra taps by error value
    unsigned dl32[PlCfg<LPC>::DL32 ? 2 * PL_DL32_HALF : 1];              // the same, 5 x 6 bits (pl_pack_taps6)
    uint32_t base[PlCfg<LPC>::CPW][256];              // symbol_frequency at the start of the row
    // BM: per chain, filter and band ("bucket") of the symbol axis the candidate key of the symbol that
    // currently wins the choice inside that band (see pl_row_pass)
    unsigned long long bmk[BM ? PlCfg<LPC>::CPW : 1][BM ? PL_FILTERS : 1][BM ? PL_BM_MAX : 1];
    unsigned long long cost[PlCfg<LPC>::CPW][PL_FILTERS];
    PlImageDev img[PlCfg<LPC>::CPW];
    int win[PlCfg<LPC>::CPW];
    int retry;
    PlWarpSmem<LPC> w[PL_K2_WARPS];
};

// The Sierra tap values of one error lane (reference src/optimize_state.c:398-467): d = diff / bleed,
// then twos = d/16, threes = (d - 4 twos)/8, fours = 2 (d - ...)/9, five = (...)/2 and the remainder, all
// with C's truncating division.
struct PlTaps {
    int twos, threes, fours, five, rem;
};
__device__ __forceinline__ PlTaps pl_sierra_taps(int diff, unsigned bleed_magic) {
    PlTaps t;
    int dd = pl_sdiv_magic(diff, bleed_magic);
    t.twos = pl_sdiv_pow2(dd, 4);
    dd -= 4 * t.twos;
    t.threes = pl_sdiv_pow2(dd, 3);
    dd -= 2 * t.threes;
    t.fours = pl_two_ninths(dd);
    dd -= 2 * t.fours;
    t.five = pl_sdiv_pow2(dd, 1);
    t.rem = dd - t.five;
    return t;
}
// For |diff| < PL_DLUT_HALF all five values fit in signed bytes; K2 keeps them in a per-CTA table
// indexed by diff (the bleed divider is baked in), which replaces ~20 dependent integer operations
// per pixel by one shared-memory load.
__device__ __forceinline__ unsigned long long pl_pack_taps(const PlTaps &t) {
    return (unsigned long long)(unsigned char)t.twos | ((unsigned long long)(unsigned char)t.threes << 8) |
           ((unsigned long long)(unsigned char)t.fours << 16) | ((unsigned long long)(unsigned char)t.five << 24) |
           ((unsigned long long)(unsigned char)t.rem << 32);
}
__device__ __forceinline__ PlTaps pl_unpack_taps(unsigned long long e) {
    PlTaps t;
    const unsigned lo = (unsigned)e;
    t.twos = (int)(signed char)lo;
    t.threes = (int)(signed char)(lo >> 8);
    t.fours = (int)(signed char)(lo >> 16);
    t.five = (int)(signed char)(lo >> 24);
    t.rem = (int)(signed char)(unsigned)(e >> 32);
    return t;
}

// For |diff| < PL_DL32_HALF every value lies in [-32, 31] (|d| <= 127: twos <= 7, threes <= 12, fours <= 16,
// five <= 21, rem <= 22), so the five of them fit one 32-bit word as 6-bit fields biased by 32.
__device__ __forceinline__ unsigned pl_pack_taps6(const PlTaps &t) {
    return (unsigned)(t.twos + 32) | ((unsigned)(t.threes + 32) << 6) | ((unsigned)(t.fours + 32) << 12) |
           ((unsigned)(t.five + 32) << 18) | ((unsigned)(t.rem + 32) << 24);
}
__device__ __forceinline__ PlTaps pl_unpack_taps6(unsigned e) {
    PlTaps t;
    t.twos = (int)(e & 63u) - 32;
    t.threes = (int)((e >> 6) & 63u) - 32;
    t.fours = (int)((e >> 12) & 63u) - 32;
    t.five = (int)((e >> 18) & 63u) - 32;
    t.rem = (int)(e >> 24) - 32;
    return t;
}

// colour mode of an image: forced, or detected by K1's scan (reference src/pngloss_image.c:64-93)
__device__ __forceinline__ int pl_image_mode(const PlImageDev &im) {
    if (im.force_mode) return (int)im.force_mode;
    const bool notgray = im.flags[0] != 0, notopaque = im.flags[1] != 0;
    return notgray ? (notopaque ? 4 : 3) : (notopaque ? 2 : 1);
}

// per-lane view of its chain
struct PlChain {
    const uchar4 *in;
    const uchar4 *out;
    const uchar4 *oprev;
    short4 *err;
    uchar4 *cand;
    int chmask;       // active RGBA channels: 0x2 gray, 0xA gray+alpha, 0x7 rgb, 0xF rgba
    bool gray;        // gray modes count the G difference three times (color_delta.c:10-25)
    bool alpha_rule;  // fully transparent pixels stay transparent (optimize_state.c:158-164)
    bool live;        // chain takes part in this pass
    int q;            // quantisation strength of this pass
    unsigned step_magic;
};

__device__ __forceinline__ int pl_chan16(const short4 &v, int c) {
    return ((const short *)&v)[c];
}

// Candidate key, one 64-bit word whose plain unsigned maximum is the reference's choice
// (src/optimize_state.c:212-244):
//   [63:32] symbol_frequency of the candidate
//   [31:10] rank of original_frequency[filter][symbol] among the 256 symbols (order preserving, so
//           comparing ranks == comparing the 31-bit counts; built once per image in the prologue)
//   [9]     symbol == exact (unquantised) symbol
//   [8:0]   511 - offset of the symbol in its band: earlier candidates win ties, and keys are unique
// 0 = "no candidate" (every real key has a non-zero low field).
#define PL_KEY_RANK_SHIFT 10
__device__ __forceinline__ unsigned pl_key_low(bool exact, int pos) {
    return ((unsigned)exact << 9) | (unsigned)(511 - pos);
}
#define PL_KEY_RANK_MASK (~((1ull << PL_KEY_RANK_SHIFT) - 1ull))   /* count + rank fields */

// ---- bucket maxima (BM variant) ------------------------------------------------------------------------
// Before clamping, the band of admissible symbols of a byte (reference src/optimize_state.c:186-193) is
// one of a fixed set of intervals of the symbol axis: [k step, k step + q] for want >= 0 and
// [-k step - q, -k step] for want < 0 (step = q + 1).  Instead of scanning the q + 1 candidates of every
// byte, the BM variant keeps, per chain and filter, the candidate key of the current winner of every such
// interval ("bucket") that starts inside [-128, 127] and looks it up:
//   * table index: non-negative buckets k = 0 .. P1-1 at k, negative buckets k = 0 .. N1-1 at P1 + k;
//   * the last bucket of either sign reaches beyond +-128 and wraps around the 256-bin histogram, as the
//     scan does, so its far end shares bins with buckets of the other sign (the "seam");
//   * symbol 0 ends both zero bands; it is a member of bucket 0 only - negative bucket 0 covers [-q, -1]
//     and a look-up of the band [-q, 0] adds symbol 0 by hand.  So a bin outside the seam has exactly one
//     bucket, a bin in the seam two;
//   * a table entry is two 32-bit words: a per-row base count, and the winner's key relative to it,
//     [31:15] count - base, [14:7] rank of original_frequency (0..255), [6:0] 127 - position in the bucket
//     - the order of the candidate key below without the "exact" bit.  The base is the winner's count at
//     the start of the row minus 4 W: a symbol that starts the row further behind cannot win during it
//     (a row adds at most 4 W to any count), and no count climbs more than 8 W above the base, which is
//     why a relative count fits 17 bits for W < 16384 (PL_BM_MAX_WIDTH);
//   * symbol counts only grow, so "maximum of every key the bucket's symbols had during this row" is the
//     current winner: after every byte the chosen symbol's new key enters its bucket - its two buckets
//     in the seam - through one native 32-bit shared-memory atomicMax;
//   * a band clamped to the byte range is a sub-interval of its bucket: if the bucket winner lies
//     inside, it also wins the sub-interval; if not - or if the band starts beyond the last bucket - the
//     byte falls back to the candidate scan.
struct PlBm {
    int P1, N1;           // number of non-negative / negative buckets (0: no table at this strength)
    int seam_p, seam_n;   // bins >= seam_p also belong to the last negative bucket, bins <= seam_n to the
                          // last non-negative one
};
__device__ __forceinline__ PlBm pl_bm_counts(int step, int width) {
    PlBm b;
    const bool on = step >= PL_BM_MIN_STEP && step <= PL_BM_MAX_STEP && width < PL_BM_MAX_WIDTH;
    const int KP = 128 / step, KN = 129 / step;   // buckets that lie inside [-128, 127] entirely
    b.P1 = on ? KP + 1 : 0;
    b.N1 = on ? KN + 1 : 0;
    b.seam_p = on ? 256 - KN * step - (step - 1) : 999;
    b.seam_n = on ? KP * step + (step - 1) - 256 : -999;
    return b;
}
// first symbol of bucket t
__device__ __forceinline__ int pl_bm_low(const PlBm &b, int t, int step) {
    return t < b.P1 ? t * step : -(t - b.P1) * step - (step - 1);
}

// table key of a symbol: count relative to the bucket's base, rank, position in the bucket
__device__ __forceinline__ unsigned pl_bm_key(unsigned rel_count, unsigned rank, int pos) {
    return (rel_count << PL_BM_COUNT_SHIFT) | (rank << 7) | (unsigned)(127 - pos);
}

template <int LPC, bool BM>
__device__ __forceinline__ unsigned long long pl_row_pass(PlCtaSmem<LPC, BM> &sm, const PlChain &cn, int F,
                                                          int W, int y, int parity, int prev_w,
                                                          bool adaptive, unsigned bleed_magic) {
    typedef PlCfg<LPC> C;
    const int lane = threadIdx.x & 31;
    const int ci = lane / C::GROUP, gl = lane % C::GROUP;
    const int ch = gl / LPC, sub = gl % LPC;
    PlWarpSmem<LPC> &ws = sm.w[F];
    unsigned long long *hk = &sm.hk[ci][F * 256];
    const PlHkTable hkt = pl_hk_table(hk);
    const int rot = ci * C::HROT;        // physical entry of symbol s is (s + rot) & 255
    const int EW = W + PL_ERR_PAD;
    const bool live = cn.live;
    const bool first = (y == 0);
    const int chmask = cn.chmask;
    const bool alpha_rule = cn.alpha_rule;
    const unsigned step_magic = cn.step_magic;
    const bool act = live && ((chmask >> ch) & 1);
    const int q = cn.q, step = cn.q + 1;
    const int jmax = (q + LPC) / LPC;   // ceil((q + 1) / LPC)
    const PlPredictor predictor = pl_make_predictor(F);
    // byte selector that brings a pixel word into the mode's canonical form (byte 4 = zero):
    // rgba: as is; rgb: alpha -> 0; gray+alpha: G,G,G,A; gray: G,G,G,0
    const unsigned canon = chmask == 0xF ? 0x3210u : chmask == 0x7 ? 0x4210u : chmask == 0xA ? 0x3111u : 0x4111u;

    const short4 *Ecur0 = cn.err + ((size_t)(parity * PL_FILTERS + prev_w) * 2 + 0) * EW;
    const short4 *Ecur1 = Ecur0 + EW;
    short4 *En0 = cn.err + ((size_t)((parity ^ 1) * PL_FILTERS + F) * 2 + 0) * EW;
    short4 *En1 = En0 + EW;
    const uchar4 *rin = cn.in + (size_t)y * W;
    const uchar4 *rin_up = cn.oprev;   // original row y-1, saved by the commit of row y-1
    const uchar4 *rout_up = cn.out + (size_t)(y ? y - 1 : 0) * W;
    uchar4 *rcand = cn.cand + (size_t)F * W;

    // Sierra window for this lane's channel (reference src/optimize_state.c:445-467):
    //   a0..a2 = error row 0, cells x+2 (the one consumed by pixel x), x+3, x+4
    //   b0..b4 = error row 1, cells x .. x+4   (becomes row 0 of the next image row)
    //   c0..c3 = error row 2, cells x .. x+3   (becomes row 1 of the next image row; starts at zero)
    int a0 = 0, a1 = 0, a2, b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4, c0 = 0, c1 = 0, c2 = 0, c3;
    if (!first && live) {
        a0 = pl_chan16(Ecur0[2], ch);
        a1 = pl_chan16(Ecur0[3], ch);
        b0 = pl_chan16(Ecur1[0], ch);
        b1 = pl_chan16(Ecur1[1], ch);
        b2 = pl_chan16(Ecur1[2], ch);
        b3 = pl_chan16(Ecur1[3], ch);
    }
    int left = 0, aprev = 0;
    unsigned long long derr = 0;
    unsigned as0 = 0, as1 = 0, as2 = 0, as3 = 0, as4 = 0;

    // BM: the histogram of this candidate was just cloned from the previous row's winner (or restored
    // for a retry at another strength) and the tie-break rank differs per filter, so the bucket winners
    // are rebuilt at the start of every pass; the lanes of a chain share its buckets.
    const PlBm bmc = pl_bm_counts(BM ? step : 0, W);
    unsigned long long *bmrow = sm.bmk[BM ? ci : 0][BM ? F : 0];
    if (BM) {
        if (live) {
            for (int t = gl; t < bmc.P1 + bmc.N1; t += C::GROUP) {
                const int lo_t = pl_bm_low(bmc, t, step);
                const int np = t == bmc.P1 ? q : q + 1;   // negative bucket 0 leaves symbol 0 to bucket 0
                unsigned long long best = 0;
                for (int p = 0; p < np; p++) {
                    const unsigned long long key =
                        pl_hk_load(hkt, (unsigned)(lo_t + p + rot) * 8u) | (unsigned)(511 - p);
                    best = key > best ? key : best;
                }
                const unsigned mcount = (unsigned)(best >> 32);
                const unsigned base = mcount > 4u * (unsigned)W ? mcount - 4u * (unsigned)W : 0u;
                bmrow[t] = ((unsigned long long)base << 32) |
                           pl_bm_key(mcount - base, ((unsigned)best >> PL_KEY_RANK_SHIFT) & 255u,
                                     511 - (int)((unsigned)best & 511u));
            }
        }
        __syncwarp();
    }

    const int ntiles = (W + C::TP - 1) / C::TP;
    // tile loader: lane (ci, gl) fetches pixel x0 + gl of its chain
    auto issue = [&](int t, int buf) {
        const int x1 = t * C::TP + gl;
        if (x1 < W && live) {
            pl_cp_async4(&ws.orig[buf][ci][gl + 1], rin + x1);
            if (!first) {
                pl_cp_async4(&ws.oa[buf][ci][gl + 1], rin_up + x1);
                pl_cp_async4(&ws.na[buf][ci][gl + 1], rout_up + x1);
                pl_cp_async8(&ws.e0[buf][ci][gl], Ecur0 + x1 + 4);
                pl_cp_async8(&ws.e1[buf][ci][gl], Ecur1 + x1 + 4);
            } else {
                ws.oa[buf][ci][gl + 1] = 0;
                ws.na[buf][ci][gl + 1] = 0;
                ws.e0[buf][ci][gl] = make_short4(0, 0, 0, 0);
                ws.e1[buf][ci][gl] = make_short4(0, 0, 0, 0);
            }
        }
    };
    issue(0, 0);
    pl_cp_async_wait_all();
    __syncwarp();

    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        const int x0 = t * C::TP;
        const int npx = min(C::TP, W - x0);
        if (t + 1 < ntiles) issue(t + 1, buf ^ 1);

        // this lane's channel of the tile's first pixel; later pixels are fetched one iteration ahead
        int o_n = ((const unsigned char *)&ws.orig[buf][ci][1])[ch];
        int a_n = ((const unsigned char *)&ws.na[buf][ci][1])[ch];
        int e0_n = pl_chan16(ws.e0[buf][ci][0], ch);
        int e1_n = pl_chan16(ws.e1[buf][ci][0], ch);
        for (int i = 0; i < npx; i++) {
            const int o = o_n, a = a_n;
            a2 = e0_n;
            b4 = e1_n;
            c3 = 0;
            // prefetch pixel i+1, off the dependency chain (the slot after the tile is a spare)
            o_n = ((const unsigned char *)&ws.orig[buf][ci][i + 2])[ch];
            a_n = ((const unsigned char *)&ws.na[buf][ci][i + 2])[ch];
            e0_n = pl_chan16(ws.e0[buf][ci][i + 1], ch);
            e1_n = pl_chan16(ws.e1[buf][ci][i + 1], ch);
            int pred = pl_predict_rt(predictor, a, aprev, left);
            const bool transp = alpha_rule && ch == 3 && o == 0;

            // ---- band of admissible symbols (reference src/optimize_state.c:158-210) ---------
            // wrap (:175-182): bring the exact symbol orig - predicted into [-128, 127]
            int ex = o - pred;
            const int adj = ex < -128 ? -256 : (ex > 127 ? 256 : 0);
            pred += adj;
            ex -= adj;
            // band (:186-193): the multiple-of-(q+1) bucket that holds want, away from zero
            int here = o + pl_sext16(a0);
            const int want = here - pred;
            const unsigned m = (unsigned)(want < 0 ? -want : want);
            const unsigned kq = pl_udiv_magic(m, step_magic);   // which band, counted from zero
            const int r = (int)(m - kq * (unsigned)step);
            const int lo_u = want + (want < 0 ? r - q : -r);    // first symbol of the band before clamping
            // clamp (:195-210): symbol + predicted must be a byte.  The reference clamps lo from below
            // and hi from above and collapses an emptied band onto the saturated value, which is the
            // same as clamping both ends to [-predicted, 255 - predicted].
            const int smin = -pred, smax = 255 - pred;
            int lo = min(max(lo_u, smin), smax);
            int hi = min(max(lo_u + q, smin), smax);
            if (transp) {   // fully transparent stays transparent (:158-164): the only symbol is 0 - predicted
                here = 0;
                lo = hi = ex = smin;
            }
            const int span = act ? hi - lo : -1;

            // ---- candidate scan against the histogram at the start of the pixel ---------------
            // Every lane looks at candidates sub, sub + LPC, ...; the trip count comes from the
            // strength (warp-uniform), candidates beyond the clamped band are predicated off, and the
            // body is unrolled so that the independent shared-memory loads are in flight together.
            // NACC running maxima (merged after the scan) keep the compare chain short where a lane looks
            // at many candidates.
            auto scan = [&](unsigned long long(&acc)[C::NACC]) {
                const int jl = (span - sub) >> C::LOG2LPC;          // last valid j of this lane (-1: none)
                const unsigned off0 = (unsigned)(lo + sub + rot) * 8u;   // byte offset of candidate j = 0
                const unsigned low0 = 511u - (unsigned)sub;          // its "511 - pos" field
                const int jx = ex - lo - sub;                        // j * LPC of the exact symbol, if mine
                // (BM: only this lane's own candidates - the scan runs diverged there anyway and a clamped
                // band is usually short)
                for (int j0 = 0; j0 < (BM ? jl + 1 : jmax); j0 += C::UNR) {
                    // per-chunk bases, so that each candidate below only adds compile-time constants
                    const unsigned off_c = off0 + (unsigned)(j0 * LPC * 8);
                    const unsigned low_c = low0 - (unsigned)(j0 * LPC);
                    const int jl_c = jl - j0, jx_c = jx - j0 * LPC;
#pragma unroll
                    for (int u = 0; u < C::UNR; u++) {
                        unsigned long long key = pl_hk_load(hkt, off_c + (unsigned)(u * LPC * 8));
                        key |= low_c - (unsigned)(u * LPC);          // low 10 bits of an entry are zero
                        if (!C::EXSEP) key |= (unsigned)(u * LPC == jx_c) << 9;
                        unsigned long long &a = acc[u % C::NACC];
                        a = (u <= jl_c && key > a) ? key : a;
                    }
                }
            };
            // the exact symbol as a candidate of its own, with its bonus bit; every lane of the group may
            // add it (max is idempotent)
            auto exact_candidate = [&](unsigned long long &a) {
                const int pos = ex - lo;
                const unsigned long long key =
                    pl_hk_load(hkt, (unsigned)(ex + rot) * 8u) | pl_key_low(true, pos & 255);
                a = (pos >= 0 && pos <= span && key > a) ? key : a;
            };

            unsigned long long acc0;
            int tl = -1;                 // BM: bucket of the unclamped band (-1: beyond the table)
            unsigned base_l = 0;         // BM: its base count
            if (BM) {
                // ---- look the band's winner up ------------------------------------------------------
                const bool neg = want < 0;
                const bool tvalid = kq < (unsigned)(neg ? bmc.N1 : bmc.P1);
                tl = tvalid ? (int)kq + (neg ? bmc.P1 : 0) : -1;
                const unsigned long long e = *(volatile unsigned long long *)&bmrow[tvalid ? tl : 0];
                const unsigned k32 = (unsigned)e;
                base_l = (unsigned)(e >> 32);
                // the winner as a candidate key, position field still relative to the bucket (= to lo_u)
                unsigned long long wk =
                    ((unsigned long long)(base_l + (k32 >> PL_BM_COUNT_SHIFT)) << 32) |
                    (((k32 >> 7) & 255u) << PL_KEY_RANK_SHIFT) | ((k32 & 127u) + 384u);
                // band [-q, 0]: symbol 0 lives in bucket 0, add it by hand (unless the clamp cut it off)
                if (neg && kq == 0 && (unsigned)(-lo) <= (unsigned)span) {
                    const unsigned long long k0 = pl_hk_load(hkt, (unsigned)rot * 8u) | (unsigned)(511 - q);
                    wk = k0 > wk ? k0 : wk;
                }
                const int wsym = lo_u + 511 - (int)((unsigned)wk & 511u);
                const bool inr = tvalid && wsym >= lo && wsym <= hi;
                // position field relative to the clamped band: 511 - (wsym - lo)
                acc0 = act && inr ? wk + (unsigned long long)(unsigned)(lo - lo_u) : 0ull;
                // a band of one symbol that is the exact symbol is covered by the exact candidate below
                const bool need_scan = act && !inr && !(span == 0 && ex == lo);
#ifdef PL_SIMT_EMU
                // path statistics, per warp iteration: does any lane of the warp scan?
                if (__any_sync(PL_FULL, need_scan)) PL_EMU_COUNT(PL_CNT_BM_SCAN);
                else PL_EMU_COUNT(PL_CNT_BM_LOOKUP);
#endif
                // Only the bytes whose look-up failed scan; need_scan is the same in all lanes of a channel
                // group and the block contains no warp-level operation.
                if (need_scan) {
                    unsigned long long acc[C::NACC];
#pragma unroll
                    for (int k = 0; k < C::NACC; k++) acc[k] = 0;
                    scan(acc);
#pragma unroll
                    for (int k = C::NACC / 2; k >= 1; k >>= 1)
#pragma unroll
                        for (int m = 0; m < k; m++) acc[m] = acc[m + k] > acc[m] ? acc[m + k] : acc[m];
                    acc0 = acc[0];
                }
                exact_candidate(acc0);
            } else {
                unsigned long long acc[C::NACC];
#pragma unroll
                for (int k = 0; k < C::NACC; k++) acc[k] = 0;
                scan(acc);
                if (C::EXSEP) exact_candidate(acc[C::NACC - 1]);
#pragma unroll
                for (int k = C::NACC / 2; k >= 1; k >>= 1)
#pragma unroll
                    for (int m = 0; m < k; m++) acc[m] = acc[m + k] > acc[m] ? acc[m + k] : acc[m];
                acc0 = acc[0];
            }
            unsigned long long bkey = acc0;
#pragma unroll
            for (int mk = 1; mk < LPC; mk <<= 1) {
                const unsigned long long okey = __shfl_xor_sync(PL_FULL, bkey, mk);
                bkey = okey > bkey ? okey : bkey;
            }
            int bpos = 511 - (int)((unsigned)bkey & 511u);

            // ---- fix-up: replay the channel order ---------------------------------------------
            // The reference increments symbol_frequency[best] before the next channel looks at it
            // (:253).  Raising one count can only promote that one symbol, so the true winner of
            // channel k is either its provisional winner or one of the symbols chosen by channels
            // 0..k-1 (with their counts as updated so far).
            //
            // Fast path: a symbol v chosen by an earlier channel can only overtake this channel's
            // winner if v lies in this band, is not the winner itself, and count[v] + (at most 3
            // increments) reaches the winner's count.  If that holds nowhere in the warp, every
            // provisional winner is final (induction over the channel order).  The test needs only
            // provisional values, so its three steps are independent of each other.
            bool conflict = false;
            unsigned dup = 0;   // BM: earlier channels whose provisional winner is this channel's
            {
                // one word per channel group: winner's count (saturated to 24 bits, which only makes the
                // test more conservative) and its symbol
                const unsigned bf = min((unsigned)(bkey >> 32), 0xfffff0u);
                const unsigned mine = (bf << 8) | ((unsigned)(lo + bpos) & 255u);
#pragma unroll
                for (int t2 = 0; t2 < 3; t2++) {
                    const unsigned theirs = __shfl_sync(PL_FULL, mine, ci * C::GROUP + t2 * LPC);
                    const int pos = ((int)(theirs & 255u) - lo) & 255;
                    const bool earlier = (ch > t2) & act & (bool)((chmask >> t2) & 1);
                    conflict |= earlier & (pos <= span) & (pos != bpos) & ((theirs >> 8) + 3u >= bf);
                    if (BM) dup += (unsigned)(earlier & (pos == bpos));
                }
            }
            if (!__any_sync(PL_FULL, conflict)) {
                PL_EMU_COUNT(PL_CNT_FIXUP_SKIPPED);
                // every provisional winner is final; this channel's symbol has been counted dup times
                // since the look-up
                if (BM) bkey += (unsigned long long)dup << 32;
            } else {
                PL_EMU_COUNT(PL_CNT_FIXUP_REPLAY);
                // Slow path (rare once counts have spread): the exact sequential replay.
#pragma unroll 1
                for (int t2 = 0; t2 < 3; t2++) {
                    const int src = ci * C::GROUP + t2 * LPC;
                    const int vsym = __shfl_sync(PL_FULL, lo + bpos, src);
                    const unsigned long long kv = __shfl_sync(PL_FULL, bkey, src);
                    if (ch > t2 && act && ((chmask >> t2) & 1)) {
                        const int pos = (vsym - lo) & 255;
                        if (pos <= span) {
                            if (pos == bpos) {
                                bkey += 1ull << 32;          // my own winner was chosen again: count + 1
                            } else {
                                const unsigned long long key =
                                    ((kv & PL_KEY_RANK_MASK) + (1ull << 32)) | pl_key_low(lo + pos == ex, pos);
                                if (key > bkey) {
                                    bkey = key;
                                    bpos = pos;
                                }
                            }
                        }
                    }
                }
            }

            // ---- commit the byte ----------------------------------------------------------------
            const int sym = lo + bpos;
            const int back = act ? sym + pred : 0;
            if (act && sub == 0) {
                unsigned *cnt = (unsigned *)&hk[(unsigned)(sym + rot) & 255u] + 1;   // high word = count
                atomicAdd(cnt, 1u);
                if (BM && bmc.P1 > 0) {
                    // the symbol's new key enters its bucket - in the seam: its two buckets.  The count
                    // field of bkey is the symbol's count just before this channel's increment (the replay
                    // and the dup correction above keep it so).
                    const unsigned now = (unsigned)(bkey >> 32) + 1u;
                    const unsigned rank7 = ((unsigned)bkey >> (PL_KEY_RANK_SHIFT - 7)) & (255u << 7);   // rank << 7
                    const int s8 = (int)(signed char)sym;   // the histogram bin, as a symbol in [-128, 127]
                    const unsigned as8 = (unsigned)(s8 < 0 ? -s8 : s8);
                    const unsigned ks = pl_udiv_magic(as8, step_magic);
                    const int rs = (int)(as8 - ks * (unsigned)step);            // |s8| mod step
                    const int t1 = (int)ks + (s8 < 0 ? bmc.P1 : 0);
                    unsigned *e1 = (unsigned *)&bmrow[t1];
                    const unsigned base1 = t1 == tl ? base_l : *(volatile unsigned *)(e1 + 1);
                    // position in the bucket: rs from its first symbol (s8 >= 0), q - rs (s8 < 0)
                    const unsigned key1 = ((now - base1) << PL_BM_COUNT_SHIFT) | rank7 |
                                          (unsigned)(127 - (s8 < 0 ? q - rs : rs));
                    if (now >= base1) atomicMax(e1, key1);
                    if (s8 >= bmc.seam_p || s8 <= bmc.seam_n) {
                        PL_EMU_COUNT(PL_CNT_BM_GENERAL);
                        // bins >= seam_p: also symbol s8 - 256 of the last negative bucket;
                        // bins <= seam_n: also symbol s8 + 256 of the last non-negative bucket
                        const bool up = s8 <= bmc.seam_n;
                        const int t2 = up ? bmc.P1 - 1 : bmc.P1 + bmc.N1 - 1;
                        unsigned *e2 = (unsigned *)&bmrow[t2];
                        const unsigned base2 = *(volatile unsigned *)(e2 + 1);
                        if (now >= base2)
                            atomicMax(e2, ((now - base2) << PL_BM_COUNT_SHIFT) | rank7 |
                                              (unsigned)(127 - (s8 + (up ? 256 : -256) - pl_bm_low(bmc, t2, step))));
                    } else {
                        PL_EMU_COUNT(PL_CNT_BM_FASTUPD);
                    }
                }
                ((unsigned char *)&ws.back[ci][i + 1])[ch] = (unsigned char)back;
            }
            left = back;
            aprev = a;

            // ---- Sierra diffusion of (here - back) / bleed (reference :390-467) ------------------
            const int diff = act ? pl_sext16(here - back) : 0;
            PlTaps tp;
            if (C::DLUT && __all_sync(PL_FULL, (unsigned)(diff + PL_DLUT_HALF) < 2u * PL_DLUT_HALF)) {
                PL_EMU_COUNT(PL_CNT_TAPS_TABLE);
                tp = pl_unpack_taps(sm.dlut[diff + PL_DLUT_HALF]);
            } else if (C::DL32 && (unsigned)(diff + PL_DL32_HALF) < 2u * PL_DL32_HALF) {   // per lane
                PL_EMU_COUNT(PL_CNT_TAPS_TABLE);
                tp = pl_unpack_taps6(sm.dl32[diff + PL_DL32_HALF]);
            } else {
                PL_EMU_COUNT(PL_CNT_TAPS_COMPUTED);
                tp = pl_sierra_taps(diff, bleed_magic);
            }
            const int twos = tp.twos, threes = tp.threes, fours = tp.fours, five = tp.five, dd = tp.rem;
            a1 += dd;
            a2 += threes;
            b0 += twos;
            b1 += fours;
            b2 += five;
            b3 += fours;
            b4 += twos;
            c1 += twos;
            c2 += threes;
            c3 += twos;
            if (sub == 0) {
                ((short *)&ws.n0[ci][i])[ch] = (short)b0;
                ((short *)&ws.n1[ci][i])[ch] = (short)c0;
            }
            a0 = a1; a1 = a2;
            b0 = b1; b1 = b2; b2 = b3; b3 = b4;
            c0 = c1; c1 = c2; c2 = c3;
            __syncwarp();
        }

        // ---- tile epilogue: lane (ci, gl) owns pixel x0+gl ------------------------------------------
        if (gl < npx && live) {
            const int x = x0 + gl;
            const unsigned o4 = ws.orig[buf][ci][gl + 1], q4 = ws.back[ci][gl + 1];
            const unsigned oa4 = ws.oa[buf][ci][gl + 1], na4 = ws.na[buf][ci][gl + 1];
            unsigned ol4 = 0, ql4 = 0, oad4 = 0, nad4 = 0;
            if (x > 0) {
                ol4 = ws.orig[buf][ci][gl];
                ql4 = ws.back[ci][gl];
                oad4 = ws.oa[buf][ci][gl];
                nad4 = ws.na[buf][ci][gl];
            }
            En0[x] = ws.n0[ci][gl];
            En1[x] = ws.n1[ci][gl];
            rcand[x] = pl_uc4(q4);
            // derivative error of the three neighbours (reference :265-287): for N in {above, diag, left}
            //   sum over colour lanes of ((oldN - orig) - (newN - back))^2 .
            // With the four channels of a pixel in one 32-bit word that is a handful of byte dot products
            // (IDP.4A).  canon() zeroes channels the colour mode does not use and copies G over R and B in
            // the gray modes, which is how color_difference() counts gray three times (color_delta.c:10-25).
            {
                const unsigned o = __byte_perm(o4, 0u, canon), q = __byte_perm(q4, 0u, canon);
                const unsigned n1o = __byte_perm(oa4, 0u, canon), n1n = __byte_perm(na4, 0u, canon);
                const unsigned n2o = __byte_perm(oad4, 0u, canon), n2n = __byte_perm(nad4, 0u, canon);
                const unsigned n3o = __byte_perm(ol4, 0u, canon), n3n = __byte_perm(ql4, 0u, canon);
                // sum (a - b - c + d)^2 = X - 2Y + 2Z over byte dot products
                unsigned xs = __dp4a(o, o, __dp4a(q, q, 0u)) * 3u;
                xs = __dp4a(n1o, n1o, __dp4a(n1n, n1n, xs));
                xs = __dp4a(n2o, n2o, __dp4a(n2n, n2n, xs));
                xs = __dp4a(n3o, n3o, __dp4a(n3n, n3n, xs));
                unsigned ys = __dp4a(o, q, 0u) * 3u;
                ys = __dp4a(n1o, n1n, __dp4a(n1o, o, __dp4a(n1n, q, ys)));
                ys = __dp4a(n2o, n2n, __dp4a(n2o, o, __dp4a(n2n, q, ys)));
                ys = __dp4a(n3o, n3n, __dp4a(n3o, o, __dp4a(n3n, q, ys)));
                unsigned zs = __dp4a(n1o, q, __dp4a(n1n, o, 0u));
                zs = __dp4a(n2o, q, __dp4a(n2n, o, zs));
                zs = __dp4a(n3o, q, __dp4a(n3n, o, zs));
                derr += xs + 2u * zs - 2u * ys;
            }
            if (adaptive) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if ((chmask >> c) & 1) {
                        const int qc = pl_byte(q4, c);
                        const int lq = pl_byte(ql4, c), aq = pl_byte(na4, c), dq = pl_byte(nad4, c);
                        as0 += pl_absres(qc, 0);
                        as1 += pl_absres(qc, lq);
                        as2 += pl_absres(qc, aq);
                        as3 += pl_absres(qc, (aq + lq) >> 1);
                        as4 += pl_absres(qc, pl_paeth(aq, dq, lq));
                    }
                }
            }
        }
        if (gl == 0 && live) {
            ws.orig[buf ^ 1][ci][0] = ws.orig[buf][ci][npx];
            ws.oa[buf ^ 1][ci][0] = ws.oa[buf][ci][npx];
            ws.na[buf ^ 1][ci][0] = ws.na[buf][ci][npx];
            ws.back[ci][0] = ws.back[ci][npx];
        }
        pl_cp_async_wait_all();
        __syncwarp();
    }

    // ---- row tail: cells W .. W+3 of the two outgoing error rows -----------------------------------
    if (act && sub == 0) {
        ((short *)&En0[W + 0])[ch] = (short)b0;
        ((short *)&En0[W + 1])[ch] = (short)b1;
        ((short *)&En0[W + 2])[ch] = (short)b2;
        ((short *)&En0[W + 3])[ch] = (short)b3;
        ((short *)&En1[W + 0])[ch] = (short)c0;
        ((short *)&En1[W + 1])[ch] = (short)c1;
        ((short *)&En1[W + 2])[ch] = (short)c2;
        ((short *)&En1[W + 3])[ch] = 0;
    }

    // ---- row cost (reference src/optimize_state.c:314-360) -------------------------------------------
    // bits = sum over symbols of (count gained in this row) * ulog2(UINTMAX_MAX / final count);
    // ulog2(UINTMAX_MAX / f) == 33 + clz32(f) for every f >= 1.
    unsigned bits = 0;
    for (int s = gl; s < 256; s += C::GROUP) {
        const unsigned hv = (unsigned)(hk[s] >> 32);
        bits += (hv - sm.base[ci][s]) * (33u + (unsigned)__clz((int)hv));
    }
#pragma unroll
    for (int mk = 1; mk < C::GROUP; mk <<= 1) {
        derr += __shfl_xor_sync(PL_FULL, derr, mk);
        bits += __shfl_xor_sync(PL_FULL, bits, mk);
    }
    unsigned long long cost = derr / 128ull + bits;
    if (__any_sync(PL_FULL, adaptive)) {   // chains of one warp may differ; keep the shuffles uniform
#pragma unroll
        for (int mk = 1; mk < C::GROUP; mk <<= 1) {
            as0 += __shfl_xor_sync(PL_FULL, as0, mk);
            as1 += __shfl_xor_sync(PL_FULL, as1, mk);
            as2 += __shfl_xor_sync(PL_FULL, as2, mk);
            as3 += __shfl_xor_sync(PL_FULL, as3, mk);
            as4 += __shfl_xor_sync(PL_FULL, as4, mk);
        }
        // libpng picks the first minimum in the order none, sub, up, average, paeth (:531-559)
        unsigned lowest = min(min(min(as0, as1), min(as2, as3)), as4);
        int pick = lowest >= as0 ? 0 : lowest >= as1 ? 1 : lowest >= as2 ? 2 : lowest >= as3 ? 3 : 4;
        if (adaptive && pick != F) cost = ~0ull;
    }
    return cost;
}

// Row commit: the winner's candidate row becomes row y of the output (which may be the original row
// itself, in-place batch) and the original row y is kept in the image's scratch row for the predictions of
// row y + 1.  The compiler must assume that the stores alias the next loads, so every round of this loop
// costs a full memory latency: four pixels per thread and round where the row allows it.
__device__ __forceinline__ void pl_commit_row(const PlImageDev &im, int w2, int y, int W, int tid) {
    const int mode2 = pl_image_mode(im);
    const bool notgray = mode2 >= 3, notopaque = (mode2 & 1) == 0;
    const uchar4 *src = im.cand + (size_t)w2 * W;
    const uchar4 *orig = im.in + (size_t)y * W;
    uchar4 *dst = im.out + (size_t)y * W;
    // gray modes: G over R and B (reference src/pngloss_image.c:130-139); opaque modes: alpha 255 (:134,:144)
    const unsigned sel = notgray ? 0x3210u : 0x3111u, amask = notopaque ? 0u : 0xff000000u;
    if ((W & 3) == 0) {   // rows are 16-byte aligned then (256-byte aligned buffers)
        const uint4 *src4 = (const uint4 *)src, *orig4 = (const uint4 *)orig;
        uint4 *dst4 = (uint4 *)dst, *oprev4 = (uint4 *)im.oprev;
        for (int x = tid; x < W / 4; x += PL_K2_THREADS) {
            uint4 p = src4[x];
            const uint4 o = orig4[x];
            p.x = __byte_perm(p.x, 0u, sel) | amask;
            p.y = __byte_perm(p.y, 0u, sel) | amask;
            p.z = __byte_perm(p.z, 0u, sel) | amask;
            p.w = __byte_perm(p.w, 0u, sel) | amask;
            oprev4[x] = o;
            dst4[x] = p;
        }
    } else {
        for (int x = tid; x < W; x += PL_K2_THREADS) {
            const unsigned p = __byte_perm(pl_u32(src[x]), 0u, sel) | amask;
            im.oprev[x] = orig[x];
            dst[x] = pl_uc4(p);
        }
    }
}

template <int LPC, bool BM>
__global__ void __launch_bounds__(PL_K2_THREADS, PL_K2_MIN_BLOCKS(LPC))
pl_k2_quantize(const PlImageDev *imgs, const int *slots, int strength, int bleed) {
    typedef PlCfg<LPC> C;
    PL_DYN_SMEM(smem_raw);
    // 2 KB alignment for the histogram tables (the launch reserves PL_K2_SMEM_ALIGN spare bytes)
    PlCtaSmem<LPC, BM> &sm = *(PlCtaSmem<LPC, BM> *)pl_align_shared(smem_raw, PL_K2_SMEM_ALIGN);
    const int tid = threadIdx.x;
    // Warps 0 and 4 of a CTA land on the same SM sub-partition (warp id % 4) and are measurably the
    // slow pair when few CTAs share an SM (profiles/r1_filter_warp_busy.txt), so they get the two
    // cheapest predictors (none, up) and Paeth gets a sub-partition of its own.
    const int lane = tid & 31, F = (0x21340 >> (4 * (tid >> 5))) & 7;   // warp 0..4 -> none, paeth, avg, sub, up
    const int ci = lane / C::GROUP, gl = lane % C::GROUP;
    const int *my_slots = slots + (size_t)blockIdx.x * C::CPW;

    // ---- set-up ----------------------------------------------------------------------------------------
    if (tid < C::CPW) {
        const int idx = my_slots[tid];
        sm.img[tid] = imgs[idx >= 0 ? idx : my_slots[0]];
        sm.win[tid] = 0;
    }
    if (tid == 0) sm.retry = 0;
    __syncthreads();
    const int W = (int)sm.img[0].width, H = (int)sm.img[0].height;
    const bool valid = my_slots[ci] >= 0;

    PlChain cn;
    {
        const PlImageDev &im = sm.img[ci];
        const int mode = pl_image_mode(im);
        cn.in = im.in;
        cn.out = im.out;
        cn.oprev = im.oprev;
        cn.err = im.err;
        cn.cand = im.cand;
        cn.gray = mode <= 2;
        cn.alpha_rule = (mode & 1) == 0;
        cn.chmask = PL_MODE_MASK(mode);
        cn.live = valid;
        cn.q = strength;
        cn.step_magic = pl_make_magic((unsigned)strength + 1u);
    }
    // original_frequency of the image's colour mode (K1 counted every RGBA channel separately) ...
    for (int k = tid; k < C::CPW * PL_FILTERS * 256; k += PL_K2_THREADS) {
        const int c2 = k / (PL_FILTERS * 256), r = k % (PL_FILTERS * 256);
        const int f = r / 256, s = r % 256;
        const PlImageDev &im = sm.img[c2];
        const int mask = PL_MODE_MASK(pl_image_mode(im));
        unsigned v = 0;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if ((mask >> c) & 1) v += im.chan_hist[(f * 4 + c) * 256 + s];
        sm.hk[c2][f * 256 + s] = v;      // staged in the low word, replaced by its rank below
    }
    __syncthreads();
    // ... enters the candidate key only as a tie-break, so an order-preserving rank is enough:
    // rank = number of symbols with a strictly smaller count (equal counts share a rank).
    unsigned my_rank[(PlCfg<LPC>::CPW * PL_FILTERS * 256 + PL_K2_THREADS - 1) / PL_K2_THREADS];
    {
        int n = 0;
        for (int k = tid; k < C::CPW * PL_FILTERS * 256; k += PL_K2_THREADS, n++) {
            const int c2 = k / (PL_FILTERS * 256), r = k % (PL_FILTERS * 256);
            const int f = r / 256, s = r % 256;
            const unsigned mine = (unsigned)sm.hk[c2][f * 256 + s];
            unsigned rank = 0;
            for (int s2 = 0; s2 < 256; s2++) rank += (unsigned)((unsigned)sm.hk[c2][f * 256 + s2] < mine);
            my_rank[n] = rank;
        }
    }
    __syncthreads();
    {
        int n = 0;
        for (int k = tid; k < C::CPW * PL_FILTERS * 256; k += PL_K2_THREADS, n++) {
            const int c2 = k / (PL_FILTERS * 256), r = k % (PL_FILTERS * 256);
            my_rank[n] <<= PL_KEY_RANK_SHIFT;
            if (r < 256) sm.base[c2][r] = 0;
        }
    }
    __syncthreads();
    {
        int n = 0;
        for (int k = tid; k < C::CPW * PL_FILTERS * 256; k += PL_K2_THREADS, n++) {
            const int c2 = k / (PL_FILTERS * 256), r = k % (PL_FILTERS * 256);
            const int f = r / 256, s = r % 256;
            sm.hk[c2][f * 256 + ((s + c2 * C::HROT) & 255)] = (unsigned long long)my_rank[n];   // count 0
        }
    }
    __syncthreads();

    const unsigned bleed_magic = pl_make_magic((unsigned)bleed);
    if (C::DLUT)
        for (int k = tid; k < 2 * PL_DLUT_HALF; k += PL_K2_THREADS)
            sm.dlut[k] = pl_pack_taps(pl_sierra_taps(k - PL_DLUT_HALF, bleed_magic));
    if (C::DL32)
        for (int k = tid; k < 2 * PL_DL32_HALF; k += PL_K2_THREADS)
            sm.dl32[k] = pl_pack_taps6(pl_sierra_taps(k - PL_DL32_HALF, bleed_magic));
    __syncthreads();
    int prev_w = 0;
#ifdef PL_K2_PROFILE
    // debug builds only (tools/sweep.py --profile): cycles each filter warp spends inside the row pass,
    // returned in place of the first words of final_hist
    unsigned long long prof_busy = 0;
    const long long prof_start = clock64();
#endif
    bool failed = false;       // this chain's image hit "no acceptable row" (reference abort())
    unsigned retries = 0;

    for (int y = 0; y < H; y++) {
        const bool adaptive = sm.img[ci].adaptive_all || y == 0;   // reference src/pngloss_image.c:210
        cn.q = strength;
        cn.step_magic = pl_make_magic((unsigned)strength + 1u);
        cn.live = valid && !failed;
        bool pending = cn.live;
        for (;;) {
#ifdef PL_K2_PROFILE
            const long long prof_t0 = clock64();
#endif
            const unsigned long long cost =
                pl_row_pass<LPC, BM>(sm, cn, F, W, y, y & 1, prev_w, adaptive, bleed_magic);
#ifdef PL_K2_PROFILE
            prof_busy += (unsigned long long)(clock64() - prof_t0);
#endif
            if (gl == 0 && cn.live) sm.cost[ci][F] = cost;
            __syncthreads();

            // ---- pick the winner of every image that ran this pass (reference :257-263) --------------
            int w = -1;
            if (cn.live) {
                unsigned long long best = ~0ull;
#pragma unroll
                for (int f = 0; f < PL_FILTERS; f++) {
                    const unsigned long long c = sm.cost[ci][f];
                    if (c < best) { best = c; w = f; }
                }
                if (F == 0 && gl == 0) {
                    sm.win[ci] = w;
                    if (w < 0 && cn.q > 0) sm.retry = 1;
                }
            } else if (F == 0 && gl == 0) {
                sm.win[ci] = -2;   // not part of this pass
            }
            __syncthreads();

            // ---- commit: all 160 threads, image by image ----------------------------------------------
            for (int c2 = 0; c2 < C::CPW; c2++) {
                const int w2 = sm.win[c2];
                if (w2 == -2) continue;
                if (w2 >= 0) {
                    const PlImageDev &im = sm.img[c2];
                    pl_commit_row(im, w2, y, W, tid);
                    for (int s = tid; s < 256; s += PL_K2_THREADS) {
                        const unsigned v = (unsigned)(sm.hk[c2][w2 * 256 + s] >> 32);
                        sm.base[c2][s] = v;
#pragma unroll
                        for (int f = 0; f < PL_FILTERS; f++) ((unsigned *)&sm.hk[c2][f * 256 + s])[1] = v;
                    }
                    if (tid == 0) im.filters[y] = (unsigned char)(0x08 << w2);
                } else {
                    // nobody acceptable: restore the histograms for the retry at lower strength
                    for (int s = tid; s < 256; s += PL_K2_THREADS) {
                        const unsigned v = sm.base[c2][s];
#pragma unroll
                        for (int f = 0; f < PL_FILTERS; f++) ((unsigned *)&sm.hk[c2][f * 256 + s])[1] = v;
                    }
                }
            }
            const bool again = sm.retry != 0;
            if (pending) {
                if (w >= 0) {
                    prev_w = w;
                    pending = false;
                } else if (cn.q == 0) {
                    failed = true;     // reference aborts here (src/pngloss_image.c:268-271)
                    pending = false;
                } else {
                    cn.q -= 1;         // try again at lower quantization strength (:273-274)
                    cn.step_magic = pl_make_magic((unsigned)cn.q + 1u);
                    retries++;
                }
            }
            cn.live = pending;
            __syncthreads();
            if (tid == 0) sm.retry = 0;   // next write happens only after the next cost barrier
            if (!again) break;
        }
    }

    // ---- results -------------------------------------------------------------------------------------------
    for (int c2 = 0; c2 < C::CPW; c2++) {
        if (my_slots[c2] < 0) continue;
        const PlImageDev &im = sm.img[c2];
        for (int s = tid; s < 256; s += PL_K2_THREADS)
            im.final_hist[s] = sm.base[c2][(s + c2 * C::HROT) & 255];
    }
#ifdef PL_K2_PROFILE
    __syncthreads();
    if (lane == 0 && my_slots[0] >= 0) {
        uint32_t *dbg = sm.img[0].final_hist;
        dbg[2 * F] = (uint32_t)(prof_busy >> 10);
        if (F == 0) dbg[10] = (uint32_t)((unsigned long long)(clock64() - prof_start) >> 10);
    }
#endif
    if (valid && F == 0 && gl == 0) {
        const PlImageDev &im = sm.img[ci];
        im.status[0] = failed ? PL_ST_NO_ROW : PL_ST_OK;
        im.status[1] = cn.gray ? (cn.alpha_rule ? 2u : 1u) : (cn.alpha_rule ? 4u : 3u);
        im.status[2] = retries;
    }
}

#include "pl_k2_lean.cuh"
#include "pl_k2_solo.cuh"

// --------------------------------------------------------------------------------------------------
// K3: batch histogram.  out[256] (u64) += sum over images of final_hist.  grid = any, 256 threads.
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pl_k3_batch_hist(const PlImageDev *imgs, int n,
                                                        unsigned long long *out) {
    unsigned long long acc = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x) acc += imgs[i].final_hist[threadIdx.x];
    if (acc) atomicAdd(&out[threadIdx.x], acc);
}

// --------------------------------------------------------------------------------------------------
// K4: filtered PNG scanlines of the quantised images, ready for deflate.
//
// What the reference leaves to libpng after the hot path (src/rwpng.c:557-613 colour type auto-detection,
// :488-495 row filters): narrow the RGBA rows to the colour type the *output* pixels allow (gray /
// gray+alpha / rgb / rgba), apply to every row the filter the search chose - row 0: libpng's
// min-sum-of-absolute-differences heuristic, as the reference's writer does - and lay the rows out as zlib
// wants them: one filter-type byte, then width * bpp filtered bytes.
//   pl_k4_scan_output: the gray / opaque scan of the output.  4 B/pixel read.
//   pl_k4_scanlines:   grid = images * slices CTAs of 256 threads; CTA b takes rows (b % slices),
//                      (b % slices) + slices, ... of image b / slices.  4 B/pixel read (the three neighbours
//                      come from L1/L2), bpp B/pixel written through a shared-memory stage as aligned words.
// --------------------------------------------------------------------------------------------------
#define PL_K4_THREADS 256

__global__ void __launch_bounds__(PL_K4_THREADS) pl_k4_scan_output(const PlScanDev *imgs, unsigned slices) {
    const PlScanDev im = imgs[blockIdx.x / slices];
    const size_t n = (size_t)im.width * im.height;
    unsigned notgray = 0, notopaque = 0;
    const size_t first = (size_t)(blockIdx.x % slices) * PL_K4_THREADS + threadIdx.x;
    const size_t step = (size_t)slices * PL_K4_THREADS;
    // four pixels per load, four loads in flight per thread (the image buffers are 256-byte aligned)
    const uint4 *p4 = (const uint4 *)im.px;
    const size_t n4 = n / 4;
    size_t i = first;
    for (; i + 3 * step < n4; i += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = p4[i + u * step];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // gray: R == G == B, i.e. the word equals itself with G copied over R and B
                notgray |= (unsigned)(__byte_perm(w[k], 0u, 0x3111u) != w[k]);
                notopaque |= (unsigned)(w[k] < 0xff000000u);
            }
        }
    }
    for (; i < n4; i += step) {
        const uint4 v = p4[i];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            notgray |= (unsigned)(__byte_perm(w[k], 0u, 0x3111u) != w[k]);
            notopaque |= (unsigned)(w[k] < 0xff000000u);
        }
    }
    for (size_t j = n4 * 4 + first; j < n; j += step) {   // the last n % 4 pixels
        const unsigned w = pl_u32(im.px[j]);
        notgray |= (unsigned)(__byte_perm(w, 0u, 0x3111u) != w);
        notopaque |= (unsigned)(w < 0xff000000u);
    }
    if (__any_sync(PL_FULL, notgray) && (threadIdx.x & 31) == 0) atomicOr(&im.oflags[0], 1u);
    if (__any_sync(PL_FULL, notopaque) && (threadIdx.x & 31) == 0) atomicOr(&im.oflags[1], 1u);
}

// bytes of a pixel in the narrowed layout: byte selector for __byte_perm (gray: G; gray+alpha: G,A; ...)
__device__ __forceinline__ unsigned pl_k4_narrow(unsigned rgba, int bpp) {
    return bpp == 1 ? __byte_perm(rgba, 0u, 0x4441u) : bpp == 2 ? __byte_perm(rgba, 0u, 0x4431u)
           : bpp == 3 ? rgba & 0x00ffffffu : rgba;   // bytes beyond bpp are zero
}
// the bpp filtered bytes of one pixel, packed into the low bytes of a word
__device__ __forceinline__ unsigned pl_k4_filter_pixel(int type, int bpp, unsigned cur, unsigned left, unsigned up,
                                                       unsigned ul) {
    unsigned r = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c < bpp) {
            const int x = pl_byte(cur, c), a = pl_byte(left, c), b = pl_byte(up, c), d = pl_byte(ul, c);
            const int pred = type == 1 ? a : type == 2 ? b : type == 3 ? (a + b) >> 1 : type == 4 ? pl_paeth(b, d, a) : 0;
            r |= ((unsigned)(x - pred) & 255u) << (8 * c);
        }
    }
    return r;
}

// The same for all four bytes of a word at once (bytes the colour type does not use are zero in every
// operand and stay zero): per-byte subtraction modulo 256 and the per-byte floor average.
__device__ __forceinline__ unsigned pl_bytes_sub(unsigned a, unsigned b) {
    // keep the borrow of each byte inside that byte
    return ((a | 0x80808080u) - (b & 0x7f7f7f7fu)) ^ ((a ^ ~b) & 0x80808080u);
}
__device__ __forceinline__ unsigned pl_bytes_avg(unsigned a, unsigned b) {
    return (a & b) + (((a ^ b) & 0xfefefefeu) >> 1);
}
__device__ __forceinline__ unsigned pl_k4_filter_word(int type, int bpp, unsigned cur, unsigned left, unsigned up,
                                                      unsigned ul) {
    if (type == 4) return pl_k4_filter_pixel(4, bpp, cur, left, up, ul);
    const unsigned pred = type == 1 ? left : type == 2 ? up : type == 3 ? pl_bytes_avg(left, up) : 0u;
    return pl_bytes_sub(cur, pred);
}

// filter type of row y: row 0 by the heuristic (row0_type), the others from K2's libpng masks
__device__ __forceinline__ int pl_k4_row_type(const PlScanDev &im, int y, int row0_type) {
    if (y == 0) return row0_type;
    const unsigned m = im.filters[y];
    return m == 0x10 ? 1 : m == 0x20 ? 2 : m == 0x40 ? 3 : m == 0x80 ? 4 : 0;
}

// Vector path of K4 (width a multiple of 4, so rows are 16-byte aligned): the CTA's rows are cut into
// segments of 256 * PL_K4_PX pixels; a thread takes PL_K4_GROUPS groups of four pixels of a segment - group k of
// thread t is group k * 256 + t of the segment, so that the loads of a warp are contiguous and its 16-byte stage
// stores hit every bank once - each 4 * bpp bytes = bpp whole words of the output stream, which are staged as words;
// the copy-out realigns the stream to the destination with funnel
// shifts and writes 16-byte vectors.  The loads of segment s + 1 are issued before segment s is copied out and
// the stage is double buffered, so there is one barrier per segment and loads are always in flight.
#ifndef PL_K4_GROUPS
#define PL_K4_GROUPS 4   /* 4-pixel groups per thread and segment (measured: profiles/r2_k4.txt) */
#endif
#define PL_K4_PX (4 * PL_K4_GROUPS)
#define PL_K4_STAGE_WORDS (PL_K4_THREADS * PL_K4_PX + 8)
struct PlK4Seg {
    uint4 c4[PL_K4_GROUPS], u4[PL_K4_GROUPS];   // the thread's groups of four pixels of the row and of the row above
    unsigned left[PL_K4_GROUPS], ul[PL_K4_GROUPS];   // the pixel before each group, narrowed
};
// four pixels' worth of filtered bytes: the filter type is the same for the whole CTA, so the switch is uniform
template <int BPP, int TYPE>
__device__ __forceinline__ void pl_k4_filter4(const uint4 &c4, const uint4 &u4, unsigned &left, unsigned &ul,
                                              unsigned (&r)[4]) {
    const unsigned cw[4] = {c4.x, c4.y, c4.z, c4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned cur = pl_k4_narrow(cw[k], BPP), up = pl_k4_narrow(uw[k], BPP);
        r[k] = pl_k4_filter_word(TYPE, BPP, cur, left, up, ul);
        left = cur;
        ul = up;
    }
}
template <int BPP>
__device__ __forceinline__ void pl_k4_vector_rows(const PlScanDev &im, int first_row, int slices, int row0_type,
                                                  unsigned (*stage)[PL_K4_STAGE_WORDS]) {
    const int W = (int)im.width, H = (int)im.height, tid = threadIdx.x;
    const size_t stride = 1 + (size_t)W * BPP;
    const int seg_px = PL_K4_THREADS * PL_K4_PX;
    auto load = [&](int y, int x0, PlK4Seg &g) {
        const uchar4 *row = im.px + (size_t)y * W;
#pragma unroll
        for (int k = 0; k < PL_K4_GROUPS; k++) {
            const int x = x0 + (k * PL_K4_THREADS + tid) * 4;
            // (the width is a multiple of 4, so a group is either whole or beyond the row)
            g.c4[k] = g.u4[k] = make_uint4(0u, 0u, 0u, 0u);
            g.left[k] = g.ul[k] = 0u;
            if (y < H && x < W) {
                g.c4[k] = *(const uint4 *)(row + x);
                if (x) g.left[k] = pl_k4_narrow(pl_u32(row[x - 1]), BPP);
                if (y) {
                    g.u4[k] = *(const uint4 *)(row - W + x);
                    if (x) g.ul[k] = pl_k4_narrow(pl_u32(row[x - 1 - W]), BPP);
                }
            }
        }
    };
    PlK4Seg seg;
    load(first_row, 0, seg);
    int buf = 0;
    for (int y = first_row, x0 = 0; y < H; buf ^= 1) {
        const int type = pl_k4_row_type(im, y, row0_type);
        const int npx = min(seg_px, W - x0), len = npx * BPP;       // pixels and bytes of this segment
        unsigned *stage_w = stage[buf];
        unsigned char *dst_row = im.scan + (size_t)y * stride;
        if (x0 == 0 && tid == 0) dst_row[0] = (unsigned char)type;
#pragma unroll
        for (int k = 0; k < PL_K4_GROUPS; k++) {
            const int grp = k * PL_K4_THREADS + tid;   // group of four pixels within the segment
            if (grp * 4 < npx) {
                unsigned left = seg.left[k], ul = seg.ul[k];
                unsigned r[4];
                switch (type) {
                case 0: pl_k4_filter4<BPP, 0>(seg.c4[k], seg.u4[k], left, ul, r); break;
                case 1: pl_k4_filter4<BPP, 1>(seg.c4[k], seg.u4[k], left, ul, r); break;
                case 2: pl_k4_filter4<BPP, 2>(seg.c4[k], seg.u4[k], left, ul, r); break;
                case 3: pl_k4_filter4<BPP, 3>(seg.c4[k], seg.u4[k], left, ul, r); break;
                default: pl_k4_filter4<BPP, 4>(seg.c4[k], seg.u4[k], left, ul, r); break;
                }
                unsigned *sp = stage_w + grp * BPP;
                if (BPP == 4) {
                    *(uint4 *)sp = make_uint4(r[0], r[1], r[2], r[3]);
                } else if (BPP == 3) {
                    sp[0] = r[0] | (r[1] << 24);
                    sp[1] = (r[1] >> 8) | (r[2] << 16);
                    sp[2] = (r[2] >> 16) | (r[3] << 8);
                } else if (BPP == 2) {
                    sp[0] = r[0] | (r[1] << 16);
                    sp[1] = r[2] | (r[3] << 16);
                } else {
                    sp[0] = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
                }
            }
        }
        if (tid == 0) stage_w[len >> 2] = 0u;   // the word after the stream (read by the last shift)
        // next segment: its loads are in flight during the barrier and the copy-out
        unsigned char *g0 = dst_row + 1 + (size_t)x0 * BPP;
        x0 += seg_px;
        if (x0 >= W) {
            x0 = 0;
            y += slices;
        }
        load(y, x0, seg);
        __syncthreads();
        // 16-byte vectors of the destination [g0 - mis, g0 + len): vector v holds stream bytes
        // 16 v - mis .. 16 v - mis + 15
        const int mis = (int)((size_t)g0 & 15u);
        const int nvec = (mis + len + 15) >> 4;
        const int sh = ((16 - mis) & 3) * 8;                        // stream word -> destination word shift
        const unsigned char *sb = (const unsigned char *)stage_w;
        // The five stream words a vector needs, stage_w[4 vi - c .. 4 vi - c + 4] with c = ceil(mis / 4), lie in the
        // aligned quads vi - 1 and vi: two conflict-free 16-byte reads (five 4-byte reads at a four-word stride between
        // threads cost four bank-conflict passes each; profiles/r2_k4.txt).
        const int c = (mis + 3) >> 2;
        const uint4 *sq = (const uint4 *)stage_w;
        for (int vi = tid; vi < nvec; vi += PL_K4_THREADS) {
            const int o = vi * 16 - mis;                            // stream offset of the vector's first byte
            if (o >= 0 && o + 16 <= len) {
                const uint4 pq = vi ? sq[vi - 1] : make_uint4(0u, 0u, 0u, 0u), cq = sq[vi];
                unsigned w0, w1, w2, w3, w4;
                switch (c) {   // CTA-uniform
                case 0: w0 = cq.x; w1 = cq.y; w2 = cq.z; w3 = cq.w; w4 = 0u; break;   // (aligned: no shift, w4 unused)
                case 1: w0 = pq.w; w1 = cq.x; w2 = cq.y; w3 = cq.z; w4 = cq.w; break;
                case 2: w0 = pq.z; w1 = pq.w; w2 = cq.x; w3 = cq.y; w4 = cq.z; break;
                case 3: w0 = pq.y; w1 = pq.z; w2 = pq.w; w3 = cq.x; w4 = cq.y; break;
                default: w0 = pq.x; w1 = pq.y; w2 = pq.z; w3 = pq.w; w4 = cq.x; break;
                }
                *(uint4 *)(g0 + o) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                                __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
            }
        }
        // the ragged first and last vector, one byte per thread
        if (tid < 32) {
            const int o = (tid < 16 ? 0 : (nvec - 1) * 16) - mis + (tid & 15);
            const bool ragged = tid < 16 ? mis != 0 : (nvec > 1 && ((mis + len) & 15) != 0);
            if (ragged && o >= 0 && o < len) g0[o] = sb[o];
        }
        // no second barrier: the next segment is staged in the other buffer, and the one after that only after
        // the next barrier, which every thread reaches after this copy-out
    }
}

__global__ void __launch_bounds__(PL_K4_THREADS) pl_k4_scanlines(const PlScanDev *imgs, unsigned slices) {
    // staging of a segment of the output stream, double buffered (vector path) / byte stage (scalar path)
    __shared__ __align__(16) unsigned stage2[2][PL_K4_STAGE_WORDS];
    __shared__ unsigned red[5][PL_K4_THREADS / 32];
    __shared__ int row0_type_s;
    unsigned char *stage = (unsigned char *)stage2[0];
    const PlScanDev im = imgs[blockIdx.x / slices];
    const int W = (int)im.width, H = (int)im.height;
    const bool gray = im.oflags[0] == 0, opaque = im.oflags[1] == 0;
    const int bpp = gray ? (opaque ? 1 : 2) : (opaque ? 3 : 4);
    const size_t stride = 1 + (size_t)W * bpp;
    const int tid = threadIdx.x;
    const int first_row = (int)(blockIdx.x % slices);

    int row0_type = 0;
    if (first_row == 0) {
        // libpng's heuristic on row 0 (reference src/rwpng.c:488-495 leaves it to libpng): the filter
        // with the smallest sum of |signed residual|, first minimum in the order none .. paeth
        const uchar4 *row = im.px;
        unsigned sum[5] = {0, 0, 0, 0, 0};
        for (int x = tid; x < W; x += PL_K4_THREADS) {
            const unsigned cur = pl_k4_narrow(pl_u32(row[x]), bpp);
            const unsigned left = x ? pl_k4_narrow(pl_u32(row[x - 1]), bpp) : 0u;
#pragma unroll
            for (int f = 0; f < 5; f++) {
                const unsigned r = pl_k4_filter_pixel(f, bpp, cur, left, 0u, 0u);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < bpp) {
                        const unsigned v = (r >> (8 * c)) & 255u;
                        sum[f] += v < 128u ? v : 256u - v;
                    }
            }
        }
#pragma unroll
        for (int f = 0; f < 5; f++) {
#pragma unroll
            for (int mk = 16; mk >= 1; mk >>= 1) sum[f] += __shfl_xor_sync(PL_FULL, sum[f], mk);
            if ((tid & 31) == 0) red[f][tid >> 5] = sum[f];
        }
        __syncthreads();
        if (tid == 0) {
            unsigned best = ~0u;
            int pick = 0;
            for (int f = 0; f < 5; f++) {
                unsigned t = 0;
                for (int k = 0; k < PL_K4_THREADS / 32; k++) t += red[f][k];
                if (t < best) { best = t; pick = f; }
            }
            row0_type_s = pick;
            im.oflags[2] = (unsigned)pick;
        }
        __syncthreads();
        row0_type = row0_type_s;
    }

    if ((W & 3) == 0) {
        switch (bpp) {
        case 1: pl_k4_vector_rows<1>(im, first_row, (int)slices, row0_type, stage2); break;
        case 2: pl_k4_vector_rows<2>(im, first_row, (int)slices, row0_type, stage2); break;
        case 3: pl_k4_vector_rows<3>(im, first_row, (int)slices, row0_type, stage2); break;
        default: pl_k4_vector_rows<4>(im, first_row, (int)slices, row0_type, stage2); break;
        }
        return;
    }
    // scalar path: one pixel per thread, bytes staged at the destination's misalignment, 4-byte words out
    for (int y = first_row; y < H; y += (int)slices) {
        const uchar4 *row = im.px + (size_t)y * W;
        const uchar4 *above = row - W;   // only read when y > 0
        const int type = pl_k4_row_type(im, y, row0_type);
        unsigned char *dst_row = im.scan + (size_t)y * stride;
        if (tid == 0) dst_row[0] = (unsigned char)type;
        for (int x0 = 0; x0 < W; x0 += PL_K4_THREADS) {
            const int npx = min(PL_K4_THREADS, W - x0);
            unsigned char *g0 = dst_row + 1 + (size_t)x0 * bpp;          // destination of this segment
            const unsigned mis = (unsigned)((size_t)g0 & 3u);            // stage[mis + i] <-> g0[i]
            const int len = npx * bpp;
            if (tid < npx) {
                const int x = x0 + tid;
                const unsigned cur = pl_k4_narrow(pl_u32(row[x]), bpp);
                const unsigned left = x ? pl_k4_narrow(pl_u32(row[x - 1]), bpp) : 0u;
                unsigned up = 0, ul = 0;
                if (y) {
                    up = pl_k4_narrow(pl_u32(above[x]), bpp);
                    ul = x ? pl_k4_narrow(pl_u32(above[x - 1]), bpp) : 0u;
                }
                const unsigned r = pl_k4_filter_pixel(type, bpp, cur, left, up, ul);
                unsigned char *s = stage + mis + tid * bpp;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < bpp) s[c] = (unsigned char)(r >> (8 * c));
            }
            __syncthreads();
            // aligned words of [g0 - mis, g0 + len): whole words as 32-bit stores, the ragged ends bytewise
            const int nwords = ((int)mis + len + 3) >> 2;
            for (int wi = tid; wi < nwords; wi += PL_K4_THREADS) {
                const int b0 = wi * 4;                                   // offset in stage
                if (b0 >= (int)mis && b0 + 4 <= (int)mis + len) {
                    *(unsigned *)(g0 - mis + b0) = *(const unsigned *)(stage + b0);
                } else {
                    for (int k = 0; k < 4; k++) {
                        const int o = b0 + k;
                        if (o >= (int)mis && o < (int)mis + len) g0[o - (int)mis] = stage[o];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// --------------------------------------------------------------------------------------------------
// Synthetic gradient + noise RGBA image, identical to oracle_synth_rgba.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pl_splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) pl_k_synth(uchar4 *dst, uint32_t w, uint32_t h,
                                                  unsigned long long seed) {
    const unsigned long long n = (unsigned long long)w * h;
    const unsigned dx = w > 1 ? w - 1 : 1, dy = h > 1 ? h - 1 : 1;
    const unsigned dxy = (w + h > 2) ? w + h - 2 : 1;
    // images are at most 2^29 pixels (batch_create), so x, y, x + y times 255 fit 32 bits
    for (unsigned long long p = (unsigned long long)blockIdx.x * 256 + threadIdx.x; p < n;
         p += (unsigned long long)gridDim.x * 256) {
        const unsigned p32 = (unsigned)p, y = p32 / w, x = p32 - y * w;
        int base[4];
        base[0] = (int)(x * 255u / dx);
        base[1] = (int)(y * 255u / dy);
        base[2] = (int)((x + y) * 255u / dxy);
        base[3] = 255 - base[0] / 2;
        unsigned char v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // z mod 17 for a 64-bit z: 2^32 = 1 (mod 17), so it is (high word + low word) mod 17
            const unsigned long long z = pl_splitmix64((seed << 40) + p * 4 + c);
            const unsigned fold = (unsigned)(z >> 32) % 17u + (unsigned)z % 17u;
            const int nz = (int)(fold % 17u) - 8;
            int t = base[c] + nz;
            t = t < 0 ? 0 : t > 255 ? 255 : t;
            if (c == 3 && (y / 64) % 4 == 0 && x < w / 16) t = 0;
            v[c] = (unsigned char)t;
        }
        dst[p] = make_uchar4(v[0], v[1], v[2], v[3]);
    }
}
