// K2, lean variant ("K2L"): the quantise + filter-search kernel for large batches.
//
// Same algorithm and the same bit-exact results as pl_k2_quantize<1, true> (pl_kernels.cuh: one lane per
// colour channel, 8 images per CTA, one warp per filter candidate, bucket-maxima look-up), restructured
// around what the round-1 profiles showed (profiles/r1_k2_lanes1_bm_final.txt):
//   * shared memory per CTA 112 KB -> 72 KB, so that three CTAs (24 images, 15 warps) share an SM instead of
//     two: the running symbol counts are kept as 16-bit per-row increments on top of one 32-bit table per
//     image (all five candidates of a row start from the same counts), ranks are bytes, and the row-start
//     copy is that base table itself;
//   * the five input streams of a row (original row y, original and quantised row y-1, the previous winner's
//     two error rows) are fetched ONCE per CTA by the TMA unit - 1-D bulk copies (cp.async.bulk, SASS UBLKCP)
//     into a ring of tiles guarded by mbarriers - instead of five times (once per filter warp) with per-lane
//     4-byte cp.async; the warp that releases a tile last refills it, so nobody ever waits for a free slot;
//   * tiles are 16 pixels instead of 4: a quarter of the tile overhead;
//   * the commit of a byte updates the band's own bucket directly (the chosen symbol lies in the band that
//     was looked up) instead of locating the bucket by division; symbol 0 - the extra candidate of the band
//     [-q, 0] - has a table entry of its own ("bucket Z") instead of three table look-ups;
//   * tables are padded instead of rotated (no index arithmetic against bank conflicts).
// Requirements (checked by the launcher, pl_api.cu): width % 4 == 0 (16-byte aligned rows for the bulk
// copies), width < PL_BM_MAX_WIDTH.  Any strength works (outside PL_BM_MIN_STEP..PL_BM_MAX_STEP, and while
// a row is retried at a lower strength, every byte takes the scan path), but the launcher only picks this
// kernel where the table exists.
//
// Replaces the same reference code as K2: src/pngloss_image.c:159-309, src/optimize_state.c:114-361,390-562.
#pragma once
#include <stddef.h>

#ifndef PL_L_T
#define PL_L_T 16        // pixels per tile (a multiple of 4)
#endif
#ifndef PL_L_STAGES
#define PL_L_STAGES 2    // tiles in the input ring
#endif
#define PL_L_CPW 8       // images per CTA (chains per warp)
#define PL_L_HALO 4      // pixels of left context in front of a tile (one is used; 16-byte granularity)

// One tile of the five input streams, for the 8 images of the CTA.  Row strides are = 4 (mod 32) words so
// that the 8 chains of a warp hit different banks.
struct PlLeanStage {
    uint32_t orig[PL_L_CPW][PL_L_T + PL_L_HALO];   // original row y, pixels x0-4 .. x0+15
    uint32_t oa[PL_L_CPW][PL_L_T + PL_L_HALO];     // original row y-1
    uint32_t na[PL_L_CPW][PL_L_T + PL_L_HALO];     // quantised row y-1
    short4 e0[PL_L_CPW][PL_L_T + 2];               // incoming error row 0, cells x0+4 .. x0+19
    short4 e1[PL_L_CPW][PL_L_T + 2];               // incoming error row 1, cells x0+4 .. x0+19
};
// output staging of one warp (= one filter candidate of the 8 images)
struct PlLeanOut {
    uint32_t back[PL_L_CPW][PL_L_T + PL_L_HALO];   // candidate pixels; [3] = last pixel of the previous tile
    short4 n0[PL_L_CPW][PL_L_T + 2];               // finished cells of the next error row 0
    short4 n1[PL_L_CPW][PL_L_T + 2];               // finished cells of the next error row 1
};
#define PL_L_DELTA_WORDS (128 + 4)
#define PL_L_BASE_WORDS (256 + 4)
#define PL_L_RANK_BYTES (256 + 16)
struct PlLeanSmem {
    PlLeanStage stage[PL_L_STAGES];                     // first member: bulk-copy destinations, 16-byte aligned
    unsigned long long full[PL_L_STAGES];               // mbarrier: the tile's bytes have landed
    unsigned long long empty[PL_L_STAGES];              // mbarrier: all five warps are done with the tile
    unsigned released[PL_L_STAGES];                     // release counter: elects the warp that refills a tile
    int prevw[PL_L_CPW];                                // winner of the previous row (whose error rows to read)
    int win[PL_L_CPW];
    unsigned livemask;                                  // images that take part in the current pass
    int retry;
    // symbol_frequency of chain (image, filter) = base[image][s] + 16-bit delta[image][filter][s]
    uint32_t delta[PL_L_CPW][PL_FILTERS][PL_L_DELTA_WORDS];
    uint32_t base[PL_L_CPW][PL_L_BASE_WORDS];
    unsigned char rank[PL_L_CPW][PL_FILTERS][PL_L_RANK_BYTES];   // rank of original_frequency[filter][s]
    uint2 bmk[PL_L_CPW][PL_FILTERS][PL_BM_MAX + 1];     // bucket winners: .x = relative key, .y = base count;
                                                        // entry P1 + N1 ("Z") is symbol 0 alone, at position q
    unsigned dl32[2 * PL_DL32_HALF];                    // Sierra taps by error value (pl_pack_taps6)
    PlLeanOut out[PL_K2_WARPS];
    unsigned long long cost[PL_L_CPW][PL_FILTERS];
    PlImageDev img[PL_L_CPW];
};
#define PL_L_SMEM_ALIGN 128

// Issues the bulk copies of tile t of row y into ring stage s and arms its mbarrier.  Warp-collective (all 32
// lanes converged): lane c2 < 8 issues the copies of image c2, lane 0 announces the byte count.
__device__ __forceinline__ void pl_lean_fill(PlLeanSmem &sm, int s, int t, int W, int y, int parity) {
    const int lane = threadIdx.x & 31;
    const unsigned live = sm.livemask;
    const int x0 = t * PL_L_T;
    const int npx = min(PL_L_T, W - x0);
    const int hal = t ? PL_L_HALO : 0;
    const bool first = (y == 0);
    const unsigned px_bytes = (unsigned)(npx + hal) * 4u, er_bytes = (unsigned)npx * 8u;
    PlLeanStage &st = sm.stage[s];
    if (lane < PL_L_CPW && ((live >> lane) & 1u)) {
        const PlImageDev &im = sm.img[lane];
        pl_bulk_g2s(&st.orig[lane][PL_L_HALO - hal], im.in + (size_t)y * W + x0 - hal, px_bytes, &sm.full[s]);
        if (!first) {
            const int EW = W + PL_ERR_PAD;
            const short4 *E0 = im.err + ((size_t)(parity * PL_FILTERS + sm.prevw[lane]) * 2 + 0) * EW;
            pl_bulk_g2s(&st.oa[lane][PL_L_HALO - hal], im.oprev + x0 - hal, px_bytes, &sm.full[s]);
            pl_bulk_g2s(&st.na[lane][PL_L_HALO - hal], im.out + (size_t)(y - 1) * W + x0 - hal, px_bytes,
                        &sm.full[s]);
            pl_bulk_g2s(&st.e0[lane][0], E0 + x0 + 4, er_bytes, &sm.full[s]);
            pl_bulk_g2s(&st.e1[lane][0], E0 + EW + x0 + 4, er_bytes, &sm.full[s]);
        }
    }
    if (lane == 0)
        pl_mbar_arrive_expect_tx(&sm.full[s],
                                 (unsigned)__popc(live) * (first ? px_bytes : 3u * px_bytes + 2u * er_bytes));
}

// One candidate row of the 8 images of the CTA: warp = filter F, lane = (image ci, channel ch).
// `use` = number of ring tiles the CTA has consumed before this pass (CTA-uniform).
// ALLACT: every lane of the warp is an active channel of a live chain (eight live RGBA images - the large-batch
// case); the per-lane activity tests then fold away.
template <bool ALLACT>
__device__ __forceinline__ unsigned long long pl_lean_row_pass(PlLeanSmem &sm, const PlChain &cn, int F, int W,
                                                               int y, int parity, bool adaptive,
                                                               unsigned bleed_magic, unsigned use) {
    const int lane = threadIdx.x & 31;
    const int ci = lane >> 2, ch = lane & 3;
    PlLeanOut &wo = sm.out[F];
    uint32_t *dl = sm.delta[ci][F];
    const volatile unsigned short *d16 = (const volatile unsigned short *)dl;   // other lanes increment it
    const uint32_t *bs = sm.base[ci];
    const unsigned char *rk = sm.rank[ci][F];
    uint2 *bmrow = sm.bmk[ci][F];
    const int EW = W + PL_ERR_PAD;
    const bool live = ALLACT || cn.live;
    const bool first = (y == 0);
    const int chmask = ALLACT ? 0xF : cn.chmask;
    const bool alpha_rule = ALLACT || cn.alpha_rule;
    const unsigned step_magic = cn.step_magic;
    const bool act = ALLACT || (live && ((chmask >> ch) & 1));
    const int q = cn.q, step = cn.q + 1;
    const PlPredictor predictor = pl_make_predictor(F);
    const unsigned canon = chmask == 0xF ? 0x3210u : chmask == 0x7 ? 0x4210u : chmask == 0xA ? 0x3111u : 0x4111u;
    const int prev_w = sm.prevw[ci];

    const short4 *Ecur0 = cn.err + ((size_t)(parity * PL_FILTERS + prev_w) * 2 + 0) * EW;
    const short4 *Ecur1 = Ecur0 + EW;
    short4 *En0 = cn.err + ((size_t)((parity ^ 1) * PL_FILTERS + F) * 2 + 0) * EW;
    short4 *En1 = En0 + EW;
    uchar4 *rcand = cn.cand + (size_t)F * W;

    // Sierra window of this lane's channel (see pl_row_pass)
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

    // ---- bucket winners of this candidate at the start of the row (all increments are zero: counts = base) --
    const PlBm bmc = pl_bm_counts(step, W);
    const int tz = bmc.P1 + bmc.N1;      // bucket Z
    if (live) {
        for (int t = ch; t < bmc.P1 + bmc.N1 + (bmc.P1 > 0 ? 1 : 0); t += 4) {
            // negative bucket 0 leaves symbol 0 to bucket 0 and to Z (see pl_bm_counts); Z is symbol 0 alone
            const int lo_t = t == tz ? -q : pl_bm_low(bmc, t, step);
            const int p0 = t == tz ? q : 0, p1 = t == bmc.P1 ? q - 1 : q;
            unsigned long long best = 0;
            for (int p = p0; p <= p1; p++) {
                const unsigned s = (unsigned)(lo_t + p) & 255u;
                const unsigned long long key =
                    ((unsigned long long)bs[s] << 32) | ((unsigned)rk[s] << PL_KEY_RANK_SHIFT) | (unsigned)(511 - p);
                best = key > best ? key : best;
            }
            const unsigned mcount = (unsigned)(best >> 32);
            const unsigned base = mcount > 4u * (unsigned)W ? mcount - 4u * (unsigned)W : 0u;
            bmrow[t] = make_uint2(pl_bm_key(mcount - base, ((unsigned)best >> PL_KEY_RANK_SHIFT) & 255u,
                                            511 - (int)((unsigned)best & 511u)),
                                  base);
        }
    }
    __syncwarp();

    const int ntiles = (W + PL_L_T - 1) / PL_L_T;
    for (int t = 0; t < ntiles; t++) {
        const unsigned u = use + (unsigned)t;
        const int s = (int)(u % PL_L_STAGES);
        const unsigned ph = (u / PL_L_STAGES) & 1u;
        const PlLeanStage &st = sm.stage[s];
        const int x0 = t * PL_L_T;
        const int npx = min(PL_L_T, W - x0);
        pl_mbar_wait(&sm.full[s], ph);

        // this lane's channel of the tile's pixels / error cells, as flat pointers: the look-ahead below reads one
        // element past the tile (into the next row of the staging arrays, never used)
        const unsigned char *po = (const unsigned char *)&st + offsetof(PlLeanStage, orig) +
                                  (ci * (PL_L_T + PL_L_HALO) + PL_L_HALO) * 4 + ch;
        const unsigned char *pa = (const unsigned char *)&st + offsetof(PlLeanStage, na) +
                                  (ci * (PL_L_T + PL_L_HALO) + PL_L_HALO) * 4 + ch;
        const short *pe0 = (const short *)((const unsigned char *)&st + offsetof(PlLeanStage, e0)) +
                           ci * (PL_L_T + 2) * 4 + ch;
        const short *pe1 = (const short *)((const unsigned char *)&st + offsetof(PlLeanStage, e1)) +
                           ci * (PL_L_T + 2) * 4 + ch;
        int o_n = po[0];
        int a_n = pa[0];
        int e0_n = pe0[0];
        int e1_n = pe1[0];
        for (int i = 0; i < npx; i++) {
            const int o = o_n, a = a_n;
            a2 = e0_n;
            b4 = e1_n;
            c3 = 0;
            // prefetch pixel i+1, off the dependency chain (one slot beyond the tile is readable padding)
            o_n = po[(i + 1) * 4];
            a_n = pa[(i + 1) * 4];
            e0_n = pe0[(i + 1) * 4];
            e1_n = pe1[(i + 1) * 4];

            // ---- band of admissible symbols (reference src/optimize_state.c:158-210) ---------------
            // wrap (:175-182): the exact symbol orig - predicted, brought into [-128, 127], is the signed
            // low byte of the difference; `pred` is the predictor shifted by the same multiple of 256
            const int pred0 = pl_predict_rt(predictor, a, aprev, left);
            int ex = pl_sext8(o - pred0);
            const int pred = o - ex;
            const int err_in = pl_sext16(a0);
            int here = o + err_in;
            const int want = ex + err_in;                       // here - pred
            const unsigned m = (unsigned)(want < 0 ? -want : want);
            const unsigned kq = pl_udiv_magic(m, step_magic);   // which band, counted from zero
            const bool neg = want < 0;
            const int ks = (int)kq * step;
            const int lo_u = neg ? -(ks + q) : ks;              // first symbol of the band before clamping
            // clamp (:195-210): symbol + predicted must be a byte (see pl_row_pass)
            const int smin = -pred, smax = 255 - pred;
            int lo = min(max(lo_u, smin), smax);
            int hi = min(max(lo_u + q, smin), smax);
            const bool transp = alpha_rule && ch == 3 && o == 0;
            if (transp) {   // fully transparent stays transparent (:158-164)
                here = 0;
                lo = hi = ex = smin;
            }
            const int span = act ? hi - lo : -1;

            // ---- the band's winner from the bucket table -------------------------------------------
            const bool tvalid = !transp && kq < (unsigned)(neg ? bmc.N1 : bmc.P1);
            const int tl = (int)kq + (neg ? bmc.P1 : 0);
            const unsigned long long be64 = *(volatile unsigned long long *)&bmrow[tvalid ? tl : 0];   // one LDS.64
            const unsigned be_x = (unsigned)be64, base_l = (unsigned)(be64 >> 32);
            const int wsym = lo_u + 127 - (int)(be_x & 127u);
            const bool inr = act && tvalid && wsym >= lo && wsym <= hi;
            // best so far as (count, low word) - the two halves of pl_row_pass's 64-bit candidate key
            unsigned bc = inr ? base_l + (be_x >> PL_BM_COUNT_SHIFT) : 0u;
            unsigned bl = inr ? ((((be_x >> 7) & 255u) << PL_KEY_RANK_SHIFT) | (unsigned)(511 - (wsym - lo))) : 0u;
            // band [-q, 0]: symbol 0 is not a member of negative bucket 0; its own entry joins in (unless the
            // clamp cut it off; without a bucket winner inside the band the scan below covers it anyway)
            if (inr && neg && kq == 0 && hi == 0) {
                const unsigned long long z64 = *(volatile unsigned long long *)&bmrow[tz];
                const unsigned zc = (unsigned)(z64 >> 32) + ((unsigned)z64 >> PL_BM_COUNT_SHIFT);
                const unsigned zl = ((((unsigned)z64 >> 7) & 255u) << PL_KEY_RANK_SHIFT) | (unsigned)(511 + lo);
                if (zc > bc || (zc == bc && zl > bl)) {
                    bc = zc;
                    bl = zl;
                }
            }
            const bool need_scan = act && !inr && !(span == 0 && ex == lo);
#ifdef PL_SIMT_EMU
            if (__any_sync(PL_FULL, need_scan)) PL_EMU_COUNT(PL_CNT_BM_SCAN);
            else PL_EMU_COUNT(PL_CNT_BM_LOOKUP);
#endif
            if (need_scan) {   // rare (clamped bands, retries below the table's strength range); diverged
                for (int p = 0; p <= span; p++) {
                    const unsigned s2 = (unsigned)(lo + p) & 255u;
                    const unsigned c = bs[s2] + d16[s2];
                    const unsigned l = ((unsigned)rk[s2] << PL_KEY_RANK_SHIFT) | (unsigned)(511 - p);
                    if (c > bc || (c == bc && l > bl)) {
                        bc = c;
                        bl = l;
                    }
                }
            }
            // the exact symbol, with its bonus bit (:228-244)
            {
                const unsigned s2 = (unsigned)ex & 255u;
                const unsigned c = bs[s2] + d16[s2];
                const int pos = ex - lo;
                const unsigned l = ((unsigned)rk[s2] << PL_KEY_RANK_SHIFT) | pl_key_low(true, pos & 255);
                if (pos >= 0 && pos <= span && (c > bc || (c == bc && l > bl))) {
                    bc = c;
                    bl = l;
                }
            }
            int bpos = 511 - (int)(bl & 511u);

            // ---- fix-up: replay the channel order (see pl_row_pass) ---------------------------------
            bool conflict = false;
            unsigned dup = 0;
            const unsigned bf = min(bc, 0xfffff0u);
            {
                const unsigned mine = (bf << 8) | ((unsigned)(lo + bpos) & 255u);
#pragma unroll
                for (int t2 = 0; t2 < 3; t2++) {
                    const unsigned theirs = __shfl_sync(PL_FULL, mine, (lane & ~3) + t2);
                    const int pos = ((int)(theirs & 255u) - lo) & 255;
                    const bool earlier = (ch > t2) & act & (bool)((chmask >> t2) & 1);
                    conflict |= earlier & (pos <= span) & (pos != bpos) & ((theirs >> 8) + 3u >= bf);
                    dup += (unsigned)(earlier & (pos == bpos));
                }
            }
            if (!__any_sync(PL_FULL, conflict)) {
                PL_EMU_COUNT(PL_CNT_FIXUP_SKIPPED);
                bc += dup;   // this channel's symbol has been counted dup times since the look-up
            } else {
                PL_EMU_COUNT(PL_CNT_FIXUP_REPLAY);
#pragma unroll 1
                for (int t2 = 0; t2 < 3; t2++) {
                    const int src = (lane & ~3) + t2;
                    const int vsym = __shfl_sync(PL_FULL, lo + bpos, src);
                    const unsigned vc = __shfl_sync(PL_FULL, bc, src);
                    const unsigned vl = __shfl_sync(PL_FULL, bl, src);
                    if (ch > t2 && act && ((chmask >> t2) & 1)) {
                        const int pos = (vsym - lo) & 255;
                        if (pos <= span) {
                            if (pos == bpos) {
                                bc += 1u;            // my own winner was chosen again
                            } else {
                                const unsigned c = vc + 1u;
                                const unsigned l = (vl & ~1023u) | pl_key_low(lo + pos == ex, pos);
                                if (c > bc || (c == bc && l > bl)) {
                                    bc = c;
                                    bl = l;
                                    bpos = pos;
                                }
                            }
                        }
                    }
                }
            }

            // ---- commit the byte ----------------------------------------------------------------------
            const int sym = lo + bpos;
            const int back = act ? sym + pred : 0;
            if (act) {
                const unsigned s2 = (unsigned)sym & 255u;
                atomicAdd(&dl[s2 >> 1], 1u << ((s2 & 1u) * 16u));
                if (bmc.P1 > 0) {
                    // the symbol's new key enters every bucket that holds its bin; bc = its count before this
                    // channel's increment
                    const unsigned now = bc + 1u;
                    const unsigned rank7 = (bl >> (PL_KEY_RANK_SHIFT - 7)) & (255u << 7);
                    const int s8 = pl_sext8(sym);
                    if (tvalid && (unsigned)(sym - lo_u) <= (unsigned)q && s8 > bmc.seam_n && s8 < bmc.seam_p) {
                        // the usual case: the symbol lies in the band that was looked up, i.e. in bucket tl (a band
                        // that the clamp emptied collapses onto a value outside of it), and not in the seam, so
                        // that sym == s8 and tl is its only bucket - except symbol 0, which lives in bucket 0 and Z
                        PL_EMU_COUNT(PL_CNT_BM_FASTUPD);
                        const int tu = sym == 0 ? 0 : tl;
                        const unsigned base_u = tu == tl ? base_l : *(volatile unsigned *)&bmrow[0].y;
                        if (now >= base_u)
                            atomicMax(&bmrow[tu].x, ((now - base_u) << PL_BM_COUNT_SHIFT) | rank7 |
                                                        (unsigned)(127 - (sym == 0 ? 0 : sym - lo_u)));
                        if (sym == 0) {
                            const unsigned base2 = *(volatile unsigned *)&bmrow[tz].y;
                            if (now >= base2)
                                atomicMax(&bmrow[tz].x,
                                          ((now - base2) << PL_BM_COUNT_SHIFT) | rank7 | (unsigned)(127 - q));
                        }
                    } else {
                        PL_EMU_COUNT(PL_CNT_BM_GENERAL);
                        const unsigned as8 = (unsigned)(s8 < 0 ? -s8 : s8);
                        const unsigned k8 = pl_udiv_magic(as8, step_magic);
                        const int rs = (int)(as8 - k8 * (unsigned)step);
                        {
                            const int t1 = (int)k8 + (s8 < 0 ? bmc.P1 : 0);
                            const unsigned base1 = *(volatile unsigned *)&bmrow[t1].y;
                            if (now >= base1)
                                atomicMax(&bmrow[t1].x, ((now - base1) << PL_BM_COUNT_SHIFT) | rank7 |
                                                            (unsigned)(127 - (s8 < 0 ? q - rs : rs)));
                        }
                        if (s8 == 0 || s8 >= bmc.seam_p || s8 <= bmc.seam_n) {
                            // symbol 0: its own entry Z (position q); bins >= seam_p: symbol s8 - 256 of the last
                            // negative bucket; bins <= seam_n: symbol s8 + 256 of the last non-negative one
                            const bool up = s8 <= bmc.seam_n;
                            const int t2 = s8 == 0 ? tz : up ? bmc.P1 - 1 : bmc.P1 + bmc.N1 - 1;
                            const int pos2 = s8 == 0 ? q : s8 + (up ? 256 : -256) - pl_bm_low(bmc, t2, step);
                            const unsigned base2 = *(volatile unsigned *)&bmrow[t2].y;
                            if (now >= base2)
                                atomicMax(&bmrow[t2].x,
                                          ((now - base2) << PL_BM_COUNT_SHIFT) | rank7 | (unsigned)(127 - pos2));
                        }
                    }
                }
                ((unsigned char *)&wo.back[ci][PL_L_HALO + i])[ch] = (unsigned char)back;
            }
            left = back;
            aprev = a;

            // ---- Sierra diffusion of (here - back) / bleed (reference :390-467) ----------------------
            const int diff = act ? pl_sext16(here - back) : 0;
            PlTaps tp;
            if ((unsigned)(diff + PL_DL32_HALF) < 2u * PL_DL32_HALF) {   // per lane
                PL_EMU_COUNT(PL_CNT_TAPS_TABLE);
                tp = pl_unpack_taps6(sm.dl32[diff + PL_DL32_HALF]);
            } else {
                PL_EMU_COUNT(PL_CNT_TAPS_COMPUTED);
                tp = pl_sierra_taps(diff, bleed_magic);
            }
            a1 += tp.rem;
            a2 += tp.threes;
            b0 += tp.twos;
            b1 += tp.fours;
            b2 += tp.five;
            b3 += tp.fours;
            b4 += tp.twos;
            c1 += tp.twos;
            c2 += tp.threes;
            c3 += tp.twos;
            ((short *)&wo.n0[ci][i])[ch] = (short)b0;
            ((short *)&wo.n1[ci][i])[ch] = (short)c0;
            a0 = a1; a1 = a2;
            b0 = b1; b1 = b2; b2 = b3; b3 = b4;
            c0 = c1; c1 = c2; c2 = c3;
            __syncwarp();
        }

        // ---- tile epilogue: lane (ci, ch) takes pixels ch, ch + 4, ch + 8, ch + 12 of its chain's tile ----
        if (live) {
#pragma unroll 1
            for (int p = ch; p < npx; p += 4) {
                const int x = x0 + p;
                const unsigned o4 = st.orig[ci][PL_L_HALO + p], q4 = wo.back[ci][PL_L_HALO + p];
                const unsigned oa4 = st.oa[ci][PL_L_HALO + p], na4 = st.na[ci][PL_L_HALO + p];
                unsigned ol4 = 0, ql4 = 0, oad4 = 0, nad4 = 0;
                if (x > 0) {
                    ol4 = st.orig[ci][PL_L_HALO + p - 1];
                    ql4 = wo.back[ci][PL_L_HALO + p - 1];
                    oad4 = st.oa[ci][PL_L_HALO + p - 1];
                    nad4 = st.na[ci][PL_L_HALO + p - 1];
                }
                En0[x] = wo.n0[ci][p];
                En1[x] = wo.n1[ci][p];
                rcand[x] = pl_uc4(q4);
                // derivative error of the three neighbours (reference :265-287), see pl_row_pass
                {
                    const unsigned o = __byte_perm(o4, 0u, canon), qq = __byte_perm(q4, 0u, canon);
                    const unsigned n1o = __byte_perm(oa4, 0u, canon), n1n = __byte_perm(na4, 0u, canon);
                    const unsigned n2o = __byte_perm(oad4, 0u, canon), n2n = __byte_perm(nad4, 0u, canon);
                    const unsigned n3o = __byte_perm(ol4, 0u, canon), n3n = __byte_perm(ql4, 0u, canon);
                    unsigned xs = __dp4a(o, o, __dp4a(qq, qq, 0u)) * 3u;
                    xs = __dp4a(n1o, n1o, __dp4a(n1n, n1n, xs));
                    xs = __dp4a(n2o, n2o, __dp4a(n2n, n2n, xs));
                    xs = __dp4a(n3o, n3o, __dp4a(n3n, n3n, xs));
                    unsigned ys = __dp4a(o, qq, 0u) * 3u;
                    ys = __dp4a(n1o, n1n, __dp4a(n1o, o, __dp4a(n1n, qq, ys)));
                    ys = __dp4a(n2o, n2n, __dp4a(n2o, o, __dp4a(n2n, qq, ys)));
                    ys = __dp4a(n3o, n3n, __dp4a(n3o, o, __dp4a(n3n, qq, ys)));
                    unsigned zs = __dp4a(n1o, qq, __dp4a(n1n, o, 0u));
                    zs = __dp4a(n2o, qq, __dp4a(n2n, o, zs));
                    zs = __dp4a(n3o, qq, __dp4a(n3n, o, zs));
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
        }
        __syncwarp();
        if (ch == 0 && live) wo.back[ci][PL_L_HALO - 1] = wo.back[ci][PL_L_HALO + npx - 1];

        // ---- release the tile.  The warp that does so FIRST (the fastest of the five) waits for the other four and
        // refills the stage: the slowest warp - the one everybody waits for at the end of the row - never issues
        // copies and never waits for a tile (the refill it needs next was issued a whole tile ago).
        unsigned order = 1;
        if (lane == 0) {
            order = atomicAdd(&sm.released[s], 1u) % PL_K2_WARPS;   // counts up forever: 5 releases per use of a stage
            pl_mbar_arrive(&sm.empty[s]);
        }
        order = __shfl_sync(PL_FULL, order, 0);
        if (order == 0 && t + PL_L_STAGES < ntiles) {
            pl_mbar_wait(&sm.empty[s], ph);     // all five warps are done with the tile
            pl_lean_fill(sm, s, t + PL_L_STAGES, W, y, parity);
        }
        __syncwarp();
    }

    // ---- row tail: cells W .. W+3 of the two outgoing error rows -----------------------------------------
    if (act) {
        ((short *)&En0[W + 0])[ch] = (short)b0;
        ((short *)&En0[W + 1])[ch] = (short)b1;
        ((short *)&En0[W + 2])[ch] = (short)b2;
        ((short *)&En0[W + 3])[ch] = (short)b3;
        ((short *)&En1[W + 0])[ch] = (short)c0;
        ((short *)&En1[W + 1])[ch] = (short)c1;
        ((short *)&En1[W + 2])[ch] = (short)c2;
        ((short *)&En1[W + 3])[ch] = 0;
    }

    // ---- row cost (reference src/optimize_state.c:314-360), see pl_row_pass ------------------------------
    unsigned bits = 0;
    for (int s2 = ch; s2 < 256; s2 += 4) {
        const unsigned dv = d16[s2];
        bits += dv * (33u + (unsigned)__clz((int)(bs[s2] + dv)));
    }
#pragma unroll
    for (int mk = 1; mk < 4; mk <<= 1) {
        derr += __shfl_xor_sync(PL_FULL, derr, mk);
        bits += __shfl_xor_sync(PL_FULL, bits, mk);
    }
    unsigned long long cost = derr / 128ull + bits;
    if (__any_sync(PL_FULL, adaptive)) {
#pragma unroll
        for (int mk = 1; mk < 4; mk <<= 1) {
            as0 += __shfl_xor_sync(PL_FULL, as0, mk);
            as1 += __shfl_xor_sync(PL_FULL, as1, mk);
            as2 += __shfl_xor_sync(PL_FULL, as2, mk);
            as3 += __shfl_xor_sync(PL_FULL, as3, mk);
            as4 += __shfl_xor_sync(PL_FULL, as4, mk);
        }
        unsigned lowest = min(min(min(as0, as1), min(as2, as3)), as4);
        int pick = lowest >= as0 ? 0 : lowest >= as1 ? 1 : lowest >= as2 ? 2 : lowest >= as3 ? 3 : 4;
        if (adaptive && pick != F) cost = ~0ull;
    }
    return cost;
}

#ifndef PL_K2L_MIN_BLOCKS
#define PL_K2L_MIN_BLOCKS 3
#endif

__global__ void __launch_bounds__(PL_K2_THREADS, PL_K2L_MIN_BLOCKS)
pl_k2_lean(const PlImageDev *imgs, const int *slots, int strength, int bleed) {
    PL_DYN_SMEM(smem_raw);
    PlLeanSmem &sm = *(PlLeanSmem *)pl_align_shared(smem_raw, PL_L_SMEM_ALIGN);
    const int tid = threadIdx.x;
    const int lane = tid & 31, F = (0x21340 >> (4 * (tid >> 5))) & 7;   // warp 0..4 -> none, paeth, avg, sub, up
    const int ci = lane >> 2, ch = lane & 3;
    const int *my_slots = slots + (size_t)blockIdx.x * PL_L_CPW;

    // ---- set-up ----------------------------------------------------------------------------------------
    if (tid < PL_L_CPW) {
        const int idx = my_slots[tid];
        sm.img[tid] = imgs[idx >= 0 ? idx : my_slots[0]];
        sm.win[tid] = 0;
        sm.prevw[tid] = 0;
    }
    if (tid == 0) {
        sm.retry = 0;
        for (int s = 0; s < PL_L_STAGES; s++) {
            pl_mbar_init(&sm.full[s], 1);
            pl_mbar_init(&sm.empty[s], PL_K2_WARPS);
            sm.released[s] = 0;
        }
        pl_fence_mbar_init();
    }
    // the inputs of row 0 that do not exist (row above, incoming errors) read as zero: the bulk copies of
    // row 0 never touch these parts of the ring
    for (int k = tid; k < (int)(sizeof(sm.stage) / 4); k += PL_K2_THREADS) ((uint32_t *)sm.stage)[k] = 0;
    for (int k = tid; k < (int)(sizeof(sm.out) / 4); k += PL_K2_THREADS) ((uint32_t *)sm.out)[k] = 0;
    for (int k = tid; k < (int)(sizeof(sm.delta) / 4); k += PL_K2_THREADS) ((uint32_t *)sm.delta)[k] = 0;
    for (int k = tid; k < (int)(sizeof(sm.base) / 4); k += PL_K2_THREADS) ((uint32_t *)sm.base)[k] = 0;
    pl_fence_proxy_async();
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
    // ---- ranks of original_frequency (K1 counted every RGBA channel separately), image by image; the counts
    // are staged in the (still unused) output staging area ------------------------------------------------------
    {
        uint32_t *stg = (uint32_t *)sm.out;   // PL_FILTERS * 256 words
        for (int c2 = 0; c2 < PL_L_CPW; c2++) {
            const PlImageDev &im = sm.img[c2];
            const int mask = PL_MODE_MASK(pl_image_mode(im));
            for (int k = tid; k < PL_FILTERS * 256; k += PL_K2_THREADS) {
                const int f = k >> 8, s = k & 255;
                unsigned v = 0;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if ((mask >> c) & 1) v += im.chan_hist[(f * 4 + c) * 256 + s];
                stg[k] = v;
            }
            __syncthreads();
            for (int k = tid; k < PL_FILTERS * 256; k += PL_K2_THREADS) {
                const int f = k >> 8, s = k & 255;
                const unsigned mine = stg[k];
                unsigned rank = 0;
                for (int s2 = 0; s2 < 256; s2++) rank += (unsigned)(stg[f * 256 + s2] < mine);
                sm.rank[c2][f][s] = (unsigned char)rank;
            }
            __syncthreads();
        }
        for (int k = tid; k < (int)(sizeof(sm.out) / 4); k += PL_K2_THREADS) ((uint32_t *)sm.out)[k] = 0;
    }
    const unsigned bleed_magic = pl_make_magic((unsigned)bleed);
    for (int k = tid; k < 2 * PL_DL32_HALF; k += PL_K2_THREADS)
        sm.dl32[k] = pl_pack_taps6(pl_sierra_taps(k - PL_DL32_HALF, bleed_magic));
    __syncthreads();

    bool failed = false;
    unsigned retries = 0;
    unsigned use = 0;   // ring tiles consumed so far (CTA-uniform)
    const int ntiles = (W + PL_L_T - 1) / PL_L_T;

    for (int y = 0; y < H; y++) {
        const bool adaptive = sm.img[ci].adaptive_all || y == 0;   // reference src/pngloss_image.c:210
        cn.q = strength;
        cn.step_magic = pl_make_magic((unsigned)strength + 1u);
        cn.live = valid && !failed;
        bool pending = cn.live;
        for (;;) {
            // ---- who runs this pass; prime the ring --------------------------------------------------------
            {
                const unsigned lm = __ballot_sync(PL_FULL, cn.live && ch == 0);   // bit 4 ci per live image
                if (tid < 32) {
                    unsigned m8 = 0;
#pragma unroll
                    for (int c2 = 0; c2 < PL_L_CPW; c2++) m8 |= ((lm >> (4 * c2)) & 1u) << c2;
                    if (lane == 0) sm.livemask = m8;
                    __syncwarp();
                    for (int k = 0; k < PL_L_STAGES && k < ntiles; k++)
                        pl_lean_fill(sm, (int)((use + (unsigned)k) % PL_L_STAGES), k, W, y, y & 1);
                }
            }
            __syncthreads();
            const bool allact = __all_sync(PL_FULL, cn.live && cn.chmask == 0xF);
            const unsigned long long cost =
                allact ? pl_lean_row_pass<true>(sm, cn, F, W, y, y & 1, adaptive, bleed_magic, use)
                       : pl_lean_row_pass<false>(sm, cn, F, W, y, y & 1, adaptive, bleed_magic, use);
            use += (unsigned)ntiles;
            if (ch == 0 && cn.live) sm.cost[ci][F] = cost;
            __threadfence();          // this thread's candidate / error rows reach L2 (the TMA unit reads there) ...
            pl_fence_proxy_async();   // ... and are ordered before the next pass's bulk reads
            __syncthreads();

            // ---- pick the winner of every image that ran this pass (reference :257-263) ------------------
            int w = -1;
            if (cn.live) {
                unsigned long long best = ~0ull;
#pragma unroll
                for (int f = 0; f < PL_FILTERS; f++) {
                    const unsigned long long c = sm.cost[ci][f];
                    if (c < best) { best = c; w = f; }
                }
                if (F == 0 && ch == 0) {
                    sm.win[ci] = w;
                    if (w >= 0) sm.prevw[ci] = w;
                    if (w < 0 && cn.q > 0) sm.retry = 1;
                }
            } else if (F == 0 && ch == 0) {
                sm.win[ci] = -2;   // not part of this pass
            }
            __syncthreads();

            // ---- commit: all 160 threads, image by image ------------------------------------------------------
            for (int c2 = 0; c2 < PL_L_CPW; c2++) {
                const int w2 = sm.win[c2];
                if (w2 == -2) continue;
                if (w2 >= 0) {
                    const PlImageDev &im = sm.img[c2];
                    pl_commit_row(im, w2, y, W, tid);
                    if (tid == 0) im.filters[y] = (unsigned char)(0x08 << w2);
                }
                // the winner's increments join the base counts; every candidate starts the next pass from zero
                // (no winner: the row is retried from the unchanged base)
                for (int k = tid; k < 128; k += PL_K2_THREADS) {
                    if (w2 >= 0) {
                        const unsigned d = sm.delta[c2][w2][k];
                        sm.base[c2][2 * k] += d & 0xffffu;
                        sm.base[c2][2 * k + 1] += d >> 16;
                    }
#pragma unroll
                    for (int f = 0; f < PL_FILTERS; f++) sm.delta[c2][f][k] = 0;
                }
            }
            const bool again = sm.retry != 0;
            if (pending) {
                if (w >= 0) {
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
            __threadfence();
            pl_fence_proxy_async();   // the committed row, before the next row's bulk reads
            __syncthreads();
            if (tid == 0) sm.retry = 0;
            if (!again) break;
        }
    }

    // ---- results -------------------------------------------------------------------------------------------
    for (int c2 = 0; c2 < PL_L_CPW; c2++) {
        if (my_slots[c2] < 0) continue;
        const PlImageDev &im = sm.img[c2];
        for (int s = tid; s < 256; s += PL_K2_THREADS) im.final_hist[s] = sm.base[c2][s];
    }
    if (valid && F == 0 && ch == 0) {
        const PlImageDev &im = sm.img[ci];
        im.status[0] = failed ? PL_ST_NO_ROW : PL_ST_OK;
        im.status[1] = cn.gray ? (cn.alpha_rule ? 2u : 1u) : (cn.alpha_rule ? 4u : 3u);
        im.status[2] = retries;
    }
}
