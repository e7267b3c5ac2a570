// K2, latency variant ("K2S"): the quantise + filter-search kernel for batches that leave every image a CTA of
// its own (up to four per SM) - the reference's actual use (one image per call, src/pngloss.c:173-205,266), the
// few-large-image configurations, command lines over a few hundred files.
//
// Same algorithm and bit-identical results as pl_k2_quantize / pl_k2_lean.  Those kernels let one warp do
// everything a candidate row needs.  A lone warp issues one instruction about every four cycles (fixed-latency
// dependencies, in-order issue: every exposed shared-memory or branch latency is paid in full), so a pixel step
// costs (instructions on that warp) x 3-4 cycles ~ 1000 cycles whatever the lane mapping.  Here the roles are split
// across the warps of one CTA (one image per CTA) and the one warp that carries the serial chain executes as few
// instructions as the algorithm allows:
//   * chain warp(s): a lane is (filter, channel); the bucket-maxima look-up (pl_kernels.cuh) needs no candidate
//     scan, so four lanes serve a candidate.  FPW = 5: one warp carries all five candidates (20 lanes, the warp
//     has a scheduler to itself: it is the highest warp id of its sub-partition and the other warp there idles at
//     the CTA barrier); FPW = 1: one warp per candidate.  Per pixel the chain reads ONE pre-digested word
//     (original byte, quantised byte above, incoming error) and writes ONE word (quantised byte, diffused
//     difference); the band of a byte, the table entries of a histogram bin and the error taps come from
//     per-pass look-up tables in shared memory; every table is addressed by 32-bit shared address.
//   * a producer warp, tiles ahead: packs original row y, quantised row y-1 and the previous winner's error row 0
//     into those words (ring of PL_S_STAGES tiles, mbarrier hand-off).
//   * post warps (five, one per candidate; two in the four-warp layout), tiles behind, one lane per pixel: everything
//     that is separable - the
//     two outgoing error rows are a 5-tap / 3-tap stencil over the differences (src/optimize_state.c:445-467),
//     the derivative error (:265-287), libpng's heuristic sums (:492-562), the candidate row.
// Strengths 0 .. 126 (the bucket tables have room for the one-symbol buckets of strength 0).
// The row end (cost, winner, commit, histogram clone, table rebuild) is done by all warps between CTA barriers.
//
// Replaces the same reference code as K2: src/pngloss_image.c:159-309, src/optimize_state.c:114-361,390-562.
#pragma once

#ifndef PL_S_T
#define PL_S_T 32          // pixels per tile = lanes of a producer / post warp
#endif
#ifndef PL_S_STAGES
#define PL_S_STAGES 4      // tiles in the rings (post warps release a tile one tile late: >= 3)
#endif
#ifndef PL_S_ONEVOTE
#define PL_S_ONEVOTE 0     // measured (profiles/r2_solo_variants.txt)
#endif
// Who arrives on the ring's mbarriers: lane 0 on behalf of its warp, after a __syncwarp (the product), or every lane
// itself (-DPL_S_ALL_ARRIVE=1: the sanitizer build - racecheck follows a thread's own arrive / wait only, and reports
// the one-lane idiom as a race between the other lanes' accesses and the waiters'; profiles/r2_compute_sanitizer.txt).
#ifndef PL_S_ALL_ARRIVE
#define PL_S_ALL_ARRIVE 0
#endif
#define PL_S_ARRIVERS (PL_S_ALL_ARRIVE ? 32 : 1)
#define PL_S_ARRIVES(lane) (PL_S_ALL_ARRIVE || (lane) == 0)
#define PL_S_TAPC_HALF 512 // the chain's tap table covers differences -512 .. 511
#define PL_S_BOFF 1024     // the band table covers here - predicted = -1024 .. 1023
#define PL_S_HPAD 4        // entries between the candidates' histograms (bank stagger)
#define PL_S_BM_ROW 260     // entries of a candidate's bucket table: up to 129 + 130 buckets (strength 0) + Z+; like the
                            // histograms' row, 8 banks (mod 32) per candidate: the candidates of a half-warp never collide

// Warp roles.  The hardware arbiter prefers the higher warp id of a sub-partition (warp id % 4), so a chain warp is the
// higher warp of its sub-partition.
//   FPW = 5 (8 warps, up to two CTAs per SM): the chain warp is warp 4 + r and warp r idles at the CTA barriers, so that
//     the chain has the scheduler of sub-partition r to itself; r rotates with the CTA index (wave number), so that the
//     chain warps of the CTAs that share an SM sit on different schedulers.  The other six warps are, in ascending
//     order, the producer and the five post warps.
//   FPW = 5, compact (4 warps, up to four CTAs per SM - batches of 297 .. 592 images): warp r is the chain warp, the
//     other three are the producer and two post warps (candidates 0 2 4 and 1 3).
//   FPW = 1 (12 warps): warps 0 producer, 1 .. 5 post, 6 idle, 7 .. 11 chain.
template <int FPW, bool COMPACT = false>
struct PlSoloCfg {
    static const int NCHAIN = PL_FILTERS / FPW;             // chain warps
    static const int NWARPS = FPW == 5 ? (COMPACT ? 4 : 8) : 12;
    static const int THREADS = 32 * NWARPS;
    static const int NPOSTW = COMPACT ? 2 : PL_FILTERS;     // post warps
    static const int NF = COMPACT ? 3 : 1;                  // candidates per post warp (at most)
    static const int MIN_BLOCKS = FPW == 5 ? (COMPACT ? 4 : 2) : 1;
    // role of warp w: -3 idle, -2 chain, -1 producer, 0 .. post warp index
    __device__ static int role(int w, int rot) {
        if (FPW == 5 && COMPACT) {
            if (w == rot) return -2;
            return w - (w > rot ? 1 : 0) - 1;               // the other three, ascending: -1, 0, 1
        }
        if (FPW == 5) {
            if (w == 4 + rot) return -2;
            if (w == rot) return -3;
            const int k = w - (w > rot ? 1 : 0) - (w > 4 + rot ? 1 : 0);   // index among the other six
            return k - 1;
        }
        return w == 0 ? -1 : w <= 5 ? w - 1 : w == 6 ? -3 : -2;
    }
    __device__ static int chain_index(int w) { return FPW == 5 ? 0 : w - 7; }
};

// Buckets of this kernel: pl_bm_counts without its lower strength limit (that one comes from the 18-entry tables of
// pl_k2_quantize / pl_k2_lean; here a candidate's table has room for the 259 one-symbol buckets of strength 0).
__device__ __forceinline__ PlBm pl_solo_bm_counts(int step, int width) {
    PlBm b;
    const bool on = step >= 1 && step <= PL_BM_MAX_STEP && width < PL_BM_MAX_WIDTH;
    const int KP = 128 / step, KN = 129 / step;
    b.P1 = on ? KP + 1 : 0;
    b.N1 = on ? KN + 1 : 0;
    b.seam_p = on ? 256 - KN * step - (step - 1) : 999;
    b.seam_n = on ? KP * step + (step - 1) - 256 : -999;
    return b;
}

struct PlSoloSmem {
    // per candidate and symbol: high word = running symbol_frequency, low word = rank of
    // original_frequency[filter][symbol] << PL_KEY_RANK_SHIFT (the two halves of K2's candidate key)
    unsigned long long hk[PL_FILTERS][256 + PL_S_HPAD];
    uint2 bmk[PL_FILTERS][PL_S_BM_ROW];     // bucket winners: .x = relative key, .y = base count; [P1 + N1] = band [-q, 0]
    uint32_t base[256];                     // symbol_frequency at the start of the row
    uint4 bins[256];                        // the table entries of a histogram bin: {entry 0, its base, entry 1, its base}
    uint2 bins3[256];                       // ... and a third one (large strengths only)
    uint32_t band[2 * PL_S_BOFF];           // here - predicted -> band start << 16 | zero-band flag | entry (pl_solo_band_entry)
    uint32_t tapc[2 * PL_S_TAPC_HALF];      // chain: difference -> the two taps that stay in the row (rem | threes << 16)
    uint32_t dl32[2 * PL_DL32_HALF];        // post warps: all five taps (pl_pack_taps6)
    uint4 pre[PL_S_STAGES][PL_S_T + 1];     // per pixel, word ch: orig | above << 8 | incoming error << 16
    uint4 outw[PL_S_STAGES][PL_FILTERS][PL_S_T + 1];   // per candidate and pixel, word ch: back | difference << 16
                                                       // (+ 1: the candidates' rows are 4 banks apart)
    unsigned long long pre_full[PL_S_STAGES], pre_empty[PL_S_STAGES];
    unsigned long long out_full[PL_S_STAGES], out_empty[PL_S_STAGES];
    unsigned long long derr[PL_FILTERS];
    unsigned asum[PL_FILTERS][5];
    unsigned bits[PL_FILTERS];
    uint4 trash[PL_S_T];                    // where the idle lanes of a chain warp store
    uint32_t sink[32];                      // ... and where a lane without an active channel sends its atomics (a word each)
    PlImageDev img;
};

// The entries of the bucket table that hold histogram bin b at the strength of the pass (the same for all
// candidates): the bin's own bucket, in the seam the last bucket of the other sign (pl_bm_counts), and for the bins
// of [-q, 0] the zero band "Z+" (pl_solo_chain).  An entry is (byte offset of its key in a candidate's row) << 16 |
// (127 - position of the bin in the bucket), and comes with the bucket's base count (0xffffffff: no such entry).
struct PlBinEntries { unsigned e[3], base[3]; };
__device__ __forceinline__ PlBinEntries pl_solo_bin_entries(const PlSoloSmem &sm, const PlBm &bmc, int b, int q, int step,
                                                            unsigned step_magic) {
    PlBinEntries r;
#pragma unroll
    for (int k = 0; k < 3; k++) { r.e[k] = 0; r.base[k] = 0xffffffffu; }
    if (bmc.P1 == 0) return r;
    const int s8 = pl_sext8(b);
    const unsigned as8 = (unsigned)(s8 < 0 ? -s8 : s8);
    const unsigned k8 = pl_udiv_magic(as8, step_magic);
    const int rs = (int)(as8 - k8 * (unsigned)step);
    int n = 0;
    {
        const int t = (int)k8 + (s8 < 0 ? bmc.P1 : 0), pos = s8 < 0 ? q - rs : rs;
        r.e[n] = ((unsigned)t * 8u) << 16 | (unsigned)(127 - pos);
        r.base[n++] = sm.bmk[0][t].y;
    }
    if (s8 >= bmc.seam_p || s8 <= bmc.seam_n) {
        const bool up = s8 <= bmc.seam_n;
        const int t = up ? bmc.P1 - 1 : bmc.P1 + bmc.N1 - 1;
        const int pos = s8 + (up ? 256 : -256) - pl_bm_low(bmc, t, step);
        r.e[n] = ((unsigned)t * 8u) << 16 | (unsigned)(127 - pos);
        r.base[n++] = sm.bmk[0][t].y;
    }
    if ((unsigned)(s8 + q) <= (unsigned)q) {
        const int t = bmc.P1 + bmc.N1;
        r.e[n] = ((unsigned)t * 8u) << 16 | (unsigned)(127 - (s8 + q));
        r.base[n++] = sm.bmk[0][t].y;
    }
    return r;
}

// The band of admissible symbols of a byte whose here - predicted is `want` (reference src/optimize_state.c:186-193)
// and the table entry that holds its winner: band start << 16 | 0x8000 if the band is [-q, 0] (Z+ instead of negative
// bucket 0 when symbol 0 is admissible) | entry (0x7ff: none).
__device__ __forceinline__ unsigned pl_solo_band_entry(const PlBm &bmc, int want, int q, int step, unsigned step_magic) {
    const bool neg = want < 0;
    const unsigned m = (unsigned)(neg ? -want : want);
    const unsigned kq = pl_udiv_magic(m, step_magic);
    const int ks = (int)kq * step;
    const int lo_u = neg ? -(ks + q) : ks;
    const bool tvalid = kq < (unsigned)(neg ? bmc.N1 : bmc.P1);
    const unsigned t = tvalid ? kq + (neg ? (unsigned)bmc.P1 : 0u) : 0x7ffu;
    const unsigned z = (tvalid && neg && kq == 0u) ? 0x8000u : 0u;
    return ((unsigned)lo_u << 16) | z | t;
}

// Commit of a byte: count the symbol, and let its new key (count `now`, rank) enter every table entry that holds its
// bin.  No predicates (they would become branches): a key that must not enter is 0.
// `three`: some bin of this strength has a third entry (warp-uniform).  The loads are a call of their own: the fast
// path issues them for its provisional symbol before the channel vote, so that their latency is not exposed.
struct PlBinLoaded { uint4 bi; unsigned long long bj; };
__device__ __forceinline__ PlBinLoaded pl_solo_bin_load(PlSh bins_sh, PlSh bins3_sh, int sym, bool three) {
    const unsigned bin = (unsigned)sym & 255u;
    PlBinLoaded r;
    r.bi = pl_lds128(bins_sh + bin * 16u);
    r.bj = three ? pl_lds64(bins3_sh + bin * 8u) : 0ull;
    return r;
}
__device__ __forceinline__ void pl_solo_commit(PlSh hk_sh, PlSh bm_sh, const PlBinLoaded &ld, int sym, unsigned now,
                                               unsigned rank7, unsigned actm, PlSh sink_sh, bool three) {
    // A lane without an active channel (actm = 0) sends its atomics to a word of its own: on the real tables they would
    // be no-ops, but atomics on one address are served one after the other.
    const unsigned bin = (unsigned)sym & 255u;
    const uint4 bi = ld.bi;
    pl_atoms_add32(pl_sh_select(hk_sh + bin * 8u + 4u, sink_sh, actm), 1u);
    // (an entry that does not exist - base 0xffffffff - or that the key must not enter goes to the sink too: all such
    // lanes would otherwise meet on entry 0 of their candidate's row)
    const unsigned k0 = ((now - bi.y) << PL_BM_COUNT_SHIFT) | rank7 | (bi.x & 127u);
    const unsigned k1 = ((now - bi.w) << PL_BM_COUNT_SHIFT) | rank7 | (bi.z & 127u);
    pl_atoms_max32(pl_sh_select(bm_sh + (bi.x >> 16), sink_sh, now >= bi.y ? actm : 0u), k0);
    pl_atoms_max32(pl_sh_select(bm_sh + (bi.z >> 16), sink_sh, now >= bi.w ? actm : 0u), k1);
    if (three) {
        const unsigned e2 = (unsigned)ld.bj, b2 = (unsigned)(ld.bj >> 32);
        const unsigned k2 = ((now - b2) << PL_BM_COUNT_SHIFT) | rank7 | (e2 & 127u);
        pl_atoms_max32(pl_sh_select(bm_sh + (e2 >> 16), sink_sh, now >= b2 ? actm : 0u), k2);
    }
}

// ---- chain warp: the dependent chain of one candidate row (FPW = 1) or of all five (FPW = 5) ------------------
//
// Bucket table of a candidate (see pl_kernels.cuh "bucket maxima"), as this kernel lays it out: non-negative
// buckets at 0 .. P1-1, negative ones at P1 .. P1+N1-1 (negative bucket 0 is [-q, -1]), and at P1 + N1 the winner of
// the whole zero band [-q, 0] ("Z+"): a band [-q, 0] whose symbol 0 is admissible (0 <= predicted <= 255) looks Z+
// up instead of negative bucket 0, so no second entry has to be merged on the chain.
//
// Every pixel takes the FAST path first when the strength has a table: straight-line code for the bytes whose answer
// is the looked-up winner, the exact symbol, or forced.  With b = band start + predicted (the band in byte values,
// [b, b + q], clamped to [lob, hib] inside 0 .. 255): the bucket's winner at position p is admissible iff
// 0 <= b + p <= 255, the exact symbol iff b <= orig <= b + q, and the chosen byte is b + p or orig; a band of one value
// (cut down to it, collapsed onto 0 or 255, or a fully transparent pixel) needs no table at all.  One vote asks
// whether every lane got its answer that way, a second one whether an earlier channel of the pixel disturbs a later
// one (if so the channel order is replayed exactly, still on this path).  If a lane has no answer - a band that the
// byte range cut short and that lost its bucket's winner that way (saturated regions), a band beyond the table - the
// warp redoes the pixel on the GENERAL path (the code of pl_lean_row_pass: scan, general bands), which also serves
// the strengths without a table.  (Letting such a lane scan its band inside the fast path was measured: slower on
// every image, profiles/r2_solo_variants.txt.)  Both paths give the reference's answer; the fast one only commits
// where it is provably the same.  Returns the number of pixels that took the general path (the caller stops
// attempting the fast path on images where it mostly fails).
template <int FPW>
__device__ __forceinline__ unsigned pl_solo_chain(PlSoloSmem &sm, int f, int ch, bool lane_act, int chmask,
                                                  bool alpha_rule, int q, unsigned step_magic, int W,
                                                  unsigned bleed_magic, unsigned use, bool try_fast) {
    const int lane = threadIdx.x & 31;
    const bool act = lane_act && ((chmask >> ch) & 1);
    const int step = q + 1;
    const int ff = lane_act ? f : 0;   // idle lanes (FPW = 5: lanes 20 .. 31) shadow candidate 0 and write nothing
    // (opaque: ptxas would otherwise re-derive every shared address from SR_CgaCtaId inside the loop)
    const PlSh hk_sh = pl_sh_opaque(pl_sh(sm.hk[ff])), bm_sh = pl_sh_opaque(pl_sh(sm.bmk[ff]));
    const PlSh bins_sh = pl_sh_opaque(pl_sh(sm.bins)), bins3_sh = pl_sh_opaque(pl_sh(sm.bins3));
    const PlSh band_sh = pl_sh_opaque(pl_sh(sm.band));
    const PlSh tapc_sh = pl_sh_opaque(pl_sh(sm.tapc + PL_S_TAPC_HALF));
    const unsigned actm = act ? ~0u : 0u;   // as a mask: predicates are scarce and get recomputed
    const PlSh sink_sh = pl_sh_opaque(pl_sh(&sm.sink[lane]));
    // predictor of this lane's candidate, branch-free (FPW = 5: the lanes of a warp differ)
    const int ma = (ff == 2 || ff == 3) ? 255 : 0, ml = (ff == 1 || ff == 3) ? 255 : 0, sh = ff == 3 ? 1 : 0;
    const unsigned pm = ff == 4 ? ~0u : 0u;
    const PlBm bmc = pl_solo_bm_counts(step, W);
    const int P1 = bmc.P1, N1 = bmc.N1, tz = bmc.P1 + bmc.N1;
    const bool table = try_fast && P1 > 0;
    const bool al = alpha_rule && ch == 3;
    // a bin has a third table entry where the seam reaches the zero band: last non-negative bucket's wrap >= -q
    // (measured, profiles/r2_solo_variants.txt: with one chain warp the third, usually idle, atomic is cheaper than
    // the branch around it; with five chain warps the atomics are what the warps queue for)
    const bool three = FPW == 5 || (P1 > 0 && bmc.seam_n >= -q);
    // earlier channels of the pixel that are active (the channel order of the fix-up)
    unsigned emask = 0;
#pragma unroll
    for (int t2 = 0; t2 < 3; t2++)
        if (ch > t2 && act && ((chmask >> t2) & 1)) emask |= 1u << t2;
    const int src0 = lane & ~3;

    int left = 0, aprev = 0;
    int carry_a = 0, carry_b = 0;   // error this row's pixels x-1, x-2 send to pixel x / pixel x-1 sends to x+1
    unsigned general_px = 0;
    const int ntiles = (W + PL_S_T - 1) / PL_S_T;
    for (int t = 0; t < ntiles; t++) {
        const unsigned u = use + (unsigned)t;
        const int s = (int)(u % PL_S_STAGES);
        const unsigned ph = (u / PL_S_STAGES) & 1u;
        const int npx = min(PL_S_T, W - t * PL_S_T);
        pl_mbar_wait(&sm.pre_full[s], ph);
        pl_mbar_wait(&sm.out_empty[s], ph ^ 1u);
        const PlSh pp = pl_sh_opaque(pl_sh((uint32_t *)&sm.pre[s][0] + ch));
        // idle lanes store into a scratch tile
        const PlSh po = pl_sh_opaque(pl_sh(lane_act ? (uint32_t *)&sm.outw[s][ff][0] + ch : (uint32_t *)&sm.trash[0] + ch));
        uint32_t w_n = pl_lds32(pp);
#pragma unroll 1
        for (int i = 0; i < npx; i++) {
            const uint32_t w = w_n;
            w_n = pl_lds32(pp + (unsigned)(i + 1) * 16u);   // one pixel ahead, off the chain (slot T is readable padding)
            const int o = (int)(w & 255u), a = (int)((w >> 8) & 255u);
            const int err_in = pl_sext16(((int)w >> 16) + carry_a);
            // predictor; wrap (:175-182): the exact symbol orig - predicted, brought into [-128, 127], is the signed
            // low byte of the difference; `pred` is the predictor shifted by the same multiple of 256
            const int plin = ((a & ma) + (left & ml)) >> sh;
            const int ppae = pl_paeth(a, aprev, left);
            const int pred0 = (int)(((unsigned)ppae & pm) | ((unsigned)plin & ~pm));
            const int ex0 = pl_sext8(o - pred0);
            const int pred = o - ex0;
            const int want = ex0 + err_in;                      // here - pred
            const bool transp = al && o == 0;                   // fully transparent stays transparent (:158-164)
            int back, diff;

            if (table) {
                // ================================ FAST path ================================
                const unsigned bw = pl_lds32(band_sh + (((unsigned)(want + PL_S_BOFF)) & (2u * PL_S_BOFF - 1u)) * 4u);
                const unsigned long long xe64 = pl_lds64(hk_sh + ((unsigned)ex0 & 255u) * 8u);
                const unsigned t8 = ((bw & 0x8000u) && (unsigned)pred <= 255u) ? (unsigned)tz : (bw & 0x7ffu);
                const unsigned long long be64 = pl_lds64(bm_sh + t8 * 8u);
                const bool lutok = (unsigned)(want + PL_S_BOFF) < 2u * PL_S_BOFF, tv = t8 != 0x7ffu;
                const int bl = ((int)bw >> 16) + pred;          // the band in byte values: [bl, bl + q]
                const unsigned be_x = (unsigned)be64, base_l = (unsigned)(be64 >> 32);
                const int wbyte = bl + 127 - (int)(be_x & 127u);
                // The clamped band in byte values: [lob, hib].  A band that the byte range cuts down to one value - or
                // that lies outside of it altogether and collapses onto 0 or 255 (:195-210) - has one admissible symbol,
                // whatever the tables say; so has a fully transparent pixel (its band is {0}).
                const int lob = transp ? 0 : min(max(bl, 0), 255), hib = transp ? 0 : min(max(bl + q, 0), 255);
                const bool single = lob == hib;
                const int fs = lob - pred;                      // ... that symbol
                const unsigned long long fx64 = pl_lds64(hk_sh + ((unsigned)fs & 255u) * 8u);
                const bool inr = tv && (unsigned)wbyte <= 255u;
                // the exact symbol wins against the bucket winner iff its (count, rank) is not smaller (:228-244)
                const bool fe = (unsigned)(o - bl) <= (unsigned)q;
                const unsigned ce = (unsigned)(xe64 >> 32), re = (unsigned)xe64 >> PL_KEY_RANK_SHIFT;
                const unsigned ke = ((ce - base_l) << 8) | re, kw = be_x >> 7;
                const bool e_wins = fe && ce >= base_l && ke >= kw;
                back = single ? lob : e_wins ? o : wbyte;
                int sym = back - pred;
                // (the forced symbol's count and rank arrive late - their load needed the band - and are kept off the
                // chain: they are only used behind the vote)
                const unsigned bc_ns = e_wins ? ce : base_l + (be_x >> PL_BM_COUNT_SHIFT);
                // what the commit needs from the tables, for the provisional symbol (it is final unless the channel order
                // is replayed below): issued here, the loads are back by the time the vote is
                PlBinLoaded ld = pl_solo_bin_load(bins_sh, bins3_sh, sym, three);
                diff = (transp || !act) ? 0 : want - sym;   // = here - back; beyond the band's width only where it collapsed
                const bool tap_ok = (unsigned)(diff + PL_S_TAPC_HALF) < 2u * PL_S_TAPC_HALF;
                uint32_t te = pl_lds32(tapc_sh + (tap_ok ? diff : 0) * 4);
                // channel order (see pl_row_pass "fix-up"): does a symbol chosen by an earlier channel disturb this one?
                // The clamped band in symbols is [lo, lo + span]; a bin v lies in it iff ((v - lo) & 255) <= span.
                const int lo = lob - pred;
                const unsigned span = (unsigned)(hib - lob);
                // What the other channels see of this one: its symbol and its count.  A one-value band publishes the
                // largest count instead of its own (late) one: that can only make a later channel's test more
                // conservative, and this lane's own test cannot fire (a symbol inside a one-value band is the same symbol).
                const unsigned bf = single ? 0xfffff0u : min(bc_ns, 0xfffff0u);
                const unsigned mine = (bf << 8) | ((unsigned)sym & 255u);
                const unsigned mhi = mine & ~255u, nlo24 = (unsigned)(-lo) << 24, span24 = span << 24;
                bool conflict = false;
                unsigned dup = 0;
#pragma unroll
                for (int t2 = 0; t2 < 3; t2++) {
                    const unsigned theirs = __shfl_sync(PL_FULL, mine, src0 + t2);
                    const bool same = ((theirs ^ mine) & 255u) == 0u;
                    const bool inb = (theirs << 24) + nlo24 <= span24;
                    const bool near = theirs + 768u >= mhi;       // their count + 3 >= my winner's count
                    const bool earlier = (emask >> t2) & 1u;
                    conflict |= earlier & inb & !same & near;
                    dup += (unsigned)(earlier & same);
                }
                const bool fails = act && (!lutok || !tap_ok || (!single && !inr));
#if PL_S_ONEVOTE
                // one vote for the common case; the rare cases are told apart behind it
                bool any_fail = false, any_conflict = false;
                if (__any_sync(PL_FULL, fails || conflict)) {
                    any_fail = __any_sync(PL_FULL, fails);
                    any_conflict = __any_sync(PL_FULL, conflict);
                }
#else
                const bool any_fail = __any_sync(PL_FULL, fails);
                const bool any_conflict = __any_sync(PL_FULL, conflict);
#endif
                if (!any_fail) {
                    PL_EMU_COUNT(PL_CNT_SOLO_FAST);
                    unsigned bc = single ? (unsigned)(fx64 >> 32) : bc_ns;
                    unsigned rk = single ? (unsigned)fx64 >> PL_KEY_RANK_SHIFT : e_wins ? re : (kw & 255u);
                    if (!any_conflict) {
                        bc += dup;   // every provisional winner is final; mine has been counted dup times since the look-up
                    } else {
                        // exact sequential replay of the channel order (flat colours: two channels whose bands
                        // coincide leapfrog each other's counts): key order = count, rank, exact symbol, earlier position
                        PL_EMU_COUNT(PL_CNT_FIXUP_REPLAY);
#pragma unroll 1
                        for (int t2 = 0; t2 < 3; t2++) {
                            const int vsym = __shfl_sync(PL_FULL, sym, src0 + t2);
                            const unsigned vc = __shfl_sync(PL_FULL, bc, src0 + t2);
                            const unsigned vrk = __shfl_sync(PL_FULL, rk, src0 + t2);
                            if ((emask >> t2) & 1u) {
                                const unsigned pos = (unsigned)(vsym - lo) & 255u;
                                if (pos <= span) {
                                    if (((unsigned)(vsym ^ sym) & 255u) == 0u) {
                                        bc += 1u;
                                    } else {
                                        const int v = lo + (int)pos;   // their bin as a symbol of my band
                                        const unsigned c = vc + 1u;
                                        const unsigned kv = (vrk << 10) | ((unsigned)(v == ex0) << 9) | (511u - pos);
                                        const unsigned km = (rk << 10) | ((unsigned)(sym == ex0) << 9) | (511u - (unsigned)(sym - lo));
                                        if (c > bc || (c == bc && kv > km)) {
                                            bc = c;
                                            rk = vrk;
                                            sym = v;
                                        }
                                    }
                                }
                            }
                        }
                        back = sym + pred;
                        ld = pl_solo_bin_load(bins_sh, bins3_sh, sym, three);
                        diff = (transp || !act) ? 0 : want - sym;   // (a replayed symbol lies in the band: within the table)
                        te = pl_lds32(tapc_sh + diff * 4);
                    }
                    pl_solo_commit(hk_sh, bm_sh, ld, sym, bc + 1u, rk << 7, actm, sink_sh, three);
                    back &= (int)actm;
                    pl_sts32(po + (unsigned)i * 16u, ((unsigned)back & 255u) | ((unsigned)diff << 16));
                    left = back;
                    aprev = a;
                    // the two Sierra taps that stay in this row (reference :390-467): rem -> pixel x+1, threes -> x+2
                    carry_a = (int)(short)(te & 0xffffu) + carry_b;
                    carry_b = (int)te >> 16;
                    __syncwarp();
                    continue;
                }
            }

            {
                // ================================ GENERAL path (see pl_lean_row_pass) ================================
                PL_EMU_COUNT(PL_CNT_SOLO_GENERAL);
                general_px++;
                int ex = ex0;
                const unsigned m = (unsigned)(want < 0 ? -want : want);
                const unsigned kq = pl_udiv_magic(m, step_magic);
                const bool neg = want < 0;
                const int ks = (int)kq * step;
                const int lo_u = neg ? -(ks + q) : ks;
                const int smin = -pred, smax = 255 - pred;
                int lo = min(max(lo_u, smin), smax);
                int hi = min(max(lo_u + q, smin), smax);
                if (transp) lo = hi = ex = smin;
                const int span = act ? hi - lo : -1;

                const bool zband = neg && kq == 0u && (unsigned)pred <= 255u;
                const bool tvalid = !transp && kq < (unsigned)(neg ? N1 : P1);
                const int tl = zband ? tz : (int)kq + (neg ? P1 : 0);
                const unsigned long long be64 = pl_lds64(bm_sh + (unsigned)(tvalid ? tl : 0) * 8u);
                const unsigned long long xe64 = pl_lds64(hk_sh + ((unsigned)ex & 255u) * 8u);
                const unsigned be_x = (unsigned)be64, base_l = (unsigned)(be64 >> 32);
                const int wsym = lo_u + 127 - (int)(be_x & 127u);
                const bool inr = act && tvalid && wsym >= lo && wsym <= hi;
                unsigned bc = inr ? base_l + (be_x >> PL_BM_COUNT_SHIFT) : 0u;
                unsigned bl = inr ? ((((be_x >> 7) & 255u) << PL_KEY_RANK_SHIFT) | (unsigned)(511 - (wsym - lo))) : 0u;
                const bool need_scan = act && !inr && !(span == 0 && ex == lo);
#ifdef PL_SIMT_EMU
                if (__any_sync(PL_FULL, need_scan)) PL_EMU_COUNT(PL_CNT_BM_SCAN);
                else PL_EMU_COUNT(PL_CNT_BM_LOOKUP);
#endif
                if (need_scan) {   // clamped bands, strengths without a table; diverged
                    for (int p = 0; p <= span; p++) {
                        const unsigned long long e = pl_lds64(hk_sh + ((unsigned)(lo + p) & 255u) * 8u);
                        const unsigned c = (unsigned)(e >> 32), l = (unsigned)e | (unsigned)(511 - p);
                        if (c > bc || (c == bc && l > bl)) {
                            bc = c;
                            bl = l;
                        }
                    }
                }
                {   // the exact symbol, with its bonus bit (:228-244)
                    const unsigned c = (unsigned)(xe64 >> 32);
                    const int pos = ex - lo;
                    const unsigned l = (unsigned)xe64 | pl_key_low(true, pos & 255);
                    if (pos >= 0 && pos <= span && (c > bc || (c == bc && l > bl))) {
                        bc = c;
                        bl = l;
                    }
                }
                int bpos = 511 - (int)(bl & 511u);

                // ---- fix-up: replay the channel order (see pl_row_pass) ------------------------------------------
                bool conflict = false;
                unsigned dup = 0;
                const unsigned bf = min(bc, 0xfffff0u);
                {
                    const unsigned mine = (bf << 8) | ((unsigned)(lo + bpos) & 255u);
#pragma unroll
                    for (int t2 = 0; t2 < 3; t2++) {
                        const unsigned theirs = __shfl_sync(PL_FULL, mine, src0 + t2);
                        const int pos = ((int)(theirs & 255u) - lo) & 255;
                        const bool earlier = (emask >> t2) & 1u;
                        conflict |= earlier & (pos <= span) & (pos != bpos) & ((theirs >> 8) + 3u >= bf);
                        dup += (unsigned)(earlier & (pos == bpos));
                    }
                }
                if (!__any_sync(PL_FULL, conflict)) {
                    PL_EMU_COUNT(PL_CNT_FIXUP_SKIPPED);
                    bc += dup;
                } else {
                    PL_EMU_COUNT(PL_CNT_FIXUP_REPLAY);
#pragma unroll 1
                    for (int t2 = 0; t2 < 3; t2++) {
                        const int vsym = __shfl_sync(PL_FULL, lo + bpos, src0 + t2);
                        const unsigned vc = __shfl_sync(PL_FULL, bc, src0 + t2);
                        const unsigned vl = __shfl_sync(PL_FULL, bl, src0 + t2);
                        if ((emask >> t2) & 1u) {
                            const int pos = (vsym - lo) & 255;
                            if (pos <= span) {
                                if (pos == bpos) {
                                    bc += 1u;
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

                // ---- commit the byte ------------------------------------------------------------------------------
                const int sym = lo + bpos;
                diff = (act && !transp) ? pl_sext16(want - sym) : 0;   // here - back (0 for a transparent pixel)
                back = act ? sym + pred : 0;
                pl_solo_commit(hk_sh, bm_sh, pl_solo_bin_load(bins_sh, bins3_sh, sym, three), sym, bc + 1u,
                               (bl >> (PL_KEY_RANK_SHIFT - 7)) & (255u << 7), actm, sink_sh, three);
            }

            pl_sts32(po + (unsigned)i * 16u, ((unsigned)back & 255u) | ((unsigned)diff << 16));
            left = back;
            aprev = a;

            // ---- the two Sierra taps that stay in this row (reference :390-467): rem -> pixel x+1, threes -> x+2 ---
            int rem, threes;
            if ((unsigned)(diff + PL_S_TAPC_HALF) < 2u * PL_S_TAPC_HALF) {
                PL_EMU_COUNT(PL_CNT_TAPS_TABLE);
                const uint32_t e = pl_lds32(tapc_sh + diff * 4);
                rem = (int)(short)(e & 0xffffu);
                threes = (int)e >> 16;
            } else {
                PL_EMU_COUNT(PL_CNT_TAPS_COMPUTED);
                const PlTaps tp = pl_sierra_taps(diff, bleed_magic);
                rem = tp.rem;
                threes = tp.threes;
            }
            carry_a = rem + carry_b;
            carry_b = threes;
            __syncwarp();
        }
        // hand the tile to the post warps, the input slot back to the producer
        if (PL_S_ARRIVES(lane)) {
            pl_mbar_arrive(&sm.out_full[s]);
            pl_mbar_arrive(&sm.pre_empty[s]);
        }
    }
    return general_px;
}

// ---- producer warp: one lane per pixel of a tile ------------------------------------------------------------------
__device__ __forceinline__ void pl_solo_producer(PlSoloSmem &sm, int W, int y, int parity, int prev_w, unsigned use) {
    const int lane = threadIdx.x & 31;
    const PlImageDev &im = sm.img;
    const int EW = W + PL_ERR_PAD;
    const uchar4 *rin = im.in + (size_t)y * W;
    const uchar4 *rout_up = im.out + (size_t)(y ? y - 1 : 0) * W;
    const short4 *E0 = im.err + ((size_t)(parity * PL_FILTERS + prev_w) * 2 + 0) * EW;
    const int ntiles = (W + PL_S_T - 1) / PL_S_T;
    for (int t = 0; t < ntiles; t++) {
        const unsigned u = use + (unsigned)t;
        const int s = (int)(u % PL_S_STAGES);
        const unsigned ph = (u / PL_S_STAGES) & 1u;
        const int x = t * PL_S_T + lane;
        unsigned o4 = 0, n4 = 0;
        short4 e = make_short4(0, 0, 0, 0);
        if (x < W) {
            o4 = pl_u32(rin[x]);
            if (y) {
                n4 = pl_u32(rout_up[x]);
                e = E0[x + 2];
            }
        }
        pl_mbar_wait_relaxed(&sm.pre_empty[s], ph ^ 1u);
        uint4 v;
        v.x = (o4 & 255u) | ((n4 & 255u) << 8) | ((unsigned)(unsigned short)e.x << 16);
        v.y = ((o4 >> 8) & 255u) | (((n4 >> 8) & 255u) << 8) | ((unsigned)(unsigned short)e.y << 16);
        v.z = ((o4 >> 16) & 255u) | (((n4 >> 16) & 255u) << 8) | ((unsigned)(unsigned short)e.z << 16);
        v.w = (o4 >> 24) | ((n4 >> 24) << 8) | ((unsigned)(unsigned short)e.w << 16);
        sm.pre[s][lane] = v;
        __syncwarp();
        if (PL_S_ARRIVES(lane)) pl_mbar_arrive(&sm.pre_full[s]);
    }
}

// ---- post warps: one lane per error cell / pixel ---------------------------------------------------------------------
struct PlPostAcc {
    unsigned long long derr;
    unsigned as0, as1, as2, as3, as4;
};
// cells / pixels t * 32 .. t * 32 + 31 of candidate pf
__device__ __forceinline__ void pl_solo_post_tile(PlSoloSmem &sm, int pf, int t, int chmask, int W, int y, int parity,
                                                  int prev_w, bool adaptive, unsigned bleed_magic, unsigned use,
                                                  PlPostAcc &acc) {
    const int lane = threadIdx.x & 31;
    const PlImageDev &im = sm.img;
    const int EW = W + PL_ERR_PAD;
    const bool first = (y == 0);
    const unsigned canon = chmask == 0xF ? 0x3210u : chmask == 0x7 ? 0x4210u : chmask == 0xA ? 0x3111u : 0x4111u;
    const short4 *Ecur1 = im.err + ((size_t)(parity * PL_FILTERS + prev_w) * 2 + 1) * EW;
    short4 *En0 = im.err + ((size_t)((parity ^ 1) * PL_FILTERS + pf) * 2 + 0) * EW;
    short4 *En1 = En0 + EW;
    const uchar4 *rin = im.in + (size_t)y * W;
    const uchar4 *rin_up = im.oprev;
    const uchar4 *rout_up = im.out + (size_t)(y ? y - 1 : 0) * W;
    uchar4 *rcand = im.cand + (size_t)pf * W;
    const int c = t * PL_S_T + lane;
    // the chain's words of pixels c, c-1, .. c-4 (tile t or t-1: released one tile late, so still there)
    int n0[4], n1[4];
    unsigned q4 = 0, ql4 = 0;
    if (!first && c < W + 4) {
        const short4 e1 = Ecur1[c];
        n0[0] = e1.x; n0[1] = e1.y; n0[2] = e1.z; n0[3] = e1.w;
    } else {
        n0[0] = n0[1] = n0[2] = n0[3] = 0;
    }
    n1[0] = n1[1] = n1[2] = n1[3] = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int p = c - k;
        if (p >= 0 && p < W) {
            const unsigned up = use + (unsigned)(p / PL_S_T);
            const uint4 wv = sm.outw[up % PL_S_STAGES][pf][p % PL_S_T];
            const unsigned wd[4] = {wv.x, wv.y, wv.z, wv.w};
            if (k <= 1) {
                const unsigned b4 = (wv.x & 255u) | ((wv.y & 255u) << 8) | ((wv.z & 255u) << 16) | ((wv.w & 255u) << 24);
                if (k == 0) q4 = b4;
                else ql4 = b4;
            }
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                const int diff = (int)wd[cc] >> 16;
                PlTaps tp;
                if ((unsigned)(diff + PL_DL32_HALF) < 2u * PL_DL32_HALF) tp = pl_unpack_taps6(sm.dl32[diff + PL_DL32_HALF]);
                else tp = pl_sierra_taps(diff, bleed_magic);
                // next error row 0 (reference row 1): twos, fours, five, fours, twos at cells x .. x+4;
                // next error row 1 (reference row 2): twos, threes, twos at cells x+1 .. x+3
                if (k == 0) n0[cc] += tp.twos;
                if (k == 1) { n0[cc] += tp.fours; n1[cc] += tp.twos; }
                if (k == 2) { n0[cc] += tp.five; n1[cc] += tp.threes; }
                if (k == 3) { n0[cc] += tp.fours; n1[cc] += tp.twos; }
                if (k == 4) n0[cc] += tp.twos;
            }
        }
    }
    if (c < W + 4) {
        En0[c] = make_short4((short)n0[0], (short)n0[1], (short)n0[2], (short)n0[3]);
        En1[c] = make_short4((short)n1[0], (short)n1[1], (short)n1[2], (short)n1[3]);
    }
    if (c < W) {
        const int x = c;
        rcand[x] = pl_uc4(q4);
        const unsigned o4 = pl_u32(rin[x]);
        unsigned oa4 = 0, na4 = 0, ol4 = 0, oad4 = 0, nad4 = 0;
        if (!first) {
            oa4 = pl_u32(rin_up[x]);
            na4 = pl_u32(rout_up[x]);
        }
        if (x > 0) {
            ol4 = pl_u32(rin[x - 1]);
            if (!first) {
                oad4 = pl_u32(rin_up[x - 1]);
                nad4 = pl_u32(rout_up[x - 1]);
            }
        } else {
            ql4 = 0;
        }
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
            acc.derr += xs + 2u * zs - 2u * ys;
        }
        if (adaptive) {
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                if ((chmask >> cc) & 1) {
                    const int qc = pl_byte(q4, cc);
                    const int lq = pl_byte(ql4, cc), aq = pl_byte(na4, cc), dq = pl_byte(nad4, cc);
                    acc.as0 += pl_absres(qc, 0);
                    acc.as1 += pl_absres(qc, lq);
                    acc.as2 += pl_absres(qc, aq);
                    acc.as3 += pl_absres(qc, (aq + lq) >> 1);
                    acc.as4 += pl_absres(qc, pl_paeth(aq, dq, lq));
                }
            }
        }
    }
}

// A post warp's row: candidates pf0, pf0 + pf_step, ... (NF of them at most), tile by tile behind the chain.
template <int NF>
__device__ __forceinline__ void pl_solo_post(PlSoloSmem &sm, int pf0, int pf_step, int chmask, int W, int y, int parity,
                                             int prev_w, bool adaptive, unsigned bleed_magic, unsigned use) {
    const int lane = threadIdx.x & 31;
    const int ntiles = (W + PL_S_T - 1) / PL_S_T;
    const int ncell_tiles = (W + 4 + PL_S_T - 1) / PL_S_T;   // the error rows are four cells longer than the row
    PlPostAcc acc[NF];
#pragma unroll
    for (int k = 0; k < NF; k++) acc[k] = PlPostAcc{0ull, 0u, 0u, 0u, 0u, 0u};

    for (int t = 0; t < ncell_tiles; t++) {
        if (t < ntiles) {
            const unsigned u = use + (unsigned)t;
            pl_mbar_wait_relaxed(&sm.out_full[u % PL_S_STAGES], (u / PL_S_STAGES) & 1u);
        }
        // (the chain's words of pixels c .. c-4 lie in tile t or t-1: tiles are released one tile late, so still there)
#pragma unroll
        for (int k = 0; k < NF; k++) {
            const int pf = pf0 + k * pf_step;
            if (pf < PL_FILTERS)
                pl_solo_post_tile(sm, pf, t, chmask, W, y, parity, prev_w, adaptive, bleed_magic, use, acc[k]);
        }
        // give the previous tile back to the chain
        __syncwarp();
        if (PL_S_ARRIVES(lane) && t >= 1 && t - 1 < ntiles) pl_mbar_arrive(&sm.out_empty[(use + (unsigned)(t - 1)) % PL_S_STAGES]);
    }
    if (PL_S_ARRIVES(lane) && ncell_tiles == ntiles) pl_mbar_arrive(&sm.out_empty[(use + (unsigned)(ntiles - 1)) % PL_S_STAGES]);

#pragma unroll
    for (int k = 0; k < NF; k++) {
        const int pf = pf0 + k * pf_step;
        if (pf >= PL_FILTERS) continue;
        PlPostAcc a = acc[k];
#pragma unroll
        for (int mk = 1; mk < 32; mk <<= 1) {
            a.derr += __shfl_xor_sync(PL_FULL, a.derr, mk);
            a.as0 += __shfl_xor_sync(PL_FULL, a.as0, mk);
            a.as1 += __shfl_xor_sync(PL_FULL, a.as1, mk);
            a.as2 += __shfl_xor_sync(PL_FULL, a.as2, mk);
            a.as3 += __shfl_xor_sync(PL_FULL, a.as3, mk);
            a.as4 += __shfl_xor_sync(PL_FULL, a.as4, mk);
        }
        if (lane == 0) {
            sm.derr[pf] = a.derr;
            sm.asum[pf][0] = a.as0;
            sm.asum[pf][1] = a.as1;
            sm.asum[pf][2] = a.as2;
            sm.asum[pf][3] = a.as3;
            sm.asum[pf][4] = a.as4;
        }
    }
}

template <int FPW, bool COMPACT>
__global__ void __launch_bounds__(PlSoloCfg<FPW, COMPACT>::THREADS, PlSoloCfg<FPW, COMPACT>::MIN_BLOCKS)
pl_k2_solo(const PlImageDev *imgs, const int *slots, int strength, int bleed, unsigned sm_count) {
    typedef PlSoloCfg<FPW, COMPACT> C;
    PL_DYN_SMEM(smem_raw);
    PlSoloSmem &sm = *(PlSoloSmem *)pl_align_shared(smem_raw, 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int idx = slots[blockIdx.x];
    if (idx < 0) return;
    // (CTAs are placed round-robin over the SMs, so the CTAs that share an SM differ in blockIdx / #SMs)
    const int role = C::role(warp, (int)((blockIdx.x / sm_count) & 3u));
    const int pw = role >= 0 ? role : -1;   // post warp index (-1: none): takes candidates pw, pw + NPOSTW, ...

    // ---- set-up ------------------------------------------------------------------------------------------------
    if (tid == 0) {
        sm.img = imgs[idx];
        for (int s = 0; s < PL_S_STAGES; s++) {
            pl_mbar_init(&sm.pre_full[s], PL_S_ARRIVERS);
            pl_mbar_init(&sm.pre_empty[s], C::NCHAIN * PL_S_ARRIVERS);
            pl_mbar_init(&sm.out_full[s], C::NCHAIN * PL_S_ARRIVERS);
            pl_mbar_init(&sm.out_empty[s], C::NPOSTW * PL_S_ARRIVERS);
        }
        pl_fence_mbar_init();
    }
    __syncthreads();
    const PlImageDev &im = sm.img;
    const int W = (int)im.width, H = (int)im.height;
    const int mode = pl_image_mode(im);
    const int chmask = PL_MODE_MASK(mode);
    const bool alpha_rule = (mode & 1) == 0;
    // original_frequency of the image's colour mode (K1 counted every RGBA channel separately), staged in the low
    // words; then replaced by its rank (see pl_k2_quantize)
    for (int k = tid; k < PL_FILTERS * 256; k += C::THREADS) {
        const int f = k >> 8, s = k & 255;
        unsigned v = 0;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if ((chmask >> c) & 1) v += im.chan_hist[(f * 4 + c) * 256 + s];
        sm.hk[f][s] = v;
    }
    for (int k = tid; k < 256; k += C::THREADS) sm.base[k] = 0;
    __syncthreads();
    unsigned my_rank[(PL_FILTERS * 256 + C::THREADS - 1) / C::THREADS];
    {
        int n = 0;
        for (int k = tid; k < PL_FILTERS * 256; k += C::THREADS, n++) {
            const int f = k >> 8, s = k & 255;
            const unsigned mine = (unsigned)sm.hk[f][s];
            unsigned rank = 0;
            for (int s2 = 0; s2 < 256; s2++) rank += (unsigned)((unsigned)sm.hk[f][s2] < mine);
            my_rank[n] = rank;
        }
    }
    __syncthreads();
    {
        int n = 0;
        for (int k = tid; k < PL_FILTERS * 256; k += C::THREADS, n++)
            sm.hk[k >> 8][k & 255] = (unsigned long long)(my_rank[n] << PL_KEY_RANK_SHIFT);   // count 0
    }
    const unsigned bleed_magic = pl_make_magic((unsigned)bleed);
    for (int k = tid; k < 2 * PL_S_TAPC_HALF; k += C::THREADS) {
        const PlTaps tp = pl_sierra_taps(k - PL_S_TAPC_HALF, bleed_magic);
        sm.tapc[k] = ((unsigned)tp.rem & 0xffffu) | ((unsigned)tp.threes << 16);
    }
    for (int k = tid; k < 2 * PL_DL32_HALF; k += C::THREADS)
        sm.dl32[k] = pl_pack_taps6(pl_sierra_taps(k - PL_DL32_HALF, bleed_magic));
    __syncthreads();

    int prev_w = 0;
    bool failed = false;
    unsigned retries = 0;
    unsigned use = 0;       // ring tiles consumed so far (CTA-uniform)
    const int ntiles = (W + PL_S_T - 1) / PL_S_T;
    bool try_fast = true;   // chain warps: attempt the fast path on the next row
    int band_q = -1;        // strength sm.band was built for

    for (int y = 0; y < H && !failed; y++) {
        const bool adaptive = im.adaptive_all || y == 0;   // reference src/pngloss_image.c:210
        int q = strength;
        for (;;) {
            const int step = q + 1;
            const unsigned step_magic = pl_make_magic((unsigned)step);
            const PlBm bmc = pl_solo_bm_counts(step, W);
            // ---- tables of the pass.  Bucket winners of every candidate at the start of the row (all candidates hold
            // the same counts, the tie-break rank differs), see pl_lean_row_pass; entry tz ("Z+") is the whole zero
            // band [-q, 0], negative bucket 0 is [-q, -1] ---------------------------------------------------------------
            {
                const int nb = bmc.P1 + bmc.N1 + (bmc.P1 > 0 ? 1 : 0), tz = bmc.P1 + bmc.N1;
                for (int k = tid; k < PL_FILTERS * nb; k += C::THREADS) {
                    const int f = k / nb, t = k - f * nb;
                    const int lo_t = t == tz ? -q : pl_bm_low(bmc, t, step);
                    const int p1 = t == bmc.P1 ? q - 1 : q;
                    unsigned long long best = 0;
                    for (int p = 0; p <= p1; p++) {
                        const unsigned long long key = sm.hk[f][(unsigned)(lo_t + p) & 255u] | (unsigned)(511 - p);
                        best = key > best ? key : best;
                    }
                    const unsigned mcount = (unsigned)(best >> 32);
                    const unsigned base = mcount > 4u * (unsigned)W ? mcount - 4u * (unsigned)W : 0u;
                    sm.bmk[f][t] = make_uint2(pl_bm_key(mcount - base, ((unsigned)best >> PL_KEY_RANK_SHIFT) & 255u,
                                                        511 - (int)((unsigned)best & 511u)),
                                              base);
                }
                if (q != band_q) {   // the band of every here - predicted: changes with the strength only
                    for (int k = tid; k < 2 * PL_S_BOFF; k += C::THREADS)
                        sm.band[k] = pl_solo_band_entry(bmc, k - PL_S_BOFF, q, step, step_magic);
                    band_q = q;
                }
            }
            __syncthreads();
            // ... and the table entries of every histogram bin, with their base counts (the same for all candidates)
            for (int k = tid; k < 256; k += C::THREADS) {
                const PlBinEntries be = pl_solo_bin_entries(sm, bmc, k, q, step, step_magic);
                sm.bins[k] = make_uint4(be.e[0], be.base[0], be.e[1], be.base[1]);
                sm.bins3[k] = make_uint2(be.e[2], be.base[2]);
            }
            __syncthreads();

            // ---- the row, by role -----------------------------------------------------------------------------
            if (role == -2) {
                const int f = FPW == 5 ? lane >> 2 : C::chain_index(warp);
                const bool lane_act = FPW == 5 ? lane < 4 * PL_FILTERS : lane < 4;
                const unsigned g = pl_solo_chain<FPW>(sm, f, lane & 3, lane_act, chmask, alpha_rule, q, step_magic, W,
                                                      bleed_magic, use, try_fast);
                // images where the fast path mostly fails skip the attempt (and retry it every 16th row)
                try_fast = 4u * g <= 3u * (unsigned)W || (y & 15) == 15;
            } else if (role == -1) {
                pl_solo_producer(sm, W, y, y & 1, prev_w, use);
            } else if (pw >= 0) {
                pl_solo_post<C::NF>(sm, pw, C::NPOSTW, chmask, W, y, y & 1, prev_w, adaptive, bleed_magic, use);
            }
            use += (unsigned)ntiles;
            __syncthreads();

            // ---- row cost (reference src/optimize_state.c:314-360), see pl_row_pass: a post warp takes its candidate
            if (pw >= 0) {
                for (int pf = pw; pf < PL_FILTERS; pf += C::NPOSTW) {
                    unsigned bits = 0;
                    for (int s = lane; s < 256; s += 32) {
                        const unsigned hv = (unsigned)(sm.hk[pf][s] >> 32);
                        bits += (hv - sm.base[s]) * (33u + (unsigned)__clz((int)hv));
                    }
#pragma unroll
                    for (int mk = 1; mk < 32; mk <<= 1) bits += __shfl_xor_sync(PL_FULL, bits, mk);
                    if (lane == 0) sm.bits[pf] = bits;
                }
            }
            __syncthreads();
            // ---- the winner (reference src/pngloss_image.c:257-263): strict < in filter order ---------------------
            int w = -1;
            {
                unsigned long long best = ~0ull;
#pragma unroll
                for (int f = 0; f < PL_FILTERS; f++) {
                    unsigned long long c = sm.derr[f] / 128ull + sm.bits[f];
                    if (adaptive) {   // libpng picks the first minimum in the order none, sub, up, average, paeth (:531-559)
                        const unsigned a0 = sm.asum[f][0], a1 = sm.asum[f][1], a2 = sm.asum[f][2], a3 = sm.asum[f][3],
                                       a4 = sm.asum[f][4];
                        const unsigned lowest = min(min(min(a0, a1), min(a2, a3)), a4);
                        const int pick = lowest >= a0 ? 0 : lowest >= a1 ? 1 : lowest >= a2 ? 2 : lowest >= a3 ? 3 : 4;
                        if (pick != f) c = ~0ull;
                    }
                    if (c < best) { best = c; w = f; }
                }
            }
            // ---- commit (see pl_commit_row) or restore ---------------------------------------------------------------
            if (w >= 0) {
                const bool notgray = mode >= 3, notopaque = (mode & 1) == 0;
                const unsigned sel = notgray ? 0x3210u : 0x3111u, amask = notopaque ? 0u : 0xff000000u;
                const uchar4 *src = im.cand + (size_t)w * W;
                const uchar4 *orig = im.in + (size_t)y * W;
                uchar4 *dst = im.out + (size_t)y * W;
                if ((W & 3) == 0) {
                    const uint4 *src4 = (const uint4 *)src, *orig4 = (const uint4 *)orig;
                    uint4 *dst4 = (uint4 *)dst, *oprev4 = (uint4 *)im.oprev;
                    for (int x = tid; x < W / 4; x += C::THREADS) {
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
                    for (int x = tid; x < W; x += C::THREADS) {
                        const unsigned p = __byte_perm(pl_u32(src[x]), 0u, sel) | amask;
                        im.oprev[x] = orig[x];
                        dst[x] = pl_uc4(p);
                    }
                }
                for (int s = tid; s < 256; s += C::THREADS) {
                    const unsigned v = (unsigned)(sm.hk[w][s] >> 32);
                    sm.base[s] = v;
#pragma unroll
                    for (int f = 0; f < PL_FILTERS; f++) ((unsigned *)&sm.hk[f][s])[1] = v;
                }
                if (tid == 0) im.filters[y] = (unsigned char)(0x08 << w);
                prev_w = w;
            } else {
                for (int s = tid; s < 256; s += C::THREADS) {
                    const unsigned v = sm.base[s];
#pragma unroll
                    for (int f = 0; f < PL_FILTERS; f++) ((unsigned *)&sm.hk[f][s])[1] = v;
                }
            }
            __syncthreads();
            if (w >= 0) break;
            if (q == 0) {          // reference aborts here (src/pngloss_image.c:268-271)
                failed = true;
                break;
            }
            q -= 1;                // try again at lower quantization strength (:273-274)
            retries++;
        }
    }

    // ---- results ---------------------------------------------------------------------------------------------------
    for (int s = tid; s < 256; s += C::THREADS) im.final_hist[s] = sm.base[s];
    if (tid == 0) {
        im.status[0] = failed ? PL_ST_NO_ROW : PL_ST_OK;
        im.status[1] = (unsigned)mode;
        im.status[2] = retries;
    }
}
