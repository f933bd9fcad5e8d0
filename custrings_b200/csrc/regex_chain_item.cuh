// Item-buffered linear-chain kernel (boolean results: contains_re / match / literal contains) — included by regex_item.cu
// inside namespace custr::bits, after regex_chain64.cuh (whose stream helpers, plan view and UTF-8 classifier it shares).
//
// k_chain64 walks the offsets array once per 2 KiB window: ROWSTART scatter, look-ahead byte, row finalisation, NUL
// bookkeeping — about 190 of its 567 warp-instructions per window (ncu source page, profiles/r2_*).  This kernel takes the
// row bookkeeping OUT of the window loop.  A work item (a contiguous row range of about 32 KiB of chars, as before) is
// processed in SEGMENTS of up to SEG_WINS windows whose bit streams live in shared memory for the whole segment:
//
//   phase 0   ROWSTART bits of every row start of the segment are scattered into `bits` (red.shared.or), 32 rows per step;
//   phase A   window loop — per lane: 64 staged bytes -> 8 bit planes -> class streams -> marker chain -> match-end stream
//             E -> per-row sticky OR  F = spread(E, ~ROWSTART), which REPLACES the window's ROWSTART word in `bits`.
//             No offsets, no result stores, no look-ahead load in here;
//   phase B   one pass over the rows that end in the segment: result = bit of F at the row's last byte, 32 rows per step,
//             every row of the item is written (no pre-clearing memset of the result array).
//
// What else changed against k_chain64:
//   * look-ahead (\b / $ / "last byte of a character" behind the last byte of a window) is DEFERRED instead of loaded: when
//     no row starts right behind the window, the match-end bit of the window's last position is withheld and decided at the
//     top of the next window, where the class of the following byte is in registers anyway; it enters the sticky stream
//     through the carry.  (When a row does start there the look-ahead is "end of row" and nothing is deferred.)
//   * char staging: every cp.async instruction copies 512 CONTIGUOUS bytes (lane L: 16 bytes at 512 k + 16 L) — 4 global
//     lines per instruction instead of 16 — into a chunk-major, rotated layout (chunk kk of owner lane o at
//     kk * 512 + 128 (o >> 3) + 16 ((o + 2 kk) & 7)) that makes both the cp.async writes and the LDS.128 read-back
//     bank-conflict free with immediate offsets only.
//   * windows without non-ASCII bytes (warp-uniform test) run a chain body without the UTF-8 terms; windows with them run
//     the general body.  NUL bytes are only ACCUMULATED per lane; an item that saw one re-checks its rows in phase B (rare).
#pragma once

#ifdef CUSTR_ITEM_TIMING  // development aid (tools/build_variant.py T -DCUSTR_ITEM_TIMING): per-warp time stamps of the first items
constexpr int ITEM_TIMING_SLOTS = 4 + 4 * 6;  // [entry, first item known, -, -] then per item [start, phase 0 done, phase A done, phase B done]
__device__ unsigned long long g_item_times[148 * 3 * WARPS * ITEM_TIMING_SLOTS];
__device__ __forceinline__ unsigned long long item_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define ITEM_STAMP(slot)                                                                                        \
    do {                                                                                                        \
        if (L.lane == 0 && (slot) < ITEM_TIMING_SLOTS)                                                          \
            g_item_times[(blockIdx.x * WARPS + (threadIdx.x >> 5)) * ITEM_TIMING_SLOTS + (slot)] = item_now();  \
    } while (0)
#else
#define ITEM_STAMP(slot) do { } while (0)
#endif

#ifndef ITEM_SEG_WINS
#define ITEM_SEG_WINS 18
#endif
#ifndef ITEM_MIN_CTAS
#define ITEM_MIN_CTAS 3
#endif
constexpr int SEG_WINS = ITEM_SEG_WINS;                 // windows per segment: a 32 KiB item, its unaligned head and its last row
constexpr int SEG_WORDS = SEG_WINS * 32;     // 64-bit stream words per segment
struct __align__(128) WarpSmItem {
    char ring[RING_STAGES][WIN64];
    u64 bits[SEG_WORDS + 16];                // ROWSTART, then F, per window; one extra word: a row start right behind the segment
};
constexpr uint32_t ITEM_SM_BITS = RING_STAGES * WIN64;
constexpr int ITEM_SMEM_BYTES = WARPS * (int)sizeof(WarpSmItem);

template <int NS>
struct ItemState {  // top words of the previous window's streams (lane 31's copy is the one that is used)
    uint32_t last[NS];
    uint32_t last_al, last_nl, last_f;
    uint32_t pend;  // bit 31: a match may end at the previous window's last position, subject to the look-ahead assertions
};

static __device__ __noinline__ void ring_issue_item_tail(uint32_t wr, const char* __restrict__ chars, uint32_t ws, uint32_t end, uint32_t lane)
{
    for (uint32_t k = 0; k < 4; ++k) {
        const uint32_t pos = ws + 512u * k + 16u * lane;
        const uint32_t dst = wr + 128u * k;
        if (pos + 16u <= end) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(chars + pos) : "memory");
            continue;
        }
        // bytes behind the end of the buffer are staged as SPACES, not zeros: a zero would count as a NUL byte and send the rows
        // of the column's last item through the per-row NUL check of phase B (item_dirty_row) on every call — 75-100 us for one
        // warp while the rest of the GPU idles (tools/item_timing.py); a row start sits on `end`, so no row sees these bytes
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0x20202020u) : "memory");
        for (uint32_t q = pos; q < end; ++q) asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst + (q - pos)), "r"((uint32_t)(uint8_t)chars[q]) : "memory");
    }
}
// copy of window [ws, ws + 2048) into a ring stage: `wr` = this lane's write address in that stage, `gsrc` = chars + 16 * lane
__device__ __forceinline__ void ring_issue_item(uint32_t wr, const char* __restrict__ gsrc, const char* __restrict__ chars, uint32_t ws, uint32_t end,
                                                uint32_t lane)
{
    if (ws + WIN64 <= end) {
        const char* src = gsrc + ws;
#pragma unroll
        for (int k = 0; k < 4; ++k) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(wr + 128u * k), "l"(src + 512 * k) : "memory");
    } else
        ring_issue_item_tail(wr, chars, ws, end, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// x >> 1 with the bit behind the lane's word taken from `nxt` (low word of the next lane's / next window's stream word)
__device__ __forceinline__ u64 shift_down64_nb(u64 x, uint32_t nxt) { return mk64(funnel1r_fma(lo32(x), hi32(x)), funnel1r_fma(hi32(x), nxt)); }

// Marker chain of one window.  Returns the stream of positions behind which a match may end, BEFORE the zero-width
// assertions of END are applied.  UTF8 = false: the window holds ASCII bytes only (cont == 0, every byte is a character).
template <int NS, int NCLS, int SPEC, bool UTF8>
__device__ __forceinline__ u64 chain_item64(const ChainDev& cd, const u64 (&c)[NCLS], const Assertions64& as, u64 rs, int rounds, u64 cont, u64 lead,
                                            uint32_t cont0, ItemState<NS>& st, const LaneCtx& L)
{
    using PL = PlanLit<SPEC>;
    const u64 nrs = ~rs;
    u64 fin = ~0ull;  // last byte of a character (the window's last position: decided by the deferred look-ahead)
    if (UTF8) fin = ~shift_down64(cont, 0u, L);
    u64 P = 0, ready = 0, done = 0;
    uint32_t old_prev = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        u64 t;
        if (s == 0) {
            t = PV_ANCHORED ? rs : ~0ull;
            const uint32_t before = PV_BEFORE0;
            if (before) t = apply_before64(t, before, as);
        } else {
            t = adv64(P, (UTF8 && cont0) ? 0u : old_prev, L) & nrs;
            if (PV_STEP_OPT(s - 1)) t |= ready;
        }
        ready = t;
        const u64 ck = sel_class64<NCLS>(c, PV_STEP_CLS(s));
        t &= UTF8 ? (ck & ~cont) : ck;
        const uint32_t old = st.last[s];
        u64 Z = t;
        if (PV_STEP_LOOP(s)) Z = spread64(t, ck & nrs, old, L);
        else if (UTF8) {
            // a marker on the lead byte of a multi-byte character moves to the character's last byte, one round per
            // continuation byte of the longest character of the window — only when some marker of this step sits on such a
            // lead byte, or one is still travelling in from the previous window (warp-uniform test: usually neither)
            if (__any_sync(FULL, (t & lead) != 0 || (cont0 && L.is31 && (old & 0x80000000u)))) {
#pragma unroll 1
                for (int r = 0; r < rounds; ++r) Z |= adv64(Z, old, L) & cont;
            }
        }
        st.last[s] = hi32(Z);
        old_prev = old;
        P = UTF8 ? (Z & fin) : Z;
        if (PV_STEP_EXIT(s)) done |= P;
    }
    return done;
}

// phase B of an item that saw a NUL byte: does THIS row hold one?  Such rows are appended to the work list of the exact VM
// kernel and report no hit here.  Called by the whole warp (cold path).
static __device__ __noinline__ uint32_t item_dirty_row(const Args& A, bool in, int o0, int o1, int row, uint32_t hit, uint32_t lane)
{
    bool dirty = false;
    if (in)
        for (int q = o0; q < o1 && !dirty; ++q) dirty = A.chars[q] == 0;
    const unsigned dm = __ballot_sync(FULL, dirty);
    if (dm) {
        unsigned basei = 0;
        if (lane == 0) basei = atomicAdd(A.dirty_count, __popc(dm));
        basei = __shfl_sync(FULL, basei, 0);
        if (dirty) A.dirty_rows[basei + __popc(dm & ((1u << lane) - 1))] = row;
    }
    return dirty ? 0u : hit;
}

template <int NS, int NCLS, int SPEC>
__device__ __forceinline__ void chain_item_body(const ChainDev& cd, const Args& A)
{
    using PL = PlanLit<SPEC>;
    extern __shared__ __align__(128) unsigned char item_smem[];
    LaneCtx L;
    L.lane = lane_id();
    asm volatile("" : "+r"(L.lane));
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    L.m31 = L.lane == 31 ? 1u : 0u;
    asm volatile("" : "+r"(L.m31));
    const uint32_t lane = L.lane;
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(item_smem) + (threadIdx.x >> 5) * (uint32_t)sizeof(WarpSmItem);  // this warp's block
    uint32_t bits0 = wb + ITEM_SM_BITS;
    uint32_t wr0 = wb + (lane & 3u) * 512u + 16u * (((lane >> 2) + 2u * (lane & 3u)) & 7u);  // my cp.async destination (+ 128 k, + stage)
    uint32_t rd0 = wb + 128u * (lane >> 3);                                                    // my read-back base (+ chunk terms, + stage)
    const uint32_t r1 = 512u + 16u * ((lane + 2u) & 7u), r2 = 1024u + 16u * ((lane + 4u) & 7u), r3 = 1536u + 16u * ((lane + 6u) & 7u);
    const uint32_t r0 = 16u * (lane & 7u);
    const char* gsrc = A.chars + 16 * (int)lane;
    asm volatile("" : "+r"(wb), "+r"(bits0), "+r"(wr0), "+r"(rd0));
    const uint32_t bneed = PV_BUILTIN_UNION | ((PV_NEEDS & (AS_BOW | AS_NBOW)) ? (1u << AK_ALNUM) : 0u);
    const bool need_nl = (PV_NEEDS & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) != 0;
    const uint32_t uend = (uint32_t)A.end;
    uint32_t top31 = L.is31 ? 0x80000000u : 0u;  // the window's last position: top bit of lane 31's stream words
    asm volatile("" : "+r"(top31));
    uint32_t cnt = 0;  // rows with a match finalised by this lane

    // Work items are handed out dynamically (one atomicAdd each).  The NEXT item's index, row bounds and byte bounds are
    // fetched while the current item is being processed (the chain atomic -> item_bounds -> offsets is three dependent
    // memory latencies, ~7 % of an item's time when it sits at the top of the item)
    int nxt_item = 0, nxt_ra = 0, nxt_rb = 0, nxt_ba = 0, nxt_bb = 0;
    int timing_item = 0;
    (void)timing_item;
    ITEM_STAMP(0);
#ifdef CUSTR_ITEM_TIMING
    if (L.lane == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_item_times[(blockIdx.x * WARPS + (threadIdx.x >> 5)) * ITEM_TIMING_SLOTS + 2] = smid;
    }
#endif
    if (lane == 0) nxt_item = (int)atomicAdd(A.item_counter, 1u);
    nxt_item = __shfl_sync(FULL, nxt_item, 0);
    if (nxt_item < A.nitems) {
        nxt_ra = __ldg(A.item_bounds + nxt_item);
        nxt_rb = __ldg(A.item_bounds + nxt_item + 1);
        nxt_ba = __ldg(A.offsets + nxt_ra);
        nxt_bb = __ldg(A.offsets + nxt_rb);
    }
    ITEM_STAMP(1);
    for (;;) {
        const int item = nxt_item;
        if (item >= A.nitems) break;
        ITEM_STAMP(4 + 4 * timing_item);
        const int ra = nxt_ra, rb = nxt_rb;
        const int byte_a = nxt_ba, byte_b = nxt_bb;
        int fetched = 0;  // lane 0: index of the item after this one (requested now, looked at after phase 0)
        if (lane == 0) fetched = (int)atomicAdd(A.item_counter, 1u);
        int prefetch_stage = 0;  // 0: index requested, 1: row bounds requested, 2: byte bounds requested
        if (ra >= rb) {  // no rows (cannot happen with the item index of k_item_bounds, kept for safety): just move on
            nxt_item = __shfl_sync(FULL, fetched, 0);
            if (nxt_item < A.nitems) {
                nxt_ra = __ldg(A.item_bounds + nxt_item);
                nxt_rb = __ldg(A.item_bounds + nxt_item + 1);
                nxt_ba = __ldg(A.offsets + nxt_ra);
                nxt_bb = __ldg(A.offsets + nxt_rb);
            }
            continue;
        }
        // (an item that holds only empty rows runs the segment loop once with no window: phase B writes its zeros.  A separate
        // loop + `continue` here makes ptxas give up structured reconvergence for the whole kernel: BRA.DIV guards everywhere)
        const bool empty_item = byte_a >= byte_b;
        ItemState<NS> st;
#pragma unroll
        for (int s = 0; s < NS; ++s) st.last[s] = 0;
        st.last_al = st.last_nl = st.last_f = st.pend = 0;
        u64 zacc = 0;  // NUL bytes seen by this lane in this item
        uint32_t ws = empty_item ? (uint32_t)byte_a : ((uint32_t)byte_a & ~(uint32_t)(WIN64 - 1));
        int wins_left = empty_item ? 0 : (int)((((uint32_t)byte_b - 1u) >> 11) - ((uint32_t)byte_a >> 11)) + 1;
        // the item's slice of the offsets array is read twice (phase 0 and phase B), 32 consecutive entries per lane and step:
        // touch all of its lines now, in parallel, so that those loads are cache hits instead of one DRAM latency per step
        for (int j0 = ra; j0 <= rb; j0 += 1024) {
            const int j = j0 + 32 * (int)lane;
            if (j <= rb) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.offsets + j));
        }
        int krs = ra;   // next offsets index whose ROWSTART bit is not set yet
        int kfin = ra;  // next row to finalise
        uint32_t stage = 0;
        __syncwarp();  // the previous item's reads of the ring and of `bits` are done
        if (!empty_item) ring_issue_item(wr0, gsrc, A.chars, ws, uend, lane);

        do {
            const int nw = wins_left < SEG_WINS ? wins_left : SEG_WINS;
            const uint32_t seg_ws = ws;
            const uint32_t span = (uint32_t)nw * WIN64;
            // ---- phase 0: ROWSTART bits of the segment (and of the position right behind it: extra word)
            for (uint32_t i = 2u * lane; i <= 32u * (uint32_t)nw; i += 64u)
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(bits0 + 8u * i), "r"(0u) : "memory");
            __syncwarp();
            {   // two rows per lane and step; the next step's offsets are in flight while this one is scattered;
                // rel = 0xffffffff marks "no such row" (a real one is < 2^31)
                int j = krs + (int)lane;
                int oa = j <= rb ? __ldg(A.offsets + j) : 0, ob = j + 32 <= rb ? __ldg(A.offsets + j + 32) : 0;
                for (;;) {
                    const int na = j + 64 <= rb ? __ldg(A.offsets + j + 64) : 0, nb = j + 96 <= rb ? __ldg(A.offsets + j + 96) : 0;
                    const uint32_t rel0 = j <= rb ? (uint32_t)oa - seg_ws : 0xffffffffu, rel1 = j + 32 <= rb ? (uint32_t)ob - seg_ws : 0xffffffffu;
                    if (rel0 <= span) reds_or(bits0 + ((rel0 >> 3) & ~3u), 1u << (rel0 & 31));
                    if (rel1 <= span) reds_or(bits0 + ((rel1 >> 3) & ~3u), 1u << (rel1 & 31));
                    const unsigned m0 = __ballot_sync(FULL, rel0 < span), m1 = __ballot_sync(FULL, rel1 < span);
                    krs += __popc(m0) + __popc(m1);
                    if (m1 != FULL) break;
                    j += 64;
                    oa = na;
                    ob = nb;
                }
            }
            __syncwarp();
            if (prefetch_stage == 0) {  // the next item's index has arrived by now: request its row bounds
                nxt_item = __shfl_sync(FULL, fetched, 0);
                if (nxt_item < A.nitems) {
                    nxt_ra = __ldg(A.item_bounds + nxt_item);
                    nxt_rb = __ldg(A.item_bounds + nxt_item + 1);
                }
                prefetch_stage = 1;
            }

            ITEM_STAMP(4 + 4 * timing_item + 1);
            // ---- phase A: the windows of the segment
            for (int w = 0; w < nw; ++w, ws += WIN64, stage ^= 1u) {
                const bool more = (w + 1 < nw) || (wins_left > nw);
                const uint32_t so = stage * WIN64;
                if (more) ring_issue_item(wr0 + (so ^ WIN64), gsrc, A.chars, ws + WIN64, uend, lane);  // next window in flight
                const uint32_t wa = bits0 + 8u * (32u * (uint32_t)w + lane);  // my word of the segment's stream
                const u64 rs = lds64(wa);
                const uint32_t rs_nb = lds32(wa + 8u);                                   // low word of the next stream word
                const bool rs_next = (lds32(bits0 + 256u * (uint32_t)(w + 1)) & 1u) != 0;  // a row starts right behind this window
                if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();  // every lane's copies of this window have landed
                u64 p[8];
                {
                    uint32_t pl[8], ph[8];
                    const uint32_t rb_ = rd0 + so;
                    const uint4 v0 = lds128(rb_ + r0);
                    const uint4 v1 = lds128(rb_ + r1);
                    transpose_planes(v0, v1, pl);
                    const uint4 v2 = lds128(rb_ + r2);
                    const uint4 v3 = lds128(rb_ + r3);
                    transpose_planes(v2, v3, ph);
#pragma unroll
                    for (int b = 0; b < 8; ++b) p[b] = mk64(pl[b], ph[b]);
                }
                const u64 na = p[7];
                const u64 nz5 = p[4] | p[3] | p[2] | p[1] | p[0];
                const u64 letter5 = nz5 & ~(p[4] & p[3] & (p[2] | (p[1] & p[0]))), digit = cls_digit(p);  // low five bits in 1..26
                const u64 alnum = (p[6] & letter5) | digit, word = alnum | cls_underscore(p);
                zacc |= ~(nz5 | p[5] | p[6] | p[7]);
                u64 space = 0;
                if (bneed & (1u << AK_SPACE)) space = cls_space(p);
                NaClasses<NCLS> nc;
                u64 (&c)[NCLS] = nc.c;
                if constexpr (PL::jit) classes_literal64<PL, NCLS>(c, p, letter5, digit, alnum, word, space);
                else
#pragma unroll
                for (int k = 0; k < NCLS; ++k) {
                    u64 v = 0;
                    if (k < (int)PV_NCLASSES) {
                        const uint32_t f = PV_CLS_BUILTINS(k);
                        if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_WORD)) v = word;  // single builtin: inline
                        else if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_DIGIT)) v = digit;
                        else if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_ALNUM)) v = alnum;
                        else if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_SPACE)) v = space;
                        else if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_LOWER)) v = p[6] & p[5] & letter5;
                        else v = class_generic64(cd.classes[k], p, letter5, digit, alnum, word, space);
                        if (PV_CLS_NEGATE(k)) v = ~v;
                    }
                    c[k] = v;
                }
                nc.al = alnum;
                u64 nl = need_nl ? (cls_eq(p, '\n') & ~na) : 0ull;
                const bool utf8 = __any_sync(FULL, na != 0);
                u64 cont = 0;
                uint32_t cont0 = 0;
                int rounds = 0;
                if (utf8) {
                    if (na) nc = classify_non_ascii64<NCLS, 1>(cd, A, rd0 + so, (int)ws + 64 * (int)lane, na, nc);
                    cont = p[7] & ~p[6];
                    const u64 lead3 = p[7] & p[6] & p[5];  // lead byte of a 3- or 4-byte character
                    rounds = 1 + (int)__any_sync(FULL, lead3 != 0) + (int)__any_sync(FULL, (lead3 & p[4]) != 0);
                    cont0 = __shfl_sync(FULL, lo32(cont), 0) & 1u;
                }
                const u64 al = nc.al;
                const u64 nrs = ~rs;

                // ---- the match end withheld at the previous window's last position: its look-ahead is this window's first byte
                {
                    uint32_t ok = st.pend;
                    const uint32_t em = PV_END_MASK;
                    if (em & (AS_BOW | AS_NBOW)) {
                        const uint32_t differ = st.last_al ^ (__shfl_sync(FULL, lo32(al), 0) << 31);  // alnum before / behind the boundary
                        if (em & AS_BOW) ok &= differ;
                        if (em & AS_NBOW) ok &= ~differ;
                    }
                    if (em & AS_EOL_DOLLAR) ok &= __shfl_sync(FULL, lo32(nl), 0) << 31;  // '$' in the middle of a row: a newline follows
                    if (em & AS_EOL_Z) ok = 0;
                    if (utf8 && cont0) ok = 0;  // the window boundary lies inside a character
                    st.last_f |= ok & 0x80000000u;
                }

                // ---- zero-width assertion streams (look-ahead behind the window's last position: "end of row", see below)
                Assertions64 as;
                as.rs = rs;
                as.nl = nl;
                as.bow_b = as.bow_a = as.bolc_b = as.lb = as.eold_a = 0;
                if (PV_NEEDS & (AS_BOW | AS_NBOW)) {
                    as.bow_b = al ^ (adv64(al, st.last_al, L) & nrs);
                    as.bow_a = al ^ shift_down64(al & nrs, 0u, L);
                    st.last_al = hi32(al);
                }
                if (PV_NEEDS & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) {
                    as.bolc_b = rs | (adv64(nl, st.last_nl, L) & nrs);
                    as.lb = shift_down64_nb(rs, rs_nb);
                    as.eold_a = as.lb | shift_down64(nl & nrs, 0u, L);
                    st.last_nl = hi32(nl);
                }
                u64 done;
                if (!utf8) done = chain_item64<NS, NCLS, SPEC, false>(cd, c, as, rs, 0, 0ull, 0ull, 0u, st, L);
                else done = chain_item64<NS, NCLS, SPEC, true>(cd, c, as, rs, rounds, cont, p[7] & p[6], cont0, st, L);
                u64 E = PV_END_MASK ? apply_after64(done, PV_END_MASK, as) : done;
                // the last position of the window: with a row start right behind it the look-ahead used above ("nothing
                // follows") is exact; otherwise the bit is withheld and decided by the next window
                {
                    u64 d2 = done;  // END assertions that need no look-ahead
                    if (PV_END_MASK & AS_BOL_CARET) d2 &= nl;
                    if (PV_END_MASK & AS_BOL_A) d2 = 0;
                    const uint32_t hold = rs_next ? 0u : top31;
                    st.pend = hi32(d2) & hold;
                    E &= ~((u64)hold << 32);
                }
                // ---- sticky per-row OR of the match ends; it replaces the window's ROWSTART word
                const u64 F = spread64(E, nrs, st.last_f, L);
                st.last_f = hi32(F);
                sts64(wa, lo32(F), hi32(F));
                __syncwarp();  // ring stage and stream words are free for the next iteration / phase B
            }

            ITEM_STAMP(4 + 4 * timing_item + 2);
            // ---- phase B: the rows that end inside the segment
            if (prefetch_stage == 1) {  // ... and now its byte bounds (they are back when phase B is through)
                if (nxt_item < A.nitems) {
                    nxt_ba = __ldg(A.offsets + nxt_ra);
                    nxt_bb = __ldg(A.offsets + nxt_rb);
                }
                prefetch_stage = 2;
            }
            const bool item_dirty = __any_sync(FULL, zacc != 0);
            {   // two rows per lane and step, the next step's offsets in flight
                int j = kfin + (int)lane;
                int oa0 = 0, oa1 = 0, ob0 = 0, ob1 = 0;
                if (j < rb) { oa0 = __ldg(A.offsets + j); oa1 = __ldg(A.offsets + j + 1); }
                if (j + 32 < rb) { ob0 = __ldg(A.offsets + j + 32); ob1 = __ldg(A.offsets + j + 33); }
                for (;;) {
                    int na0 = 0, na1 = 0, nb0 = 0, nb1 = 0;
                    if (j + 64 < rb) { na0 = __ldg(A.offsets + j + 64); na1 = __ldg(A.offsets + j + 65); }
                    if (j + 96 < rb) { nb0 = __ldg(A.offsets + j + 96); nb1 = __ldg(A.offsets + j + 97); }
                    // (end of the row) - (segment start); 0xffffffff: no such row
                    const uint32_t rela = j < rb ? (uint32_t)oa1 - seg_ws : 0xffffffffu, relb = j + 32 < rb ? (uint32_t)ob1 - seg_ws : 0xffffffffu;
                    const bool ina = rela <= span, inb = relb <= span;
                    uint32_t hita = 0, hitb = 0;
                    if (ina && oa1 > oa0) hita = (lds32(bits0 + (((rela - 1u) >> 3) & ~3u)) >> ((rela - 1u) & 31)) & 1u;
                    if (inb && ob1 > ob0) hitb = (lds32(bits0 + (((relb - 1u) >> 3) & ~3u)) >> ((relb - 1u) & 31)) & 1u;
                    if (__builtin_expect(item_dirty, 0)) {  // a NUL byte somewhere in this item: rows holding one go to the exact VM
                        hita = item_dirty_row(A, ina, oa0, oa1, j, hita, lane);
                        hitb = item_dirty_row(A, inb, ob0, ob1, j + 32, hitb, lane);
                    }
                    if (ina) A.out[j] = (uint8_t)hita;
                    if (inb) A.out[j + 32] = (uint8_t)hitb;
                    cnt += hita + hitb;
                    const unsigned ma = __ballot_sync(FULL, ina), mb = __ballot_sync(FULL, inb);
                    kfin += __popc(ma) + __popc(mb);
                    if (mb != FULL) break;
                    j += 64;
                    oa0 = na0; oa1 = na1; ob0 = nb0; ob1 = nb1;
                }
            }
            ITEM_STAMP(4 + 4 * timing_item + 3);
            wins_left -= nw;
            __syncwarp();
        } while (wins_left > 0);
        ++timing_item;
    }
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0 && cnt) atomicAdd(A.total, (unsigned long long)cnt);
}

template <int NS, int NCLS, int SPEC = 0>
__global__ void __launch_bounds__(THREADS, ITEM_MIN_CTAS)
k_chain_item(const __grid_constant__ ChainDev cd, const __grid_constant__ Args A)
{
    chain_item_body<NS, NCLS, SPEC>(cd, A);
}
#ifdef CUSTR_JIT  // the one kernel of a run-time compiled module (regex_jit.cpp): the plan's own view, PlanLit<100>
extern "C" __global__ void __launch_bounds__(THREADS, 3)
custr_jit_chain_item(const __grid_constant__ ChainDev cd, const __grid_constant__ Args A)
{
    chain_item_body<CUSTR_JIT_NS, CUSTR_JIT_NCLS, 100>(cd, A);
}
#endif

#if !defined(ITEM_EXPERIMENT) && !defined(CUSTR_JIT) && !defined(CUSTR_NO_ITEM_LAUNCHERS)
// function attributes are per (function, device): set on the first launch there, not on every call
static bool item_attrs_needed(const void* fn)
{
    static std::mutex mu;
    static std::unordered_map<const void*, unsigned long long> done;  // bit = device ordinal
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long& m = done[fn];
    if (m & bit) return false;
    m |= bit;
    return true;
}
template <int NS>
static void launch_item_ns(const ChainDev& cd, const Args& a, int blocks)
{
#define ITEM_LAUNCH(K)                                                                                                   \
    do {                                                                                                                 \
        auto kfn = K;                                                                                                    \
        if (item_attrs_needed((const void*)kfn)) {                                                                       \
            CUSTR_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ITEM_SMEM_BYTES));         \
            CUSTR_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        }                                                                                                                \
        LAUNCH(kfn, blocks, THREADS, ITEM_SMEM_BYTES, cd, a);                                                            \
    } while (0)
    bool plain = true;  // no optional step, END only behind the last step
    for (uint32_t s = 0; s < cd.nsteps; ++s) plain = plain && !cd.steps[s].opt && ((cd.steps[s].exit != 0) == (s + 1 == cd.nsteps));
    if (!plain) {
        if (cd.nclasses <= 1) ITEM_LAUNCH((k_chain_item<NS, 1, 5>));
        else if (cd.nclasses == 2) ITEM_LAUNCH((k_chain_item<NS, 2, 5>));
        else if (cd.nclasses <= 4) ITEM_LAUNCH((k_chain_item<NS, 4, 5>));
        else ITEM_LAUNCH((k_chain_item<NS, 8, 5>));
        return;
    }
    if constexpr (NS <= 4) {
        const int spec = g_no_spec ? 0 : chain_spec_of(cd);
        if (spec == 1) { ITEM_LAUNCH((k_chain_item<NS, 1, 1>)); return; }
        if (spec == 2) { ITEM_LAUNCH((k_chain_item<NS, 1, 2>)); return; }
        if (spec == 3) { ITEM_LAUNCH((k_chain_item<NS, 1, 3>)); return; }
        if (spec == 4) { ITEM_LAUNCH((k_chain_item<NS, 1, 4>)); return; }
        if (spec == 6) { ITEM_LAUNCH((k_chain_item<NS, 1, 6>)); return; }
        if (spec == 7) { ITEM_LAUNCH((k_chain_item<NS, 1, 7>)); return; }
    }
    if (cd.nclasses <= 1) ITEM_LAUNCH((k_chain_item<NS, 1, 0>));
    else if (cd.nclasses == 2) ITEM_LAUNCH((k_chain_item<NS, 2, 0>));
    else if (cd.nclasses <= 4) ITEM_LAUNCH((k_chain_item<NS, 4, 0>));
    else ITEM_LAUNCH((k_chain_item<NS, 8, 0>));
#undef ITEM_LAUNCH
}
#elif !defined(CUSTR_JIT) && !defined(CUSTR_NO_ITEM_LAUNCHERS)
template <int NS>
static void launch_item_ns(const ChainDev&, const Args&, int) {}
#endif
