// Linear-chain bitstream kernel — included by regex_bits.cu inside namespace custr::bits.
//
// The specialised, UTF-8 aware executor for the common pattern shape (class sequences with x+ / x* tails and leading /
// trailing assertions, e.g. the headline \b\w{4,}\b).  Differences to the generic k_bitstream interpreter:
//   * the chain is unrolled at compile time (NS steps, NCLS classes): streams live in registers, no shared-memory slots
//   * the carry between windows is the previous window's stream register; an advance is SEL + SHFL + SHF
//   * builtin classes (\w \d \s alnum a-z A-Z) are computed once per window and shared by classes and by \b
//   * row bookkeeping touches the offsets array once per window (ROWSTART scatter and row finalisation share the load)
//   * multi-byte characters: a marker sits on the LAST byte of the character it consumed; class bits of non-ASCII bytes
//     come from decoding each character once (rare, warp-divergent path); a spread through continuation bytes moves a
//     marker from the lead byte to the last byte.  Only rows holding a NUL byte go to the exact Pike-VM kernel.
#pragma once

struct LaneCtx {
    uint32_t lane, src;  // src = lane-1 (mod 32): source lane of an advance
    bool is31;
};

__device__ __forceinline__ uint32_t adv_rot(uint32_t x, uint32_t last_x, const LaneCtx& L)
{
    uint32_t v = L.is31 ? last_x : x;
    return __funnelshift_l(__shfl_sync(FULL, v, L.src), x, 1);
}
// R[p] = Q[p] | (R[p-1] & K[p]); the carry R[-1] is the top bit of the previous window's result register
__device__ __forceinline__ uint32_t spread_rot(uint32_t q, uint32_t k, uint32_t last_r, const LaneCtx& L)
{
    uint32_t s = adv_rot(q, last_r, L) & k;
    uint32_t sum = s + k;
    uint32_t g = __ballot_sync(FULL, sum < s);
    uint32_t p = __ballot_sync(FULL, sum == 0xffffffffu);
    uint32_t g1 = g << 1, p1 = p << 1;
    uint32_t s2 = (g1 << 1) & p1;
    uint32_t c = g1 | ((((s2 + p1) ^ p1) | s2) & p1);
    sum += (c >> L.lane) & 1u;
    return q | (((sum ^ k) | s) & k);
}

template <int NCLS>
__device__ __forceinline__ uint32_t sel_class(const uint32_t (&c)[NCLS], uint32_t k)
{
    if (NCLS == 1) return c[0];
    if (NCLS == 2) return k ? c[1] : c[0];
    uint32_t v = c[0];
#pragma unroll
    for (int i = 1; i < NCLS; ++i)
        if (k == (uint32_t)i) v = c[i];
    return v;
}

// exact class test of a NON-ASCII character (reference regexec.inl:127-155) from the inlined class definition
__device__ __forceinline__ bool na_class_inline(const ChainClassD& cd, const uint8_t* __restrict__ uflags, uint32_t ch)
{
    for (uint32_t i = 0; i < cd.na_nranges; i += 2)
        if (ch >= cd.na_ranges[i] && ch <= cd.na_ranges[i + 1]) return true;
    const uint32_t b = cd.na_builtins;
    if (!b) return false;
    const uint32_t cp = packed_to_cp(ch);
    if (cp > 0xFFFFu) return false;
    const uint32_t f = __ldg(uflags + cp);
    const bool alnum = (f & 15u) != 0, space = (f & 16u) != 0, digit = (f & 4u) != 0;
    return ((b & rx::CB_W) && alnum) || ((b & rx::CB_S) && space) || ((b & rx::CB_D) && digit) || ((b & rx::CB_NW) && !alnum) ||
           ((b & rx::CB_NS) && !space) || ((b & rx::CB_ND) && !digit);
}
__device__ __forceinline__ bool na_char_matches(const ChainClassD& cd, const Args& A, uint32_t ch)
{
    switch (cd.na_kind) {
    case NA_ALWAYS: return true;
    case NA_CHAR_EQ: return ch == cd.na_arg;
    case NA_CLASS:
        return cd.na_inline ? na_class_inline(cd, A.uflags, ch)
                            : rxdev::class_match(rxdev::bind_program(A.prog_img, A.uflags), (int)cd.na_arg, ch);
    case NA_NCLASS:
        return !(cd.na_inline ? na_class_inline(cd, A.uflags, ch)
                              : rxdev::class_match(rxdev::bind_program(A.prog_img, A.uflags), (int)cd.na_arg, ch));
    default: return false;
    }
}

// Non-ASCII bytes of this lane: decode each character once and give ALL its bytes (inside the lane) its class bits.
template <int NCLS>
__device__ __noinline__ void classify_non_ascii(const ChainDev& cd, const Args& A, int lane_base, uint32_t na, uint32_t (&c)[NCLS],
                                                uint32_t& al)
{
    const uint8_t* base = (const uint8_t*)A.chars;
    while (na) {
        int b = __ffs(na) - 1;
        int q = lane_base + b;
        while (q > A.first && (base[q] & 0xC0) == 0x80) --q;  // only the leading continuation run has to walk back
        int w;
        uint32_t ch = utf8_packed(base + q, base + A.end, w);
        int lo = q - lane_base, hi = lo + w;  // bytes of the character, lane-relative
        if (lo < 0) lo = 0;
        if (hi > 32) hi = 32;
        if (hi <= b) hi = b + 1;  // malformed input: always make progress
        uint32_t bits = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        na &= ~bits;
#pragma unroll
        for (int k = 0; k < NCLS; ++k)
            if (k < (int)cd.nclasses) c[k] = na_char_matches(cd.classes[k], A, ch) ? (c[k] | bits) : (c[k] & ~bits);
        al = is_alnum_packed(ch, A.uflags) ? (al | bits) : (al & ~bits);
    }
}

template <int NS>
struct ChainState {  // per-warp streaming state carried between windows (values of the previous window)
    uint32_t last[NS];
    uint32_t last_al, last_nl, last_f, last_d;
};

template <int NS, int NCLS, bool UTF8>
__device__ __forceinline__ uint32_t chain_eval(const ChainDev& cd, const uint32_t (&c)[NCLS], uint32_t al, uint32_t nl, uint32_t rs,
                                               uint32_t cont, uint32_t rs_next, uint32_t next_is_cont, uint32_t a_next,
                                               uint32_t nl_next, ChainState<NS>& st, const LaneCtx& L)
{
    const uint32_t nrs = ~rs;
    const uint32_t fin = UTF8 ? ~shift_down(cont, next_is_cont) : 0xffffffffu;  // last byte of a character
    const uint32_t cont0 = UTF8 ? (__shfl_sync(FULL, cont, 0) & 1u) : 0u;        // window starts inside a character
    Assertions as;
    as.rs = rs;
    as.nl = nl;
    as.bow_b = as.bow_a = as.bolc_b = as.lb = as.eold_a = 0;
    if (cd.needs & (AS_BOW | AS_NBOW)) {
        as.bow_b = al ^ (adv_rot(al, st.last_al, L) & nrs);
        as.bow_a = al ^ shift_down(al & nrs, a_next);
        st.last_al = al;
    }
    if (cd.needs & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) {
        as.bolc_b = rs | (adv_rot(nl, st.last_nl, L) & nrs);
        as.lb = shift_down(rs, rs_next);
        as.eold_a = as.lb | shift_down(nl & nrs, nl_next);
        st.last_nl = nl;
    }
    uint32_t P = 0, old_prev = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        uint32_t t;
        if (s == 0) {
            t = cd.anchored ? rs : 0xffffffffu;
            const uint32_t before = cd.steps[0].before;
            if (before) t = apply_before(t, before, as);
        }
        else {
            // the marker of step s-1 moves to the next position; the carry (previous window's stream) is dropped when
            // this window starts inside a character: that marker was not on a final byte
            uint32_t v = L.is31 ? (cont0 ? 0u : old_prev) : P;
            t = __funnelshift_l(__shfl_sync(FULL, v, L.src), P, 1) & nrs;
        }
        const uint32_t ck = sel_class<NCLS>(c, cd.steps[s].cls);
        t &= UTF8 ? (ck & ~cont) : ck;
        const uint32_t old = st.last[s];
        uint32_t Z;
        if (cd.steps[s].loop) Z = spread_rot(t, ck & nrs, old, L);
        else if (UTF8) {  // move the marker from the lead byte to the last byte of its character (<= 3 continuation bytes)
            Z = t | (adv_rot(t, old, L) & cont);
            Z |= adv_rot(Z, old, L) & cont;
            Z |= adv_rot(Z, old, L) & cont;
        } else
            Z = t;
        st.last[s] = Z;
        old_prev = old;
        P = UTF8 ? (Z & fin) : Z;
    }
    return cd.end_mask ? apply_after(P, cd.end_mask, as) : P;
}

template <int NS, int NCLS>
__global__ void __launch_bounds__(THREADS, 4)
k_chain(const __grid_constant__ ChainDev cd, const Args A)
{
    __shared__ uint32_t sm_rs[WARPS][32], sm_f[WARPS][32], sm_d[WARPS][32];
    const int warp = threadIdx.x >> 5;
    uint32_t* S_rs = sm_rs[warp];
    uint32_t* S_f = sm_f[warp];
    uint32_t* S_d = sm_d[warp];
    LaneCtx L;
    L.lane = lane_id();
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    const uint32_t lane = L.lane;
    const int warps_total = gridDim.x * WARPS;
    unsigned long long my_matches = 0;
    const uint32_t bneed = cd.builtin_union | ((cd.needs & (AS_BOW | AS_NBOW)) ? (1u << AK_ALNUM) : 0u);
    const bool need_nl = (cd.needs & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) != 0;

    for (int item = blockIdx.x * WARPS + warp; item < A.nitems; item += warps_total) {
        const int lo_byte = A.first + item * ITEM_BYTES;
        const int ra = item == 0 ? 0 : warp_lower_bound(A.offsets, A.n, lo_byte);
        const int rb = item == A.nitems - 1 ? A.n : warp_lower_bound(A.offsets, A.n, lo_byte + ITEM_BYTES);
        if (ra >= rb) continue;
        const int byte_a = __ldg(A.offsets + ra), byte_b = __ldg(A.offsets + rb);
        if (byte_a >= byte_b) continue;  // only empty rows: results stay 0 (pre-cleared)
        ChainState<NS> st;
#pragma unroll
        for (int s = 0; s < NS; ++s) st.last[s] = 0;
        st.last_al = st.last_nl = st.last_f = st.last_d = 0;
        uint32_t d_live = 0;
        int ws = byte_a & ~(WIN - 1);
        int kcur = ra + 1;       // next offsets index to consume; offsets[j] > ws for every j >= kcur
        int prev_o = byte_a;     // offsets[kcur - 1]
        int pend = byte_a - ws;  // window-relative position of a row start already known (-1: none)
        int o_nxt = (kcur + (int)lane <= rb) ? __ldg(A.offsets + kcur + (int)lane) : 0x7fffffff;  // offsets are prefetched too
        uint4 cur_lo, cur_hi, nxt_lo, nxt_hi;
        load_window(A.chars, ws, A.end, cur_lo, cur_hi);

        for (; ws < byte_b; ws += WIN) {
            const int we = ws + WIN;
            if (we < byte_b) load_window(A.chars, we, A.end, nxt_lo, nxt_hi);  // prefetch the next window

            // ---- one pass over the offsets that fall into (ws, we]: ROWSTART bits now, row results after evaluation
            S_rs[lane] = (lane == 0 && pend == 0) ? 1u : 0u;
            __syncwarp();
            if (pend > 0) {  // first window of the item: its first row starts inside the window
                if (lane == 0) atomicOr(&S_rs[pend >> 5], 1u << (pend & 31));
            }
            const int j = kcur + (int)lane;
            const int o = o_nxt;
            const bool inw = o <= we;
            if (inw && o < we) atomicOr(&S_rs[(o - ws) >> 5], 1u << ((o - ws) & 31));
            const unsigned m_in = __ballot_sync(FULL, inw);
            bool at_we = __any_sync(FULL, inw && o == we);
            int consumed = __popc(m_in);
            if (m_in == FULL) {  // more than 32 rows end in this window (short / empty rows): generic loop
                for (;;) {
                    int j2 = kcur + consumed + (int)lane;
                    int o2 = j2 <= rb ? __ldg(A.offsets + j2) : 0x7fffffff;
                    bool in2 = o2 <= we;
                    if (in2 && o2 < we) atomicOr(&S_rs[(o2 - ws) >> 5], 1u << ((o2 - ws) & 31));
                    unsigned m2 = __ballot_sync(FULL, in2);
                    at_we = at_we || __any_sync(FULL, in2 && o2 == we);
                    consumed += __popc(m2);
                    if (m2 != FULL) break;
                }
            }
            {   // first offsets chunk of the NEXT window: issued now, consumed one iteration later
                const int jn = kcur + consumed + (int)lane;
                o_nxt = jn <= rb ? __ldg(A.offsets + jn) : 0x7fffffff;
            }
            __syncwarp();
            const uint32_t rs = S_rs[lane];
            const uint32_t rs_next = at_we || we >= A.end;
            const uint32_t next_byte = (!rs_next && we < A.end) ? (uint8_t)A.chars[we] : 0;

            // ---- bit planes, builtin class streams, classes
            uint32_t p[8];
            transpose_planes(cur_lo, cur_hi, p);
            const uint32_t na = p[7];
            const uint32_t zero = ~(p[0] | p[1] | p[2] | p[3] | p[4] | p[5] | p[6] | p[7]);
            uint32_t letter5 = 0, digit = 0, alnum = 0, word = 0, space = 0;
            if (bneed & ((1u << AK_WORD) | (1u << AK_ALNUM) | (1u << AK_LOWER) | (1u << AK_UPPER))) letter5 = cls_letter5(p);
            if (bneed & ((1u << AK_WORD) | (1u << AK_ALNUM) | (1u << AK_DIGIT))) digit = cls_digit(p);
            if (bneed & ((1u << AK_WORD) | (1u << AK_ALNUM))) alnum = (p[6] & letter5) | digit;
            if (bneed & (1u << AK_WORD)) word = alnum | cls_underscore(p);
            if (bneed & (1u << AK_SPACE)) space = cls_space(p);
            uint32_t c[NCLS];
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
                uint32_t v = 0;
                if (k < (int)cd.nclasses) {
                    const uint32_t f = cd.classes[k].builtins;
                    switch (f) {  // single builtin (the common case) without walking the flag list
                    case 0: break;
                    case 1u << AK_WORD: v = word; break;
                    case 1u << AK_ALNUM: v = alnum; break;
                    case 1u << AK_DIGIT: v = digit; break;
                    case 1u << AK_SPACE: v = space; break;
                    case 1u << AK_ANY: v = 0xffffffffu; break;
                    default:
                        if (f & (1u << AK_WORD)) v |= word;
                        if (f & (1u << AK_ALNUM)) v |= alnum;
                        if (f & (1u << AK_DIGIT)) v |= digit;
                        if (f & (1u << AK_SPACE)) v |= space;
                        if (f & (1u << AK_LOWER)) v |= p[6] & p[5] & letter5;
                        if (f & (1u << AK_UPPER)) v |= p[6] & ~p[5] & letter5;
                        if (f & (1u << AK_ANY)) v = 0xffffffffu;
                        break;
                    }
                    for (uint32_t a = 0; a < cd.classes[k].natoms; ++a) v |= cls_atom(p, cd.classes[k].atoms[a]);
                    if (cd.classes[k].negate) v = ~v;
                }
                c[k] = v;
            }
            uint32_t al = alnum;
            uint32_t nl = need_nl ? (cls_eq(p, '\n') & ~na) : 0u;
            uint32_t a_next = 0;
            const uint32_t nl_next = next_byte == '\n';
            if (cd.needs & (AS_BOW | AS_NBOW)) {
                if (next_byte < 0x80u) a_next = (next_byte - '0' < 10u) || ((next_byte | 0x20u) - 'a' < 26u);
                else if ((next_byte & 0xC0u) != 0x80u) {
                    int w;
                    a_next = is_alnum_packed(utf8_packed((const uint8_t*)A.chars + we, (const uint8_t*)A.chars + A.end, w), A.uflags);
                }
            }
            uint32_t E;
            if (__any_sync(FULL, na != 0)) {
                classify_non_ascii<NCLS>(cd, A, ws + 32 * (int)lane, na, c, al);
                const uint32_t cont = p[7] & ~p[6];
                E = chain_eval<NS, NCLS, true>(cd, c, al, nl, rs, cont, rs_next, (next_byte & 0xC0u) == 0x80u, a_next, nl_next, st, L);
            } else
                E = chain_eval<NS, NCLS, false>(cd, c, al, nl, rs, 0u, rs_next, 0u, a_next, nl_next, st, L);

            // ---- sticky per-row OR of the match bits; NUL bytes make a row "dirty" (decided by the exact VM)
            const uint32_t nrs = ~rs;
            const uint32_t F = spread_rot(E, nrs, st.last_f, L);
            st.last_f = F;
            if (m_in == FULL) S_f[lane] = F;  // only the rare multi-chunk path reads it back from shared memory
            const bool any_dirty = __any_sync(FULL, zero != 0) || d_live;
            if (any_dirty) {
                const uint32_t D = spread_rot(zero, nrs, st.last_d, L);
                st.last_d = D;
                S_d[lane] = D;
                d_live = __shfl_sync(FULL, D, 31) >> 31;
            }
            __syncwarp();

            // ---- finalise the rows whose last byte lies in this window (first chunk from registers)
            {
                int o_prev = __shfl_up_sync(FULL, o, 1);
                if (lane == 0) o_prev = prev_o;
                bool hit = false, dirty = false;
                const int b = (o - 1 - ws) & (WIN - 1);
                const uint32_t fw = __shfl_sync(FULL, F, b >> 5);  // the word of F holding this row's last byte
                if (inw && o > o_prev) {  // non-empty row j-1, last byte o-1 >= ws
                    hit = (fw >> (b & 31)) & 1u;
                    dirty = any_dirty && ((S_d[b >> 5] >> (b & 31)) & 1u);
                    if (!dirty) A.out[j - 1] = hit;
                }
                if (any_dirty) {
                    const unsigned dm = __ballot_sync(FULL, dirty);
                    if (dm) {
                        unsigned basei = 0;
                        if (lane == 0) basei = atomicAdd(A.dirty_count, __popc(dm));
                        basei = __shfl_sync(FULL, basei, 0);
                        if (dirty) A.dirty_rows[basei + __popc(dm & ((1u << lane) - 1))] = j - 1;
                    }
                }
                my_matches += __popc(__ballot_sync(FULL, hit && !dirty));
            }
            if (m_in == FULL) {  // remaining chunks: reload
                for (int k2 = kcur + 32; k2 < kcur + consumed; k2 += 32) {
                    const int j2 = k2 + (int)lane;
                    bool hit = false, dirty = false;
                    if (j2 < kcur + consumed) {
                        const int o2 = __ldg(A.offsets + j2), o2p = __ldg(A.offsets + j2 - 1);
                        if (o2 > o2p) {
                            const int b = o2 - 1 - ws;
                            hit = (S_f[b >> 5] >> (b & 31)) & 1u;
                            dirty = any_dirty && ((S_d[b >> 5] >> (b & 31)) & 1u);
                            if (!dirty) A.out[j2 - 1] = hit;
                        }
                    }
                    const unsigned dm = __ballot_sync(FULL, dirty);
                    if (dm) {
                        unsigned basei = 0;
                        if (lane == 0) basei = atomicAdd(A.dirty_count, __popc(dm));
                        basei = __shfl_sync(FULL, basei, 0);
                        if (dirty) A.dirty_rows[basei + __popc(dm & ((1u << lane) - 1))] = j2 - 1;
                    }
                    my_matches += __popc(__ballot_sync(FULL, hit && !dirty));
                }
            }
            if (consumed) {
                prev_o = m_in == FULL ? __ldg(A.offsets + kcur + consumed - 1) : __shfl_sync(FULL, o, consumed - 1);
                kcur += consumed;
            }
            pend = at_we ? 0 : -1;
            __syncwarp();
            cur_lo = nxt_lo;
            cur_hi = nxt_hi;
        }
    }
    if (lane == 0 && my_matches) atomicAdd(A.total, my_matches);
}

#ifndef CUSTR_EXPERIMENT_ONLY_4_1
template <int NS>
static void launch_chain_ns(const ChainDev& cd, const Args& a, int blocks)
{
    auto k1 = k_chain<NS, 1>;
    auto k2 = k_chain<NS, 2>;
    auto k4 = k_chain<NS, 4>;
    if (cd.nclasses <= 1) LAUNCH(k1, blocks, THREADS, 0, cd, a);
    else if (cd.nclasses == 2) LAUNCH(k2, blocks, THREADS, 0, cd, a);
    else LAUNCH(k4, blocks, THREADS, 0, cd, a);
}

static void launch_chain(const ChainDev& cd, const Args& a, int blocks)
{
    switch (cd.nsteps) {
    case 1: launch_chain_ns<1>(cd, a, blocks); break;
    case 2: launch_chain_ns<2>(cd, a, blocks); break;
    case 3: launch_chain_ns<3>(cd, a, blocks); break;
    case 4: launch_chain_ns<4>(cd, a, blocks); break;
    case 5: launch_chain_ns<5>(cd, a, blocks); break;
    case 6: launch_chain_ns<6>(cd, a, blocks); break;
    case 7: launch_chain_ns<7>(cd, a, blocks); break;
    default: launch_chain_ns<8>(cd, a, blocks); break;
    }
}
#endif
