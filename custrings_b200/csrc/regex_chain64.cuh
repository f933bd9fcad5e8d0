// Linear-chain bitstream kernel, 64-bit streams — included by regex_bits.cu inside namespace custr::bits.
//
// Same model as regex_chain.cuh (k_chain), twice the work per warp iteration: a WINDOW is 2048 bytes, lane L owns the 64
// bytes [64L, 64L+64) = one 64-bit word of every stream.  Everything that costs a fixed number of instructions per window
// — the offsets pass (ROWSTART scatter + row finalisation), shuffles of advance / look-ahead, the two ballots and the
// scalar carry solve of each `spread`, flag tests, loop overhead — is paid once per 2048 bytes instead of once per 1024.
// Char tiles are staged global -> shared with cp.async (16-byte LDGSTS, zero-filled past the end of the buffer) into a
// two-stage per-warp ring: the copy of window k+1 is in flight while window k is evaluated, and it costs no registers.
// Each lane reads back only the 64 bytes it copied itself (bank-conflict-free swizzle), so cp.async.wait_group is the only
// synchronisation the ring needs.
#pragma once

// Plan view: how the kernel reads the flags of the chain it executes.  The generic kernels (SPEC = 0) read them from the
// parameter block: one binary interprets every chain.  For the most common shapes — a run of ONE builtin class (\w or
// \d; bare also \s and [a-z]) of a minimum length (x+, x{n,}), bare or word-bounded (\b...\b), searched anywhere —
// ahead-of-time specialisations (SPEC 1..4, 6, 7) see the assertion, class and loop flags as literals, so ptxas folds the flag tests, selects and constant-bank reads away (measured on C2:
// 0.445 -> 0.42 ms with every flag literal; the assertion flags alone are worth 5 %, the class flags 4 %).
template <int SPEC>
struct PlanLit {  // SPEC 0: no flag is literal, except that the chain has no optional step and no early exit
    static constexpr bool on = false, opt = false;
    static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 0;
};
template <> struct PlanLit<5> {  // SPEC 5: like 0, for chains WITH optional steps / early exits (x?, x*, x{n,m})
    static constexpr bool on = false, opt = true;
    static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 0;
};
template <> struct PlanLit<1> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 1u << AK_WORD; };
template <> struct PlanLit<2> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 1u << AK_DIGIT; };
template <> struct PlanLit<3> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = AS_BOW, end_mask = AS_BOW, before0 = AS_BOW, builtins = 1u << AK_WORD; };
template <> struct PlanLit<4> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = AS_BOW, end_mask = AS_BOW, before0 = AS_BOW, builtins = 1u << AK_DIGIT; };
template <> struct PlanLit<6> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 1u << AK_SPACE; };
template <> struct PlanLit<7> { static constexpr bool on = true, opt = false; static constexpr uint32_t needs = 0, end_mask = 0, before0 = 0, builtins = 1u << AK_LOWER; };
constexpr int CHAIN_SPECS = 6;
// which specialisation (0 = none) covers this chain
inline int chain_spec_of(const ChainDev& cd)
{
    if (cd.nclasses != 1 || cd.anchored || cd.classes[0].natoms != 0 || cd.classes[0].negate) return 0;
    for (uint32_t s = 0; s < cd.nsteps; ++s)  // the specialisations also fix the structure: no optional step, only the last one loops
        if (cd.steps[s].opt || (cd.steps[s].exit != 0) != (s + 1 == cd.nsteps) || (cd.steps[s].loop != 0) != (s + 1 == cd.nsteps)) return 0;
    const uint32_t b = cd.classes[0].builtins;
    const bool bare = cd.needs == 0 && cd.end_mask == 0 && cd.steps[0].before == 0;
    const bool word_bounded = cd.needs == AS_BOW && cd.end_mask == AS_BOW && cd.steps[0].before == AS_BOW;
    if (b == (1u << AK_SPACE)) return bare ? 6 : 0;   // \s+
    if (b == (1u << AK_LOWER)) return bare ? 7 : 0;   // [a-z]+
    const int kind = b == (1u << AK_WORD) ? 1 : (b == (1u << AK_DIGIT) ? 2 : 0);
    if (!kind) return 0;
    if (bare) return kind;
    if (word_bounded) return 2 + kind;
    return 0;
}
#define PV_NEEDS (PL::on ? PL::needs : cd.needs)
#define PV_ANCHORED (PL::on ? 0u : cd.anchored)
#define PV_END_MASK (PL::on ? PL::end_mask : cd.end_mask)
#define PV_BEFORE0 (PL::on ? PL::before0 : cd.steps[0].before)
#define PV_STEP_CLS(s) (PL::on ? 0u : cd.steps[s].cls)
#define PV_STEP_LOOP(s) (PL::on ? (uint32_t)((s) == NS - 1) : cd.steps[s].loop)
#define PV_STEP_OPT(s) (PL::opt ? cd.steps[s].opt : 0u)
#define PV_STEP_EXIT(s) (PL::opt ? cd.steps[s].exit : (uint32_t)((s) == NS - 1))
#define PV_NCLASSES (PL::on ? 1u : cd.nclasses)
#define PV_CLS_BUILTINS(k) (PL::on ? PL::builtins : cd.classes[k].builtins)
#define PV_CLS_NATOMS(k) (PL::on ? 0u : cd.classes[k].natoms)
#define PV_CLS_NEGATE(k) (PL::on ? 0u : cd.classes[k].negate)
#define PV_BUILTIN_UNION (PL::on ? PL::builtins : cd.builtin_union)

constexpr int WIN64 = 2048;
constexpr int RING_STAGES = 2;
using u64 = unsigned long long;

__device__ __forceinline__ uint32_t lo32(u64 x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(u64 x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ u64 mk64(uint32_t lo, uint32_t hi) { return ((u64)hi << 32) | lo; }

// (hi << 1) | (lo >> 31) on the FMA pipe: hi * 2 + mulhi(lo, 2).  The ALU pipe (LOP3 / SHF / PRMT / ISETP, one warp
// instruction every other cycle) is what bounds the chain kernels; IMAD / IMAD.HI issue on the FMA pipe, which idles.
__device__ __forceinline__ uint32_t funnel1_fma(uint32_t lo, uint32_t hi)
{
#ifdef CUSTR_SHIFTS_ON_ALU
    return __funnelshift_l(lo, hi, 1);
#else
    uint32_t t, r;
    asm("mul.hi.u32 %0, %1, 2;" : "=r"(t) : "r"(lo));
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(hi), "r"(t));
    return r;
#endif
}
// move every bit one position up; bit 0 of lane 0 comes from the previous window (top word kept in last_hi)
__device__ __forceinline__ u64 adv64(u64 x, uint32_t last_hi, const LaneCtx& L)
{
#ifdef CUSTR_SHIFTS_ON_ALU
    uint32_t v = L.is31 ? last_hi : hi32(x);
#else
    // lane 31 hands the PREVIOUS window's top word to lane 0: v = hi + is31 * (last_hi - hi), an IMAD instead of a SEL
    uint32_t v = hi32(x) + L.m31 * (last_hi - hi32(x));
#endif
    uint32_t up = __shfl_sync(FULL, v, L.src);
    return mk64(funnel1_fma(up, lo32(x)), funnel1_fma(lo32(x), hi32(x)));
}
// (lo >> 1) | (hi << 31), also on the FMA pipe: hi * 2^31 + mulhi(lo, 2^31)
__device__ __forceinline__ uint32_t funnel1r_fma(uint32_t lo, uint32_t hi)
{
#ifdef CUSTR_SHIFTS_ON_ALU
    return __funnelshift_r(lo, hi, 1);
#else
    uint32_t t, r;
    asm("mul.hi.u32 %0, %1, 0x80000000;" : "=r"(t) : "r"(lo));
    asm("mad.lo.u32 %0, %1, 0x80000000, %2;" : "=r"(r) : "r"(hi), "r"(t));
    return r;
#endif
}
__device__ __forceinline__ u64 shift_down64(u64 x, uint32_t next_bit, const LaneCtx& L)
{
    uint32_t dn = __shfl_down_sync(FULL, lo32(x), 1);
    if (L.is31) dn = next_bit;
    return mk64(funnel1r_fma(lo32(x), hi32(x)), funnel1r_fma(hi32(x), dn));
}
// R[p] = Q[p] | (R[p-1] & K[p]) over the 2048 positions of the window; R[-1] = top bit of last_hi
__device__ __forceinline__ u64 spread64(u64 q, u64 k, uint32_t last_hi, const LaneCtx& L)
{
    u64 s = adv64(q, last_hi, L) & k;
    u64 sum = s + k;
    uint32_t g = __ballot_sync(FULL, sum < s);
    uint32_t p = __ballot_sync(FULL, sum == ~0ull);
    uint32_t g1 = g << 1, p1 = p << 1;
    uint32_t s2 = (g1 << 1) & p1;
    uint32_t c = g1 | ((((s2 + p1) ^ p1) | s2) & p1);
    sum += (c >> L.lane) & 1u;
    return q | (((sum ^ k) | s) & k);
}

struct Assertions64 {
    u64 rs, bow_b, bolc_b, nl, bow_a, lb, eold_a;
};
__device__ __forceinline__ u64 apply_before64_generic(u64 t, uint32_t m, const Assertions64& a)
{
    if (m & AS_BOW) t &= a.bow_b;
    if (m & AS_NBOW) t &= ~a.bow_b;
    if (m & AS_BOL_CARET) t &= a.bolc_b;
    if (m & AS_BOL_A) t &= a.rs;
    if (m & AS_EOL_DOLLAR) t &= a.nl;
    if (m & AS_EOL_Z) t = 0;
    return t;
}
__device__ __forceinline__ u64 apply_before64(u64 t, uint32_t m, const Assertions64& a)
{
    if (m == AS_BOW) return t & a.bow_b;  // the common single assertion inline, everything else out of line (code size)
    return apply_before64_generic(t, m, a);
}
__device__ __forceinline__ u64 apply_after64_generic(u64 t, uint32_t m, const Assertions64& a)
{
    if (m & AS_BOW) t &= a.bow_a;
    if (m & AS_NBOW) t &= ~a.bow_a;
    if (m & AS_BOL_CARET) t &= a.nl;
    if (m & AS_BOL_A) t = 0;
    if (m & AS_EOL_DOLLAR) t &= a.eold_a;
    if (m & AS_EOL_Z) t &= a.lb;
    return t;
}
__device__ __forceinline__ u64 apply_after64(u64 t, uint32_t m, const Assertions64& a)
{
    if (m == AS_BOW) return t & a.bow_a;
    return apply_after64_generic(t, m, a);
}
// EQ / RANGE atoms and multi-builtin classes: rare, kept out of the hot instruction stream
__device__ __forceinline__ u64 class_generic64(const ChainClassD& cc, const u64 (&p)[8], u64 letter5, u64 digit, u64 alnum, u64 word, u64 space)
{
    const uint32_t f = cc.builtins;
    u64 v = 0;
    if (f & (1u << AK_WORD)) v |= word;
    if (f & (1u << AK_ALNUM)) v |= alnum;
    if (f & (1u << AK_DIGIT)) v |= digit;
    if (f & (1u << AK_SPACE)) v |= space;
    if (f & (1u << AK_LOWER)) v |= p[6] & p[5] & letter5;
    if (f & (1u << AK_UPPER)) v |= p[6] & ~p[5] & letter5;
    if (f & (1u << AK_ANY)) v = ~0ull;
    for (uint32_t a = 0; a < cc.natoms; ++a) v |= cls_atom(p, cc.atoms[a]);
    return v;
}

template <int NCLS>
__device__ __forceinline__ u64 sel_class64(const u64 (&c)[NCLS], uint32_t k)
{
    if (NCLS == 1) return c[0];
    if (NCLS == 2) return k ? c[1] : c[0];
    u64 v = c[0];
#pragma unroll
    for (int i = 1; i < NCLS; ++i)
        if (k == (uint32_t)i) v = c[i];
    return v;
}

// Per-warp shared block.  It is addressed through ONE 32-bit shared-window address kept in a register (made opaque to
// the compiler, which otherwise re-derives it from %tid / the CTA's shared base at every use: ~35 instructions per
// window) plus immediates, with explicit ld/st.shared.
struct __align__(64) WarpSm64 {
    char ring[RING_STAGES][WIN64];
    uint32_t rs[64], f[64], d[64];
    uint32_t pc[32];  // count mode: exclusive prefix of popc over the lanes' words of the stream in f
};
constexpr uint32_t SM_RS = RING_STAGES * WIN64, SM_F = SM_RS + 256, SM_D = SM_F + 256, SM_C = SM_D + 256;
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ u64 lds64(uint32_t a)
{
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a) : "memory");
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t lo, uint32_t hi)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// bit `b` (0..2047) of a 64-word stream stored at shared address `a`
__device__ __forceinline__ uint32_t stream_bit(uint32_t a, int b) { return (lds32(a + 4u * (uint32_t)(b >> 5)) >> (b & 31)) & 1u; }

// Lane L keeps its 4 x 16-byte chunks XOR-swizzled: chunk k sits at 64L + 16 * (k ^ ((L >> 1) & 3)), so the 8 lanes of an
// LDS.128 phase hit 8 distinct 16-byte bank groups and chunk k's address is (chunk 0's address) ^ 16k.
__device__ __forceinline__ uint32_t ring_lane_offset(uint32_t lane) { return 64u * lane + 16u * ((lane >> 1) & 3u); }

// Non-ASCII bytes of this lane: decode each character once and give ALL its bytes (inside the lane) its class bits.
// Bytes are taken from the lane's own staged copy in shared memory (`my0` = shared address of the lane's chunk 0); only
// characters that straddle a lane boundary touch global memory.
template <int NCLS>
struct NaClasses {
    u64 c[NCLS];
    u64 al;
};
// shared address of byte `rel` (0..63) of this lane's 64 staged bytes.  LAYOUT 0 (k_chain64): `a0` = address of the lane's
// chunk 0, chunks XOR-swizzled; LAYOUT 1 (k_chain_item, regex_chain_item.cuh): `a0` = stage base + 128 * (lane >> 3), chunk
// kk of lane o sits at kk * 512 + 128 * (o >> 3) + 16 * ((o + 2 kk) & 7).
template <int LAYOUT>
__device__ __forceinline__ uint32_t ring_byte_addr(uint32_t a0, int rel)
{
    if (LAYOUT == 0) return (a0 ^ (uint32_t)(rel & 48)) + (uint32_t)(rel & 15);
    const uint32_t kk = (uint32_t)rel >> 4;
    return a0 + kk * 512u + 16u * ((lane_id() + 2u * kk) & 7u) + (uint32_t)(rel & 15);
}
template <int NCLS, int LAYOUT = 0>
__device__ __noinline__ NaClasses<NCLS> classify_non_ascii64(const ChainDev& cd, const Args& A, uint32_t my0, int lane_base, u64 na,
                                                             NaClasses<NCLS> r)
{
    const uint8_t* base = (const uint8_t*)A.chars;
    auto byte_at = [&](int pos) -> uint32_t {
        const int rel = pos - lane_base;
        if ((unsigned)rel < 64u) return lds8(ring_byte_addr<LAYOUT>(my0, rel));
        return pos < A.end ? base[pos] : 0u;
    };
    while (na) {
        const int b = __ffsll((long long)na) - 1;
        if (b < 63) {  // fast path: a 2-byte character (U+0080..U+07FF) that lies inside this lane's 64 bytes
            const uint32_t lead = lds8(ring_byte_addr<LAYOUT>(my0, b));
            const uint32_t next = lds8(ring_byte_addr<LAYOUT>(my0, b + 1));
            if ((lead & 0xE0u) == 0xC0u && (next & 0xC0u) == 0x80u) {
                const uint32_t cp = ((lead & 0x1Fu) << 6) | (next & 0x3Fu);
                const u64 bits = 3ull << b;
                na &= ~bits;
#pragma unroll
                for (int k = 0; k < NCLS; ++k)
                    if (k < (int)cd.nclasses) r.c[k] = ((cd.classes[k].na2[cp >> 5] >> (cp & 31)) & 1u) ? (r.c[k] | bits) : (r.c[k] & ~bits);
                r.al = ((cd.na2_alnum[cp >> 5] >> (cp & 31)) & 1u) ? (r.al | bits) : (r.al & ~bits);
                continue;
            }
        }
        int q = lane_base + b;
        while (q > A.first && (byte_at(q) & 0xC0u) == 0x80u) --q;  // only the leading continuation run has to walk back
        uint32_t ch = byte_at(q);
        const int w = utf8_width((uint8_t)ch);
        for (int k = 1; k < w; ++k) ch = (ch << 8) | byte_at(q + k);
        int lo = q - lane_base, hi = lo + w;  // bytes of the character, lane-relative
        if (lo < 0) lo = 0;
        if (hi > 64) hi = 64;
        if (hi <= b) hi = b + 1;  // malformed input: always make progress
        const u64 bits = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
        na &= ~bits;
        if (w == 2) {  // U+0080..U+07FF: exact bitmaps in the kernel parameter block, no memory traffic
            const uint32_t cp = ((ch >> 2) & 0x7C0u) | (ch & 0x3Fu);
#pragma unroll
            for (int k = 0; k < NCLS; ++k)
                if (k < (int)cd.nclasses) r.c[k] = ((cd.classes[k].na2[cp >> 5] >> (cp & 31)) & 1u) ? (r.c[k] | bits) : (r.c[k] & ~bits);
            r.al = ((cd.na2_alnum[cp >> 5] >> (cp & 31)) & 1u) ? (r.al | bits) : (r.al & ~bits);
            continue;
        }
#pragma unroll
        for (int k = 0; k < NCLS; ++k)
            if (k < (int)cd.nclasses) r.c[k] = na_char_matches(cd.classes[k], A, ch) ? (r.c[k] | bits) : (r.c[k] & ~bits);
        r.al = is_alnum_packed(ch, A.uflags) ? (r.al | bits) : (r.al & ~bits);
    }
    return r;
}

// where one lane's 64-bit words of the span streams go: this work item owns the bits of [byte_a, byte_b) — whole words
// are stored, boundary words OR-ed into the pre-cleared streams
struct SpanSink {
    u64* m;
    u64* k;
    u64* a;
    size_t widx;
    int wp, byte_a, byte_b;
};
static __device__ __noinline__ void store_spans(const SpanSink s, u64 m, u64 k, u64 a)
{
    if (s.wp >= s.byte_a && s.wp + 64 <= s.byte_b) {
        s.m[s.widx] = m;
        s.k[s.widx] = k;
        s.a[s.widx] = a;
    } else if (s.wp + 64 > s.byte_a && s.wp < s.byte_b) {
        u64 mask = ~0ull;
        if (s.wp < s.byte_a) mask &= ~0ull << (s.byte_a - s.wp);
        if (s.wp + 64 > s.byte_b) mask &= ~0ull >> (s.wp + 64 - s.byte_b);
        if (m & mask) atomicOr(s.m + s.widx, m & mask);
        if (k & mask) atomicOr(s.k + s.widx, k & mask);
        if (a & mask) atomicOr(s.a + s.widx, a & mask);
    }
}

template <int NS>
struct ChainState64 {  // top words of the previous window's streams (only their top bit is ever used)
    uint32_t last[NS];
    uint32_t last_al, last_nl, last_f, last_d;
};

// One code path for ASCII and UTF-8 windows (the UTF-8 extras sit behind the warp-uniform `utf8` flag): duplicating the
// chain for the two cases doubled the hot instruction footprint past the instruction cache.
template <int NS, int NCLS, int SPEC>
__device__ __forceinline__ u64 chain_eval64(const ChainDev& cd, const u64 (&c)[NCLS], u64 al, u64 nl, u64 rs, bool utf8, int rounds, u64 cont,
                                            uint32_t rs_next, uint32_t next_is_cont, uint32_t a_next, uint32_t nl_next,
                                            ChainState64<NS>& st, const LaneCtx& L, const SpanSink& sink)
{
    using PL = PlanLit<SPEC>;
    const u64 nrs = ~rs;
    u64 fin = ~0ull;       // last byte of a character
    uint32_t cont0 = 0u;   // window starts inside a character
    if (utf8) {
        fin = ~shift_down64(cont, next_is_cont, L);
        cont0 = __shfl_sync(FULL, lo32(cont), 0) & 1u;
    }
    Assertions64 as;
    as.rs = rs;
    as.nl = nl;
    as.bow_b = as.bow_a = as.bolc_b = as.lb = as.eold_a = 0;
    if (PV_NEEDS & (AS_BOW | AS_NBOW)) {
        as.bow_b = al ^ (adv64(al, st.last_al, L) & nrs);
        as.bow_a = al ^ shift_down64(al & nrs, a_next, L);
        st.last_al = hi32(al);
    }
    if (PV_NEEDS & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) {
        as.bolc_b = rs | (adv64(nl, st.last_nl, L) & nrs);
        as.lb = shift_down64(rs, rs_next, L);
        as.eold_a = as.lb | shift_down64(nl & nrs, nl_next, L);
        st.last_nl = hi32(nl);
    }
    // Optional steps (x?, x*): the positions ready for step s (`ready`) are also ready for step s+1.  Early exits (x{1,3},
    // trailing x?): a match may end behind every step that has an edge into END (`done`, "last consumed byte" domain).
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
            // the marker of step s-1 moves to the next position; the carry (previous window's stream) is dropped when
            // this window starts inside a character: that marker was not on a final byte
            t = adv64(P, cont0 ? 0u : old_prev, L) & nrs;
            if (PV_STEP_OPT(s - 1)) t |= ready;
        }
        ready = t;
        const u64 ck = sel_class64<NCLS>(c, PV_STEP_CLS(s));
        t &= ck & ~cont;
        const uint32_t old = st.last[s];
        u64 Z = t;
        if (PV_STEP_LOOP(s)) Z = spread64(t, ck & nrs, old, L);
        else if (utf8) {  // move the marker from the lead byte to the last byte of its character: one round per
                          // continuation byte of the longest character in the window (`rounds`, warp-uniform)
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) Z |= adv64(Z, old, L) & cont;
        }
        st.last[s] = hi32(Z);
        old_prev = old;
        P = Z & fin;
        if (PV_STEP_EXIT(s)) done |= P;
        if (s == NS - 1 && __builtin_expect(sink.m != nullptr, 0))  // span streams of the last step (span_walk.cuh); out of line: cold for contains_re / match
            store_spans(sink, t,                                                           // M: first character of the last step, lead byte
                        PV_STEP_LOOP(s) ? (ck & nrs) : (ck & cont),                       // K: the match may continue INTO this byte
                        PV_END_MASK ? apply_after64_generic(fin, PV_END_MASK, as) : fin);  // A: a match may end after this byte
    }
    return PV_END_MASK ? apply_after64(done, PV_END_MASK, as) : done;
}

// the window that holds the end of the buffer (once per column): chunks that exist completely are copied with cp.async, the
// others are zero-filled and patched with the bytes that exist (plain stores by the lane that will read them back)
static __device__ __noinline__ void ring_issue_tail(uint32_t dst0, const char* __restrict__ chars, int ws, int end, uint32_t lane)
{
    for (int k = 0; k < 4; ++k) {
        const int pos = ws + 64 * (int)lane + 16 * k;
        const uint32_t dst = dst0 ^ (16u * k);
        if (pos + 16 <= end) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(chars + pos) : "memory");
            continue;
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        for (int q = pos; q < end; ++q) asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst + (uint32_t)(q - pos)), "r"((uint32_t)(uint8_t)chars[q]) : "memory");
    }
}

// copy of window [ws, ws + 2048) into a ring stage; `dst0` = shared address of this lane's chunk 0 in that stage,
// `gsrc` = chars + 64 * lane
__device__ __forceinline__ void ring_issue(uint32_t dst0, const char* __restrict__ gsrc, const char* __restrict__ chars, int ws, int end, uint32_t lane)
{
    if (ws + WIN64 <= end) {  // whole window inside the buffer: plain 16-byte copies
        const char* src = gsrc + ws;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst0 ^ (16u * k)), "l"(src + 16 * k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        return;
    }
    ring_issue_tail(dst0, chars, ws, end, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NS, int NCLS, int SPEC = 0>
__global__ void __launch_bounds__(THREADS, 3)
k_chain64(const __grid_constant__ ChainDev cd, const __grid_constant__ Args A)
{
    using PL = PlanLit<SPEC>;
    __shared__ WarpSm64 sm[WARPS];
    LaneCtx L;
    L.lane = lane_id();
    asm volatile("" : "+r"(L.lane));
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    L.m31 = L.lane == 31 ? 1u : 0u;
    asm volatile("" : "+r"(L.m31));
    const uint32_t lane = L.lane;
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x >> 5]);  // this warp's block
    uint32_t my0 = wb + ring_lane_offset(lane);                                // my chunk 0 in stage 0
    uint32_t my_w = wb + 8u * lane;                                            // my 64-bit word of a stream (+ SM_RS / SM_F / SM_D)
    const char* gsrc = A.chars + 64 * (int)lane;
    asm volatile("" : "+r"(wb), "+r"(my0), "+r"(my_w));  // keep them in registers instead of re-deriving them per use
    const int warps_total = gridDim.x * WARPS;
    unsigned long long my_matches = 0;
    const uint32_t bneed = PV_BUILTIN_UNION | ((PV_NEEDS & (AS_BOW | AS_NBOW)) ? (1u << AK_ALNUM) : 0u);
    const bool need_nl = (PV_NEEDS & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) != 0;

    // work items are handed out dynamically (one atomicAdd per 32 KiB item): the grid is exactly the resident set, so
    // there is no partial last wave and the tail is a single item
    (void)warps_total;
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.item_counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= A.nitems) break;
        const int lo_byte = A.first + item * ITEM_BYTES;
        (void)lo_byte;
        const int ra = __ldg(A.item_bounds + item), rb = __ldg(A.item_bounds + item + 1);  // precomputed by k_item_bounds
        if (ra >= rb) continue;
        const int byte_a = __ldg(A.offsets + ra), byte_b = __ldg(A.offsets + rb);
        if (byte_a >= byte_b) continue;  // only empty rows: results stay 0 (pre-cleared)
        ChainState64<NS> st;
#pragma unroll
        for (int s = 0; s < NS; ++s) st.last[s] = 0;
        st.last_al = st.last_nl = st.last_f = st.last_d = 0;
        uint32_t d_live = 0;
        int carry = 0;  // count mode: matches of the straddling row found in earlier windows
        int ws = byte_a & ~(WIN64 - 1);
        int kcur = ra + 1;       // next offsets index to consume; offsets[j] >= ws for every j >= kcur
        int prev_o = byte_a;     // offsets[kcur - 1]
        int pend = byte_a - ws;  // window-relative position of a row start already known (-1: none)
        int o_nxt = (kcur + (int)lane <= rb) ? __ldg(A.offsets + kcur + (int)lane) : 0x7fffffff;
        int stage = 0;
        uint32_t nb_cur = 0;
        if constexpr (SPEC != 0) nb_cur = (ws + WIN64 < A.end) ? (uint32_t)(uint8_t)A.chars[ws + WIN64] : 0u;  // byte behind the first window
        __syncwarp();  // the previous item's reads of the ring are done
        ring_issue(my0, gsrc, A.chars, ws, A.end, lane);

        for (; ws < byte_b; ws += WIN64, stage ^= 1) {
            const int we = ws + WIN64;
            const bool more = we < byte_b;
            const uint32_t cur0 = my0 + (uint32_t)stage * WIN64;
            if (more) ring_issue(my0 + (uint32_t)(stage ^ 1) * WIN64, gsrc, A.chars, we, A.end, lane);  // next window in flight

            // ---- one pass over the offsets that fall into (ws, we]: ROWSTART bits now, row results after evaluation
            sts64(my_w + SM_RS, (lane == 0 && pend == 0) ? 1u : 0u, 0u);
            __syncwarp();
            if (pend > 0 && lane == 0) reds_or(wb + SM_RS + 4u * (uint32_t)(pend >> 5), 1u << (pend & 31));  // first window of the item
            const int j = kcur + (int)lane;
            const int o = o_nxt;
            const bool inw = o <= we;
            if (inw && o < we) reds_or(wb + SM_RS + 4u * (uint32_t)((o - ws) >> 5), 1u << ((o - ws) & 31));
            const unsigned m_in = __ballot_sync(FULL, inw);
            bool at_we = __any_sync(FULL, inw && o == we);
            int consumed = __popc(m_in);
            if (__builtin_expect(m_in == FULL, 0)) {  // more than 32 rows end in this window (short / empty rows): generic loop
                for (;;) {
                    int j2 = kcur + consumed + (int)lane;
                    int o2 = j2 <= rb ? __ldg(A.offsets + j2) : 0x7fffffff;
                    bool in2 = o2 <= we;
                    if (in2 && o2 < we) reds_or(wb + SM_RS + 4u * (uint32_t)((o2 - ws) >> 5), 1u << ((o2 - ws) & 31));
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
            const u64 rs = lds64(my_w + SM_RS);
            const uint32_t rs_next = at_we || we >= A.end;
            // the byte behind this window (look-ahead of \b / $): its load was issued one iteration ago — issued here it sat on
            // the critical path of the window (18 % of the stall samples)
            // (the generic kernels have enough work between the load and its use: there the extra live register costs more)
            uint32_t next_byte;
            if constexpr (SPEC != 0) {
                next_byte = (!rs_next && we < A.end) ? nb_cur : 0u;
                nb_cur = (we + WIN64 < A.end) ? (uint32_t)(uint8_t)A.chars[we + WIN64] : 0u;  // for the next window
            } else
                next_byte = (!rs_next && we < A.end) ? (uint32_t)(uint8_t)A.chars[we] : 0u;

            // ---- this window's bytes: wait for its cp.async group, read back my own 64 bytes, transpose to bit planes
            if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            u64 p[8];
            {
                uint32_t pl[8], ph[8];
                const uint4 v0 = lds128(cur0);
                const uint4 v1 = lds128(cur0 ^ 16u);
                transpose_planes(v0, v1, pl);
                const uint4 v2 = lds128(cur0 ^ 32u);
                const uint4 v3 = lds128(cur0 ^ 48u);
                transpose_planes(v2, v3, ph);
#pragma unroll
                for (int b = 0; b < 8; ++b) p[b] = mk64(pl[b], ph[b]);
            }
            const u64 na = p[7];
            const u64 zero = ~(p[0] | p[1] | p[2] | p[3] | p[4] | p[5] | p[6] | p[7]);
            const u64 letter5 = cls_letter5(p), digit = cls_digit(p);
            const u64 alnum = (p[6] & letter5) | digit, word = alnum | cls_underscore(p);
            u64 space = 0;
            if (bneed & (1u << AK_SPACE)) space = cls_space(p);
            NaClasses<NCLS> nc;
            u64 (&c)[NCLS] = nc.c;
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
                u64 v = 0;
                if (k < (int)PV_NCLASSES) {
                    const uint32_t f = PV_CLS_BUILTINS(k);
                    if (PV_CLS_NATOMS(k) == 0 && f == (1u << AK_WORD)) v = word;          // single builtin: inline
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
            const u64 nl = need_nl ? (cls_eq(p, '\n') & ~na) : 0ull;
            uint32_t a_next = 0;
            const uint32_t nl_next = next_byte == '\n';
            if (PV_NEEDS & (AS_BOW | AS_NBOW)) {
                if (next_byte < 0x80u) a_next = (next_byte - '0' < 10u) || ((next_byte | 0x20u) - 'a' < 26u);
                else if ((next_byte & 0xC0u) != 0x80u) {
                    int w;
                    a_next = is_alnum_packed(utf8_packed((const uint8_t*)A.chars + we, (const uint8_t*)A.chars + A.end, w), A.uflags);
                }
            }
            const bool utf8 = __any_sync(FULL, na != 0);
            u64 cont = 0;
            int rounds = 0;
            if (utf8) {
                if (na) nc = classify_non_ascii64<NCLS>(cd, A, cur0, ws + 64 * (int)lane, na, nc);
                cont = p[7] & ~p[6];
                const u64 lead3 = p[7] & p[6] & p[5];  // lead byte of a 3- or 4-byte character
                rounds = 1 + (int)__any_sync(FULL, lead3 != 0) + (int)__any_sync(FULL, (lead3 & p[4]) != 0);
            }
            const u64 al = nc.al;
            const SpanSink sink{A.span_m, A.span_k, A.span_a, (size_t)((ws + 64 * (int)lane - A.span_base) >> 6), ws + 64 * (int)lane, byte_a, byte_b};
            const u64 E = chain_eval64<NS, NCLS, SPEC>(cd, c, al, nl, rs, utf8, rounds, cont, rs_next, (next_byte & 0xC0u) == 0x80u, a_next, nl_next, st, L,
                                                 sink);

            // ---- sticky per-row OR of the match bits; NUL bytes make a row "dirty" (decided by the exact VM)
            // Count mode (count_re of chains whose steps all use the loop's class, e.g. \b\w{4,}\b, \d+): every match lies
            // inside one maximal run of the class and a run holds at most one match (its first admissible start, then the
            // LAST end in the run), so the number of matches of a row is the number of runs holding at least one match end
            // = the number of match-end bits that are the first one of their run:  V = E & ~(advance(spread(E, K)) & K).
            // A row's count is then a difference of prefix popcounts of V (+ a carry for the row that straddles windows).
            const u64 nrs = ~rs;
            const bool count_mode = A.counts != nullptr;
            u64 F;
            uint32_t vtotal = 0;
            if (__builtin_expect(!count_mode, 1)) {
                F = spread64(E, nrs, st.last_f, L);
                st.last_f = hi32(F);
            } else {
                const u64 K = sel_class64<NCLS>(c, PV_STEP_CLS(NS - 1)) & nrs;
                const uint32_t oldg = st.last_f;
                const u64 G = spread64(E, K, oldg, L);
                st.last_f = hi32(G);
                F = E & ~(adv64(G, oldg, L) & K);
                const uint32_t pcnt = (uint32_t)__popcll(F);
                uint32_t inc = pcnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, inc, d);
                    if ((int)lane >= d) inc += v;
                }
                vtotal = __shfl_sync(FULL, inc, 31);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(wb + SM_C + 4u * lane), "r"(inc - pcnt) : "memory");
            }
            auto prefix_at = [&](int x) -> int {  // V bits of this window at positions < x
                if (x >= WIN64) return (int)vtotal;
                const uint32_t l = (uint32_t)x >> 6;
                return (int)lds32(wb + SM_C + 4u * l) + __popcll(lds64(wb + SM_F + 8u * l) & ((1ull << (x & 63)) - 1ull));
            };
            auto row_result = [&](int o_begin, int o_end, int row, bool dirty) -> bool {  // row [o_begin, o_end), o_end in (ws, we]
                if (__builtin_expect(!count_mode, 1)) {
                    const bool hit = stream_bit(wb + SM_F, o_end - 1 - ws);
                    if (!dirty) A.out[row] = hit;
                    return hit;
                }
                const int x0 = o_begin - ws;
                const int cnt = prefix_at(o_end - ws) - (x0 > 0 ? prefix_at(x0) : 0) + (x0 < 0 ? carry : 0);
                if (!dirty) A.counts[row] = cnt;
                return cnt != 0;
            };
            sts64(my_w + SM_F, lo32(F), hi32(F));
            const bool any_dirty = __any_sync(FULL, zero != 0) || d_live;
            if (__builtin_expect(any_dirty, 0)) {
                const u64 D = spread64(zero, nrs, st.last_d, L);
                st.last_d = hi32(D);
                sts64(my_w + SM_D, lo32(D), hi32(D));
                d_live = __shfl_sync(FULL, hi32(D), 31) >> 31;
            }
            __syncwarp();

            // ---- finalise the rows whose last byte lies in this window (first chunk from registers)
            {
                int o_prev = __shfl_up_sync(FULL, o, 1);
                if (lane == 0) o_prev = prev_o;
                bool hit = false, dirty = false;
                if (inw && o > o_prev) {  // non-empty row j-1, last byte o-1 >= ws
                    dirty = any_dirty && stream_bit(wb + SM_D, o - 1 - ws);
                    hit = row_result(o_prev, o, j - 1, dirty);
                }
                if (__builtin_expect(any_dirty, 0)) {
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
            if (__builtin_expect(m_in == FULL, 0)) {  // remaining chunks: reload
                for (int k2 = kcur + 32; k2 < kcur + consumed; k2 += 32) {
                    const int j2 = k2 + (int)lane;
                    bool hit = false, dirty = false;
                    if (j2 < kcur + consumed) {
                        const int o2 = __ldg(A.offsets + j2), o2p = __ldg(A.offsets + j2 - 1);
                        if (o2 > o2p) {
                            dirty = any_dirty && stream_bit(wb + SM_D, o2 - 1 - ws);
                            hit = row_result(o2p, o2, j2 - 1, dirty);
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
            if (__builtin_expect(count_mode, 0)) carry = consumed ? (prev_o >= we ? 0 : (int)vtotal - prefix_at(prev_o - ws)) : carry + (int)vtotal;
            pend = at_we ? 0 : -1;
            __syncwarp();
        }
    }
    if (lane == 0 && my_matches) atomicAdd(A.total, my_matches);
}

#ifndef CUSTR_NO_LAUNCHERS
#ifndef CUSTR_EXPERIMENT_ONLY_4_1
template <int NS>
static void launch_chain64_ns(const ChainDev& cd, const Args& a, int blocks)
{
    auto k1 = k_chain64<NS, 1>;
    auto k2 = k_chain64<NS, 2>;
    auto k4 = k_chain64<NS, 4>;
    bool plain = true;  // no optional step, END only behind the last step
    for (uint32_t s = 0; s < cd.nsteps; ++s) plain = plain && !cd.steps[s].opt && ((cd.steps[s].exit != 0) == (s + 1 == cd.nsteps));
    if (!plain) {
        auto o1 = k_chain64<NS, 1, 5>;
        auto o2 = k_chain64<NS, 2, 5>;
        auto o4 = k_chain64<NS, 4, 5>;
        auto o8 = k_chain64<NS, 8, 5>;
        if (cd.nclasses <= 1) LAUNCH(o1, blocks, THREADS, 0, cd, a);
        else if (cd.nclasses == 2) LAUNCH(o2, blocks, THREADS, 0, cd, a);
        else if (cd.nclasses <= 4) LAUNCH(o4, blocks, THREADS, 0, cd, a);
        else LAUNCH(o8, blocks, THREADS, 0, cd, a);
        return;
    }
    if constexpr (NS <= 4) {  // shape specialisations exist for the short chains (\w+, \d{2,}, \b\w{4,}\b ...)
        const int spec = g_no_spec ? 0 : chain_spec_of(cd);
        auto s1 = k_chain64<NS, 1, 1>;
        auto s2 = k_chain64<NS, 1, 2>;
        auto s3 = k_chain64<NS, 1, 3>;
        auto s4 = k_chain64<NS, 1, 4>;
        auto s6 = k_chain64<NS, 1, 6>;
        auto s7 = k_chain64<NS, 1, 7>;
        if (spec == 6) { LAUNCH(s6, blocks, THREADS, 0, cd, a); return; }
        if (spec == 7) { LAUNCH(s7, blocks, THREADS, 0, cd, a); return; }
        if (spec == 1) { LAUNCH(s1, blocks, THREADS, 0, cd, a); return; }
        if (spec == 2) { LAUNCH(s2, blocks, THREADS, 0, cd, a); return; }
        if (spec == 3) { LAUNCH(s3, blocks, THREADS, 0, cd, a); return; }
        if (spec == 4) { LAUNCH(s4, blocks, THREADS, 0, cd, a); return; }
    }
    auto k8 = k_chain64<NS, 8>;
    if (cd.nclasses <= 1) LAUNCH(k1, blocks, THREADS, 0, cd, a);
    else if (cd.nclasses == 2) LAUNCH(k2, blocks, THREADS, 0, cd, a);
    else if (cd.nclasses <= 4) LAUNCH(k4, blocks, THREADS, 0, cd, a);
    else LAUNCH(k8, blocks, THREADS, 0, cd, a);
}
#endif

static void launch_chain64(const ChainDev& cd, const Args& a, int blocks)
{
#ifdef CUSTR_EXPERIMENT_ONLY_4_1  // quick SASS iteration on the headline instantiation (tools/sass_stat.sh)
    auto k1 = k_chain64<4, 1, CUSTR_EXPERIMENT_ONLY_4_1>;
    LAUNCH(k1, blocks, THREADS, 0, cd, a);
    return;
#else
    switch (cd.nsteps) {
    case 1: launch_chain64_ns<1>(cd, a, blocks); break;
    case 2: launch_chain64_ns<2>(cd, a, blocks); break;
    case 3: launch_chain64_ns<3>(cd, a, blocks); break;
    case 4: launch_chain64_ns<4>(cd, a, blocks); break;
    case 5: launch_chain64_ns<5>(cd, a, blocks); break;
    case 6: launch_chain64_ns<6>(cd, a, blocks); break;
    case 7: launch_chain64_ns<7>(cd, a, blocks); break;
    default: launch_chain64_ns<8>(cd, a, blocks); break;
    }
#endif
}
#endif  // CUSTR_NO_LAUNCHERS
