// Helpers of the linear-chain bitstream kernels (regex_chain64.cuh, regex_chain_item.cuh) — included inside namespace
// custr::bits.  (The first, 32-bit-stream chain kernel that lived here was superseded by the 64-bit kernels and removed.)
//
// The chain kernels are the specialised, UTF-8 aware executors for the common pattern shape (class sequences with x+ / x* tails and leading /
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
    uint32_t m31;        // 1 in lane 31, else 0 (arithmetic select on the FMA pipe)
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
#ifdef CUSTR_JIT  // run-time compiled kernels are only built for plans whose classes carry their inline definition
    case NA_CLASS: return na_class_inline(cd, A.uflags, ch);
    case NA_NCLASS: return !na_class_inline(cd, A.uflags, ch);
#else
    case NA_CLASS:
        return cd.na_inline ? na_class_inline(cd, A.uflags, ch)
                            : rxdev::class_match(rxdev::bind_program(A.prog_img, A.uflags), (int)cd.na_arg, ch);
    case NA_NCLASS:
        return !(cd.na_inline ? na_class_inline(cd, A.uflags, ch)
                              : rxdev::class_match(rxdev::bind_program(A.prog_img, A.uflags), (int)cd.na_arg, ch));
#endif
    default: return false;
    }
}
