// Device-side building blocks shared by the bit-stream kernels (regex_bits.cu: DAG interpreter, 32/64-bit chain kernels,
// tokenize; regex_item.cu: item-buffered chain kernel): launch constants, the kernel argument block, bit-plane
// transposition, class formulas over planes, zero-width assertion streams.  Included inside namespace custr::bits.
#pragma once

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int WIN = 1024;
constexpr int ITEM_BYTES = 32 * 1024;
constexpr unsigned FULL = 0xffffffffu;

// carry bit layout (one uniform 32-bit mask per warp): bits 0..23 = P_s top bit, then:
constexpr int CY_A = 24, CY_NL = 25, CY_F = 26, CY_D = 27;

struct Args {
    const char* chars;        // column chars base (offsets are absolute into it)
    const int32_t* offsets;
    int32_t n;
    int32_t first, end;       // byte span [first, end)
    int32_t nitems;
    uint8_t* out;
    unsigned long long* total;
    int32_t* dirty_rows;
    unsigned int* dirty_count;
    unsigned int* item_counter;  // dynamic work distribution (k_chain64)
    const int32_t* item_bounds;  // item_bounds[t] = first row whose start offset is >= first + t*ITEM_BYTES (k_chain64)
    const uint8_t* prog_img;  // compiled program image (exact class tests for non-ASCII characters)
    const uint8_t* uflags;
    // span streams (count_re / replace_re of last-loop chains, span_walk.cuh): one bit per byte, word w covers the bytes
    // [span_base + 64 w, span_base + 64 w + 64); null = not wanted
    int32_t* counts;          // non-null: count mode of k_chain64 (matches per row instead of the boolean result)
    unsigned long long* span_m;
    unsigned long long* span_k;
    unsigned long long* span_a;
    int32_t span_base;
};

struct WarpSmem {
    uint32_t slot[MAX_STEPS][32];   // ADV_s = advance(P_s) & ~ROWSTART, per lane
    uint32_t cls[MAX_CLASSES][32];
    uint32_t rs[32];
    uint32_t f[32];
    uint32_t d[32];
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t advance(uint32_t x, uint32_t cin_bit)
{
    uint32_t up = __shfl_up_sync(FULL, x, 1);
    if (lane_id() == 0) up = cin_bit << 31;
    return __funnelshift_l(up, x, 1);
}
__device__ __forceinline__ uint32_t top_bit(uint32_t x) { return __shfl_sync(FULL, x, 31) >> 31; }

__device__ __forceinline__ uint32_t shift_down(uint32_t x, uint32_t next_bit)
{
    uint32_t dn = __shfl_down_sync(FULL, x, 1);
    if (lane_id() == 31) dn = next_bit;
    return __funnelshift_r(x, dn, 1);
}

// R[p] = Q[p] | (R[p-1] & K[p]) over the 1024 positions of the window, R[-1] = cin_bit
__device__ __forceinline__ uint32_t spread(uint32_t q, uint32_t k, uint32_t cin_bit)
{
    uint32_t s = advance(q, cin_bit) & k;
    uint32_t sum = s + k;
    uint32_t g = __ballot_sync(FULL, sum < s);
    uint32_t p = __ballot_sync(FULL, sum == 0xffffffffu);
    // carry into lane i: c[i] = g[i-1] | (p[i-1] & c[i-1])  -- same recurrence, solved with one 32-bit add
    uint32_t g1 = g << 1, p1 = p << 1;
    uint32_t s2 = (g1 << 1) & p1;
    uint32_t c = g1 | ((((s2 + p1) ^ p1) | s2) & p1);
    sum += (c >> lane_id()) & 1u;
    return q | (((sum ^ k) | s) & k);
}

// 32 bytes (8 little-endian words) -> 8 bit planes; bit i of plane b = bit b of byte i
__device__ __forceinline__ void transpose_planes(const uint4& lo, const uint4& hi, uint32_t (&p)[8])
{
    uint32_t a0 = lo.x, a1 = lo.z, a2 = hi.x, a3 = hi.z;  // words 0,2,4,6 -> bytes y, y+8, y+16, y+24 (y<4)
    uint32_t b0 = lo.y, b1 = lo.w, b2 = hi.y, b3 = hi.w;  // words 1,3,5,7 -> y = 4..7
    uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a0, a1, 0x7362);
    uint32_t t2 = __byte_perm(a2, a3, 0x5140), t3 = __byte_perm(a2, a3, 0x7362);
    // second level (16-bit halves of two words exchanged): lo = lo16(x) | y << 16, hi = x >> 16 | hi16(y) << 16: two PRMT.
    // -DCUSTR_HALFSWAP_FMA does it as 5 instructions on the FMA pipe (with h = x >> 16: lo = x + (y - h) * 65536 mod 2^32,
    // hi = h + (y >> 16) * 65536): measured slower (0.345 vs 0.339 ms on C2), the kernel is short of issue slots, not only of
    // ALU slots
#ifndef CUSTR_HALFSWAP_FMA
#define HALF_SWAP(X, Y, LO, HI) { LO = __byte_perm(X, Y, 0x5410); HI = __byte_perm(X, Y, 0x7632); }
#else
#define HALF_SWAP(X, Y, LO, HI)                                                                  \
    {                                                                                            \
        uint32_t h_, d_, yh_;                                                                    \
        asm("mul.hi.u32 %0, %1, 65536;" : "=r"(h_) : "r"(X));                                    \
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(d_) : "r"(h_), "r"(Y));                  \
        asm("mul.hi.u32 %0, %1, 65536;" : "=r"(yh_) : "r"(Y));                                   \
        asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(LO) : "r"(d_), "r"(X));                       \
        asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(HI) : "r"(yh_), "r"(h_));                     \
    }
#endif
    HALF_SWAP(t0, t2, p[0], p[1]) HALF_SWAP(t1, t3, p[2], p[3])
    t0 = __byte_perm(b0, b1, 0x5140); t1 = __byte_perm(b0, b1, 0x7362);
    t2 = __byte_perm(b2, b3, 0x5140); t3 = __byte_perm(b2, b3, 0x7362);
    HALF_SWAP(t0, t2, p[4], p[5]) HALF_SWAP(t1, t3, p[6], p[7])
#undef HALF_SWAP
    // one delta swap = 2 shifts + 2 bit-selects: (a & m) | (x & ~m) is a single LOP3 (LUT 0xE2) — ptxas does not
    // find it on its own when m and ~m are both immediates
#define DELTA_SWAP(A, B, S, M)                                                                              \
    {                                                                                                       \
        uint32_t na, nb, bs = (B) * (1u << (S)) /* IMAD: FMA pipe, the ALU pipe is the bottleneck */, as_ = __umulhi((A), 1u << (32 - (S))); \
        asm("lop3.b32 %0, %1, %2, %3, 0xE2;" : "=r"(na) : "r"(A), "r"(M), "r"(bs));                         \
        asm("lop3.b32 %0, %1, %2, %3, 0xE2;" : "=r"(nb) : "r"(as_), "r"(M), "r"(B));                        \
        (A) = na; (B) = nb;                                                                                 \
    }
    DELTA_SWAP(p[0], p[4], 4, 0x0F0F0F0Fu) DELTA_SWAP(p[1], p[5], 4, 0x0F0F0F0Fu)
    DELTA_SWAP(p[2], p[6], 4, 0x0F0F0F0Fu) DELTA_SWAP(p[3], p[7], 4, 0x0F0F0F0Fu)
    DELTA_SWAP(p[0], p[2], 2, 0x33333333u) DELTA_SWAP(p[1], p[3], 2, 0x33333333u)
    DELTA_SWAP(p[4], p[6], 2, 0x33333333u) DELTA_SWAP(p[5], p[7], 2, 0x33333333u)
    DELTA_SWAP(p[0], p[1], 1, 0x55555555u) DELTA_SWAP(p[2], p[3], 1, 0x55555555u)
    DELTA_SWAP(p[4], p[5], 1, 0x55555555u) DELTA_SWAP(p[6], p[7], 1, 0x55555555u)
#undef DELTA_SWAP
}

// ---- character classes as boolean formulas over the planes (ASCII, plane 7 ignored); T = uint32_t or uint64_t streams
template <typename T>
__device__ __forceinline__ T cls_digit(const T (&p)[8])
{
    return ~p[6] & p[5] & p[4] & (~p[3] | (~p[2] & ~p[1]));
}
template <typename T>
__device__ __forceinline__ T cls_letter5(const T (&p)[8])  // low five bits in 1..26
{
    T nz = p[4] | p[3] | p[2] | p[1] | p[0];
    T gt26 = p[4] & p[3] & (p[2] | (p[1] & p[0]));
    return nz & ~gt26;
}
template <typename T>
__device__ __forceinline__ T cls_alnum(const T (&p)[8]) { return (p[6] & cls_letter5(p)) | cls_digit(p); }
template <typename T>
__device__ __forceinline__ T cls_underscore(const T (&p)[8])
{
    return p[6] & ~p[5] & p[4] & p[3] & p[2] & p[1] & p[0];
}
template <typename T>
__device__ __forceinline__ T cls_space(const T (&p)[8])
{
    T hi0 = ~p[6] & ~p[5];
    T c9_13 = hi0 & ~p[4] & p[3] & ((~p[2] & (p[1] | p[0])) | (p[2] & ~p[1]));
    T c28_31 = hi0 & p[4] & p[3] & p[2];
    T c32 = ~p[6] & p[5] & ~(p[4] | p[3] | p[2] | p[1] | p[0]);
    return c9_13 | c28_31 | c32;
}
template <typename T>
__device__ __forceinline__ T cls_eq(const T (&p)[8], uint32_t c)
{
    T t = ~T(0);
#pragma unroll
    for (int b = 0; b < 7; ++b) t &= ((c >> b) & 1u) ? p[b] : ~p[b];
    return t;
}
// bytes >= c (7-bit compare, MSB first)
template <typename T>
__device__ __forceinline__ T cls_ge(const T (&p)[8], uint32_t c)
{
    T gt = 0, eq = ~T(0);
#pragma unroll
    for (int b = 6; b >= 0; --b) {
        if ((c >> b) & 1u) eq &= p[b];
        else { gt |= eq & p[b]; eq &= ~p[b]; }
    }
    return gt | eq;
}
template <typename T>
__device__ __forceinline__ T cls_atom(const T (&p)[8], const AtomD a)
{
    switch (a.kind) {
    case AK_EQ: return cls_eq(p, a.lo);
    case AK_RANGE: return cls_ge(p, a.lo) & ~(a.hi >= 127 ? T(0) : cls_ge(p, a.hi + 1u));
    case AK_WORD: return cls_alnum(p) | cls_underscore(p);
    case AK_ALNUM: return cls_alnum(p);
    case AK_DIGIT: return cls_digit(p);
    case AK_SPACE: return cls_space(p);
    case AK_LOWER: return p[6] & p[5] & cls_letter5(p);
    case AK_UPPER: return p[6] & ~p[5] & cls_letter5(p);
    default: return ~T(0);
    }
}

struct Assertions {  // zero-width assertion streams of the current window
    uint32_t rs, bow_b, bolc_b, nl, bow_a, lb, eold_a;
};
__device__ __forceinline__ uint32_t apply_before(uint32_t t, uint32_t m, const Assertions& a)
{
    if (m == AS_BOW) return t & a.bow_b;  // common single-assertion cases first
    if (m & AS_BOW) t &= a.bow_b;
    if (m & AS_NBOW) t &= ~a.bow_b;
    if (m & AS_BOL_CARET) t &= a.bolc_b;
    if (m & AS_BOL_A) t &= a.rs;
    if (m & AS_EOL_DOLLAR) t &= a.nl;
    if (m & AS_EOL_Z) t = 0;
    return t;
}
__device__ __forceinline__ uint32_t apply_after(uint32_t t, uint32_t m, const Assertions& a)
{
    if (m == AS_BOW) return t & a.bow_a;
    if (m & AS_BOW) t &= a.bow_a;
    if (m & AS_NBOW) t &= ~a.bow_a;
    if (m & AS_BOL_CARET) t &= a.nl;
    if (m & AS_BOL_A) t = 0;
    if (m & AS_EOL_DOLLAR) t &= a.eold_a;
    if (m & AS_EOL_Z) t &= a.lb;
    return t;
}

// first index r in [0, n] with offsets[r] >= target (offsets has n+1 ascending entries); warp-cooperative 32-ary search
static __device__ int warp_lower_bound(const int32_t* __restrict__ offsets, int n, int target)
{
    int lo = 0, hi = n + 1;  // answer in [lo, hi]
    while (hi - lo > 0) {
        int span = hi - lo;
        int stepsz = (span + 31) / 32;
        int idx = lo + (int)lane_id() * stepsz;
        bool ge = idx >= hi ? true : (__ldg(offsets + idx) >= target);
        unsigned m = __ballot_sync(FULL, ge);
        int first_ge = m ? __ffs(m) - 1 : 32;   // lanes before it are < target
        int new_hi = lo + first_ge * stepsz;
        if (new_hi > hi) new_hi = hi;
        int new_lo = first_ge == 0 ? lo : lo + (first_ge - 1) * stepsz + 1;
        if (first_ge == 0) return lo;
        lo = new_lo;
        hi = new_hi;
        if (stepsz == 1) return hi;
    }
    return lo;
}

