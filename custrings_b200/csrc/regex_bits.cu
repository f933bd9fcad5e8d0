// Bitstream regex executor (device).  Model, plan format and the equivalence argument: regex_bits_plan.h;
// host lowering: regex_bits_lower.cpp.  DESIGN.md §4.2 has the roofline accounting.
//
// Mapping to the machine
//   * the flat chars buffer is cut into 1024-byte WINDOWS; one warp evaluates one window at a time, lane L owning the
//     32 bytes [32L, 32L+32) = one 32-bit word of every stream.  Loads are two LDG.128 per lane, fully coalesced,
//     software-prefetched one window ahead; no byte is read twice except the two boundary windows of a work item.
//   * the 32 bytes of a lane are bit-transposed in registers (16 PRMT + 48 SHF/LOP3) into 8 bit planes; character
//     classes are boolean formulas over the planes, so classification costs a few LOP3 per 32 bytes instead of a
//     table lookup per byte.
//   * advance / look-ahead are a warp shuffle + funnel shift; `spread` (x+, x*, and the per-row sticky OR) is a
//     1024-bit carry-propagating add done with two ballots; inter-window state is one carry bit per stream.
//   * rows never interact: ROWSTART bits are scattered from the offsets array (shared-memory atomicOr); a work item is a
//     contiguous ROW range whose bytes start near a 32 KiB boundary (warp-cooperative 32-ary search in offsets), so
//     warps never exchange state.
//   * rows holding a byte >= 0x80 or a NUL are appended to a work list and decided by the exact Pike-VM kernel.
#include "regex_bits.h"
#include "regex_bits_plan.h"
#include "regex_vm.cuh"
#include <cub/cub.cuh>
#include <cstddef>

namespace custr {
namespace bits {

struct Plan {
    PlanDev dev;
    bool is_chain = false;
    bool span_ok = false;
    bool chain_has_opt = false;
    ChainDev chain;
    std::string text;
};

bool g_no_spec = false;        // A/B switch: never use the shape-specialised chain kernels
bool g_force_generic = false;  // A/B switch: run chain-shaped plans on the generic interpreter kernel
bool g_chain32 = false;        // A/B switch: 32-bit-stream chain kernel instead of the 64-bit one

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
    p[0] = __byte_perm(t0, t2, 0x5410); p[1] = __byte_perm(t0, t2, 0x7632);
    p[2] = __byte_perm(t1, t3, 0x5410); p[3] = __byte_perm(t1, t3, 0x7632);
    t0 = __byte_perm(b0, b1, 0x5140); t1 = __byte_perm(b0, b1, 0x7362);
    t2 = __byte_perm(b2, b3, 0x5140); t3 = __byte_perm(b2, b3, 0x7362);
    p[4] = __byte_perm(t0, t2, 0x5410); p[5] = __byte_perm(t0, t2, 0x7632);
    p[6] = __byte_perm(t1, t3, 0x5410); p[7] = __byte_perm(t1, t3, 0x7632);
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
__device__ int warp_lower_bound(const int32_t* __restrict__ offsets, int n, int target)
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

__device__ __forceinline__ void load_window(const char* __restrict__ chars, int ws, int end, uint4& lo, uint4& hi)
{
    int b = ws + 32 * (int)lane_id();
    if (b + 32 <= end) {
        const uint4* q = (const uint4*)(chars + b);
        lo = __ldg(q);
        hi = __ldg(q + 1);
    } else {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t v = 0;
            for (int k = 0; k < 4; ++k) {
                int pos = b + 4 * j + k;
                if (pos < end && pos >= 0) v |= (uint32_t)(uint8_t)chars[pos] << (8 * k);
            }
            w[j] = v;
        }
        lo = make_uint4(w[0], w[1], w[2], w[3]);
        hi = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

__global__ void __launch_bounds__(THREADS)
k_bitstream(const __grid_constant__ PlanDev plan, const Args A)
{
    __shared__ WarpSmem smem[WARPS];
    WarpSmem& S = smem[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const int warps_total = gridDim.x * WARPS;
    unsigned long long my_matches = 0;

    for (int item = blockIdx.x * WARPS + (threadIdx.x >> 5); item < A.nitems; item += warps_total) {
        // ---- rows of this item: first row starting at or after the item's byte boundary
        const int lo_byte = A.first + item * ITEM_BYTES;
        int ra = item == 0 ? 0 : warp_lower_bound(A.offsets, A.n, lo_byte);
        int rb = item == A.nitems - 1 ? A.n : warp_lower_bound(A.offsets, A.n, lo_byte + ITEM_BYTES);
        if (ra >= rb) continue;
        const int byte_a = __ldg(A.offsets + ra), byte_b = __ldg(A.offsets + rb);
        if (byte_a >= byte_b) continue;  // only empty rows: results stay 0 (pre-cleared)
        int krs = ra;       // next offset index whose ROWSTART bit is not set yet
        int kfin = ra + 1;  // next offset index j whose row j-1 is not finalised yet
        uint32_t carry = 0;
        int ws = byte_a & ~(WIN - 1);
        uint4 cur_lo, cur_hi, nxt_lo, nxt_hi;
        load_window(A.chars, ws, A.end, cur_lo, cur_hi);

        for (; ws < byte_b; ws += WIN) {
            const int we = ws + WIN;
            if (we < byte_b) load_window(A.chars, we, A.end, nxt_lo, nxt_hi);  // prefetch

            // ---- ROWSTART bits from the offsets array
            S.rs[lane] = 0;
            __syncwarp();
            for (;;) {
                int j = krs + (int)lane;
                int o = j <= rb ? __ldg(A.offsets + j) : 0x7fffffff;
                bool in = o < we;
                if (in && o >= ws) atomicOr(&S.rs[(o - ws) >> 5], 1u << ((o - ws) & 31));
                unsigned m = __ballot_sync(FULL, in);
                krs += __popc(m);
                if (m != FULL) break;
            }
            __syncwarp();
            Assertions as;
            as.rs = S.rs[lane];
            const uint32_t nrs = ~as.rs;
            const uint32_t rs_next = (krs <= rb && __ldg(A.offsets + krs) == we) || we >= A.end;
            const uint32_t next_byte = (!rs_next && we < A.end) ? (uint8_t)A.chars[we] : 0;

            // ---- bit planes, dirty bytes
            uint32_t p[8];
            transpose_planes(cur_lo, cur_hi, p);
            const uint32_t dirty_bits = p[7] | ~(p[0] | p[1] | p[2] | p[3] | p[4] | p[5] | p[6] | p[7]);

            // ---- class streams
            for (int k = 0; k < plan.nclasses; ++k) {
                uint32_t v = 0;
                for (int a = 0; a < plan.classes[k].natoms; ++a) v |= cls_atom(p, plan.classes[k].atoms[a]);
                if (plan.classes[k].negate) v = ~v;
                S.cls[k][lane] = v;
            }
            // ---- assertion streams (only what the plan uses)
            uint32_t carry_out = 0;
            as.bow_b = as.bow_a = as.bolc_b = as.nl = as.lb = as.eold_a = 0;
            const uint32_t needs = plan.before_needs | plan.after_needs;
            if (needs & (AS_BOW | AS_NBOW)) {
                uint32_t al = cls_alnum(p);
                as.bow_b = al ^ (advance(al, (carry >> CY_A) & 1u) & nrs);
                carry_out |= top_bit(al) << CY_A;
                uint32_t nb = next_byte | 0x20u;
                uint32_t a_next = (next_byte - '0' < 10u) || (nb - 'a' < 26u);
                as.bow_a = al ^ shift_down(al & nrs, a_next);
            }
            if (needs & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) {
                as.nl = cls_eq(p, '\n');
                as.bolc_b = as.rs | (advance(as.nl, (carry >> CY_NL) & 1u) & nrs);
                carry_out |= top_bit(as.nl) << CY_NL;
                as.lb = shift_down(as.rs, rs_next);
                as.eold_a = as.lb | shift_down(as.nl & nrs, next_byte == '\n');
            }
            __syncwarp();

            // ---- marker steps
            uint32_t E = 0;
            const uint32_t start = plan.anchored ? as.rs : 0xffffffffu;
            int end_i = 0;
            for (int s = 0; s < plan.nsteps; ++s) {
                const StepD st = plan.steps[s];
                const uint32_t c = S.cls[st.cls][lane];
                uint32_t entry = 0;
                for (int j = 0; j < st.npreds; ++j) {
                    const PredD pr = st.preds[j];
                    uint32_t t = pr.src == SRC_START ? start : S.slot[pr.src][lane];
                    entry |= apply_before(t, pr.mask, as);
                }
                uint32_t P = entry & c;
                if (st.self_loop) P = spread(P, apply_before(c & nrs, st.self_mask, as), (carry >> s) & 1u);
                while (end_i < plan.nends && plan.ends[end_i].src == s) E |= apply_after(P, plan.ends[end_i++].mask, as);
                S.slot[s][lane] = advance(P, (carry >> s) & 1u) & nrs;
                carry_out |= top_bit(P) << s;
            }

            // ---- sticky per-row OR of matches (and of dirty bytes when there are any)
            uint32_t F = spread(E, nrs, (carry >> CY_F) & 1u);
            carry_out |= top_bit(F) << CY_F;
            S.f[lane] = F;
            const bool any_dirty = __any_sync(FULL, dirty_bits != 0) || ((carry >> CY_D) & 1u);
            if (any_dirty) {
                uint32_t D = spread(dirty_bits, nrs, (carry >> CY_D) & 1u);
                carry_out |= top_bit(D) << CY_D;
                S.d[lane] = D;
            }
            __syncwarp();

            // ---- finalise the rows whose last byte lies in this window
            for (;;) {
                int j = kfin + (int)lane;
                int o = j <= rb ? __ldg(A.offsets + j) : 0x7fffffff;
                bool in = o <= we;
                bool hit = false, dirty = false;
                if (in) {
                    int o_prev = __ldg(A.offsets + j - 1);
                    if (o > o_prev) {  // non-empty row j-1, last byte o-1 >= ws
                        int b = o - 1 - ws;
                        hit = (S.f[b >> 5] >> (b & 31)) & 1u;
                        dirty = any_dirty && ((S.d[b >> 5] >> (b & 31)) & 1u);
                        if (!dirty) A.out[j - 1] = hit;
                    }
                }
                unsigned dm = __ballot_sync(FULL, dirty);
                if (dm) {
                    unsigned basei = 0;
                    if (lane == 0) basei = atomicAdd(A.dirty_count, __popc(dm));
                    basei = __shfl_sync(FULL, basei, 0);
                    if (dirty) A.dirty_rows[basei + __popc(dm & ((1u << lane) - 1))] = j - 1;
                }
                my_matches += __popc(__ballot_sync(FULL, hit && !dirty));
                unsigned m = __ballot_sync(FULL, in);
                kfin += __popc(m);
                if (m != FULL) break;
            }
            __syncwarp();
            carry = carry_out;
            cur_lo = nxt_lo;
            cur_hi = nxt_hi;
        }
    }
    if (lane == 0 && my_matches) atomicAdd(A.total, my_matches);
}


// item_bounds[t] = first row r with offsets[r] >= first + t*ITEM_BYTES, for t = 0..nitems (bounds[nitems] = n): one coalesced
// pass over the offsets instead of two dependent binary searches at the head of every work item
__global__ void k_item_bounds(const int32_t* __restrict__ offsets, int n, int first, int nitems, int32_t* __restrict__ bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // row index 0..n (offsets has n+1 entries)
    if (i > n) return;
    const long long cur = (long long)offsets[i] - first;
    const long long prev = i == 0 ? -1 : (long long)offsets[i - 1] - first;
    // boundaries t with prev < t*ITEM <= cur get row i
    long long t_lo = prev < 0 ? 0 : prev / ITEM_BYTES + 1;
    long long t_hi = cur / ITEM_BYTES;
    if (t_hi > nitems) t_hi = nitems;
    for (long long t = t_lo; t <= t_hi; ++t) bounds[t] = i;
    if (i == n) {  // boundaries past the last offset (only bounds[nitems] when the span is not a multiple of the item size)
        for (long long t = t_hi + 1; t <= nitems; ++t) bounds[t] = n;
        bounds[nitems] = n;
    }
}

#include "regex_chain.cuh"
#include "regex_chain64.cuh"

const PlanDev& device_plan(const Plan& plan);

bool plan_is_chain(const Plan& plan) { return plan.is_chain; }

// count_re inside the chain kernel (see k_chain64's count mode): the last step loops and every step uses its class
bool count_in_kernel_ok(const Plan& plan)
{
    if (!plan.is_chain || !plan.span_ok) return false;
    const ChainDev& cd = plan.chain;
    if (cd.nsteps == 0 || !cd.steps[cd.nsteps - 1].loop) return false;
    for (uint32_t s = 0; s < cd.nsteps; ++s)
        if (cd.steps[s].cls != cd.steps[cd.nsteps - 1].cls) return false;
    return true;
}

bool run(const Plan& plan, const custr_column* col, const uint8_t* prog_img, const uint8_t* uflags, uint8_t* out,
         unsigned long long* total, int32_t** dirty_rows, unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count,
         SpanStreams* spans)
{
    if (spans && !(plan.is_chain && plan.span_ok && !g_force_generic && !g_chain32)) return false;
    int32_t* counts = spans ? spans->counts_out : nullptr;
    if (counts && !count_in_kernel_ok(plan)) return false;

    const int32_t n = col->n;
    if (((uintptr_t)col->chars & 15) != 0) return false;  // vector loads need a 16-byte aligned base
    if (counts) CUSTR_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n, g_stream));
    else CUSTR_CUDA(cudaMemsetAsync(out, 0, (size_t)n, g_stream));
    keep_rows = dev_alloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    keep_count = dev_alloc(2 * sizeof(unsigned int));  // [0] dirty-row count, [1] work-item counter
    CUSTR_CUDA(cudaMemsetAsync(keep_count->ptr, 0, 2 * sizeof(unsigned int), g_stream));
    *dirty_rows = (int32_t*)keep_rows->ptr;
    *dirty_count = (unsigned int*)keep_count->ptr;
    if (col->nbytes == 0) return true;
    Args a;
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    a.nitems = (int)((col->nbytes + ITEM_BYTES - 1) / ITEM_BYTES);
    a.out = out;
    a.total = total;
    a.dirty_rows = *dirty_rows;
    a.dirty_count = *dirty_count;
    a.item_counter = (unsigned int*)keep_count->ptr + 1;
    a.item_bounds = nullptr;
    a.prog_img = prog_img;
    a.uflags = uflags;
    a.span_m = a.span_k = a.span_a = nullptr;
    a.span_base = 0;
    a.counts = counts;
    if (spans && !counts) {
        a.span_base = a.first & ~(WIN64 - 1);
        const size_t words = (((size_t)(a.end - a.span_base) + WIN64 - 1) / WIN64) * (WIN64 / 64);
        spans->keep = dev_alloc(3 * words * sizeof(unsigned long long));
        CUSTR_CUDA(cudaMemsetAsync(spans->keep->ptr, 0, 3 * words * sizeof(unsigned long long), g_stream));
        a.span_m = (unsigned long long*)spans->keep->ptr;
        a.span_k = a.span_m + words;
        a.span_a = a.span_k + words;
        spans->m = a.span_m;
        spans->k = a.span_k;
        spans->a = a.span_a;
        spans->base = a.span_base;
    }
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (plan.is_chain && !g_force_generic) {
#ifndef CUSTR_EXPERIMENT_ONLY_4_1
        if (g_chain32 && !plan.chain_has_opt && plan.chain.nclasses <= 4) launch_chain(plan.chain, a, blocks);   // 1024-byte windows, 32-bit streams (A/B)
        else
#endif
        {  // 2048-byte windows, 64-bit streams, cp.async ring; grid = resident set (3 CTAs per SM), dynamic items
            if (!col->item_bounds || col->item_bounds_count != a.nitems) {  // once per column (it is immutable)
                col->item_bounds = dev_alloc(sizeof(int32_t) * (size_t)(a.nitems + 2));
                col->item_bounds_count = a.nitems;
                LAUNCH(k_item_bounds, (n + 1 + 255) / 256, 256, 0, a.offsets, n, a.first, a.nitems, (int32_t*)col->item_bounds->ptr);
            }
            a.item_bounds = (const int32_t*)col->item_bounds->ptr;
            int resident = num_sms() * 3;
            launch_chain64(plan.chain, a, blocks < resident ? blocks : resident);
        }
        return true;
    }
#ifndef CUSTR_EXPERIMENT_ONLY_4_1
    LAUNCH(k_bitstream, blocks, THREADS, 0, device_plan(plan), a);
#endif
    return true;
}

#include "tokenize_bits.cuh"

// NVText::tokenize through the bit-stream compaction kernels.  delims == nullptr: whitespace.  False = not applicable
// (unaligned chars base, empty column): the caller uses the per-row path.
bool tokenize_flat(const custr_column* col, const uint8_t* delims, int ndelims, BufPtr& out_chars, BufPtr& out_off, int64_t& ntok, int64_t& nbytes)
{
    if (((uintptr_t)col->chars & 15) != 0 || col->nbytes == 0 || col->n == 0 || ndelims > TOK_DELIMS_MAX) return false;
    const int32_t n = col->n;
    TokArgs a{};
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    a.nitems = (int)((col->nbytes + ITEM_BYTES - 1) / ITEM_BYTES);
    a.whitespace = delims ? 0u : 1u;
    a.ndelims = delims ? (uint32_t)ndelims : 0u;
    for (int k = 0; k < ndelims && delims; ++k) a.delims[k] = delims[k];
    if (!col->item_bounds || col->item_bounds_count != a.nitems) {
        col->item_bounds = dev_alloc(sizeof(int32_t) * (size_t)(a.nitems + 2));
        col->item_bounds_count = a.nitems;
        LAUNCH(k_item_bounds, (n + 1 + 255) / 256, 256, 0, a.offsets, n, a.first, a.nitems, (int32_t*)col->item_bounds->ptr);
    }
    a.item_bounds = (const int32_t*)col->item_bounds->ptr;
    const int win_base = a.first & ~(WIN64 - 1);
    const size_t nwin = ((size_t)(a.end - win_base) + WIN64 - 1) / WIN64;
    // every item adds at most one shared window; + one empty slot whose exclusive sum is the grand total
    const size_t nslots = nwin + (size_t)a.nitems + 1;
    Scratch<int32_t> item_w((size_t)a.nitems + 1), item_slot((size_t)a.nitems + 1);
    CUSTR_CUDA(cudaMemsetAsync(item_w.get() + a.nitems, 0, sizeof(int32_t), g_stream));
    LAUNCH(k_tok_item_windows, (a.nitems + 255) / 256, 256, 0, a.offsets, a.item_bounds, a.nitems, item_w.get());
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream);
        BufPtr t = dev_alloc(tb);
        CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(t->ptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    a.item_slot = item_slot.get();
    BufPtr counts = dev_alloc(sizeof(unsigned long long) * nslots), base = dev_alloc(sizeof(unsigned long long) * nslots);
    BufPtr counter = dev_alloc(2 * sizeof(unsigned int));
    CUSTR_CUDA(cudaMemsetAsync(counts->ptr, 0, sizeof(unsigned long long) * nslots, g_stream));
    CUSTR_CUDA(cudaMemsetAsync(counter->ptr, 0, 2 * sizeof(unsigned int), g_stream));
    a.slot_counts = (unsigned long long*)counts->ptr;
    a.slot_base = (const unsigned long long*)base->ptr;
    const int smem = WARPS * (int)sizeof(WarpSmTok);
    // (per device and cheap: set on every call rather than caching a flag that would be wrong after custr_set_device)
    CUSTR_CUDA(cudaFuncSetAttribute(k_tokenize64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUSTR_CUDA(cudaFuncSetAttribute(k_tokenize64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    const int resident = num_sms() * 3;
    if (blocks > resident) blocks = resident;
    a.item_counter = (unsigned int*)counter->ptr;
    auto kc = k_tokenize64<false>;
    LAUNCH(kc, blocks, THREADS, smem, a);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(tmp->ptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long total = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&total, (unsigned long long*)base->ptr + (nslots - 1), sizeof(total), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    ntok = (int64_t)(total >> 32);
    nbytes = (int64_t)(total & 0xffffffffull);
    out_chars = dev_alloc((size_t)nbytes);
    out_off = dev_alloc(sizeof(int32_t) * (size_t)(ntok + 1));
    const int32_t last = (int32_t)nbytes;
    CUSTR_CUDA(cudaMemcpyAsync((int32_t*)out_off->ptr + ntok, &last, sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    if (ntok) {
        a.tok_off = (int32_t*)out_off->ptr;
        a.out = (char*)out_chars->ptr;
        a.item_counter = (unsigned int*)counter->ptr + 1;
        auto kw = k_tokenize64<true>;
        LAUNCH(kw, blocks, THREADS, smem, a);
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));  // `last` and the scratch buffers die with this scope
    return true;
}

}  // namespace bits
}  // namespace custr
