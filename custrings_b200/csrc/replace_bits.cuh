// NVStrings::replace with a literal target and no replacement limit (strings/replace.cu:37-148; per row
// custring_view.inl:1004-1060: leftmost non-overlapping occurrences, byte compare) as a bit-stream splice — included by
// regex_bits.cu inside custr::bits after split_bits.cuh, whose window machinery it shares.
//
// In stream terms, with RS = first byte of every row and E_k = "this byte equals target byte k":
//     M    = starts of occurrences = E_0 & (E_1 >> 1) & ... & (E_{m-1} >> (m-1)), no RS bit inside (start, start + m)
//     DROP = the m bytes of every occurrence (M smeared upwards by m - 1)
//     output chars    = the bytes of ~DROP in order, with the replacement spliced in at every M bit
//     new offsets[r]  = (kept bytes + rlen * occurrences) before the first byte of row r
// The host takes this path only for a BORDER-FREE target (no proper prefix equal to a suffix, e.g. "ab", " ", "the"; not "aa" or
// "abab"): occurrences of such a target cannot overlap, so every occurrence is one the per-row leftmost scan would take.
//
// Windows of one work item overlap by one lane: the stride is 31 lanes (1984 bytes) and lane 31 is look-ahead only, so an
// occurrence that starts in an owned byte always has its m <= 32 bytes (and the row starts among them) inside the window.
// What flows forward between windows is the number of leading bytes still covered by the last occurrence of the previous one.
// Two passes like split: output bytes per (item, window) slot -> exclusive scan -> write.
//
// MODE 1 — replace_re for single-class chains (x{n,} / x+ with any leading and trailing assertions, e.g. \b\w{4,}\b, \d+, \s+):
// the same splice fed by the chain kernel's span streams (regex_bits.h SpanStreams: M = a match can begin its LAST step here,
// K = the match may continue into this byte, A = a match may end behind this byte) instead of a literal compare.  Every step has
// the class of the loop, so the k characters in front of an M bit lie in the same maximal K-run as the bit itself and a run
// holds at most one match (span_walk.cuh's rule "first M at or after the cursor, then the last A of the run" never finds a
// second one: behind that A the run has no A left).  In stream terms:
//     S   = M spread through K                      (forward carry between windows)
//     FM  = M & ~(S advanced & K)                   first M of every run
//     R   = A spread BACKWARDS through K            "an A follows inside the run" (carry from behind the window: a short scan
//                                                   of the K / A words that follow, the streams are in device memory)
//     occurrence starts = FM & R;  DROP = (S & R) | the k characters in front of every FM & R bit
#pragma once

constexpr int REPL_STRIDE = WIN64 - 64;
constexpr int REPL_TILE = 4096;
constexpr int REPL_PAT_MAX = 32, REPL_REPL_MAX = 64;

struct ReplArgs {
    const char* chars;
    const int32_t* offsets;
    int32_t n, first, end, nitems;
    unsigned int* item_counter;
    const int32_t* item_bounds;
    const int32_t* item_slot;
    unsigned long long* slot_counts;      // count pass: output bytes per slot
    const unsigned long long* slot_base;  // write pass: exclusive scan of slot_counts
    int32_t m, rlen;                      // MODE 0: target length; MODE 1: shortest match in bytes (tile bound only)
    uint8_t pat[REPL_PAT_MAX];
    uint8_t repl[REPL_REPL_MAX];
    int32_t* new_off;                     // write pass outputs
    char* out;
    // MODE 1: span streams of the chain kernel, one bit per byte from byte span_base (a multiple of 2048) on
    const unsigned long long* span_m;
    const unsigned long long* span_k;
    const unsigned long long* span_a;
    long long span_words;
    int32_t span_base, k_chars;
    unsigned int* flags;                  // bit 0: a K-run longer than the look-ahead bound (caller falls back)
};

struct __align__(64) WarpSmRepl {
    char ring[RING_STAGES][WIN64];
    uint32_t rs[64];    // ROWSTART (every row), one bit per byte of the window
    uint32_t kk[64];    // kept bytes this window owns
    uint32_t mm[64];    // occurrence starts this window owns
    uint32_t pre[32];   // exclusive prefix over the lanes of the output bytes
    char tile[tile_padded_bytes(REPL_TILE + 32)];
};

// number of windows every work item touches (0 for an item without bytes)
__global__ void k_repl_item_windows(const int32_t* __restrict__ offsets, const int32_t* __restrict__ item_bounds, int nitems, int align_mask,
                                    int32_t* __restrict__ out)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const int ra = item_bounds[item], rb = item_bounds[item + 1];
    int w = 0;
    if (ra < rb) {
        const int a = offsets[ra], b = offsets[rb];
        if (a < b) w = (b - (a & align_mask) + REPL_STRIDE - 1) / REPL_STRIDE;
    }
    out[item] = w;
}

__device__ __forceinline__ u64 eq_byte64(const u64 (&p)[8], uint32_t c)
{
    u64 t = ~0ull;
#pragma unroll
    for (int b = 0; b < 8; ++b) t &= p[b] ^ (((c >> b) & 1u) ? 0ull : ~0ull);
    return t;
}

__device__ __forceinline__ u64 mirror64(u64 x)  // position p of the window -> 2047 - p
{
    return mk64(__shfl_xor_sync(FULL, __brev(hi32(x)), 31), __shfl_xor_sync(FULL, __brev(lo32(x)), 31));
}
// R[p] = A[p] | (R[p+1] & K[p+1]) over the window; `carry_in` = R[2048] & K[2048]
__device__ __forceinline__ u64 rspread64(u64 a, u64 k, bool carry_in, const LaneCtx& L)
{
    const uint32_t cw = carry_in ? 0x80000000u : 0u;
    // mirrored coordinates q = 2047 - p: R'[q] = A'[q] | (R'[q-1] & K'[q-1]), and spread64 wants the K of position q itself
    return mirror64(spread64(mirror64(a), adv64(mirror64(k), cw, L), cw, L));
}

template <int MODE, bool WRITE>
__global__ void __launch_bounds__(THREADS, 3)
k_replace_splice64(const __grid_constant__ ReplArgs A)
{
    extern __shared__ __align__(64) unsigned char repl_dsm[];
    WarpSmRepl* sm = (WarpSmRepl*)repl_dsm;
    __shared__ uint8_t s_repl[REPL_REPL_MAX];
    if (threadIdx.x < REPL_REPL_MAX) s_repl[threadIdx.x] = A.repl[threadIdx.x];
    __syncthreads();
    LaneCtx L;
    L.lane = lane_id();
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    L.m31 = L.lane == 31 ? 1u : 0u;
    const uint32_t lane = L.lane;
    WarpSmRepl& W = sm[threadIdx.x >> 5];
    const uint32_t wb = (uint32_t)__cvta_generic_to_shared(&W);
    const uint32_t my0 = wb + ring_lane_offset(lane);
    const uint32_t rs_base = wb + (uint32_t)offsetof(WarpSmRepl, rs);
    const uint32_t kk_base = wb + (uint32_t)offsetof(WarpSmRepl, kk);
    const uint32_t mm_base = wb + (uint32_t)offsetof(WarpSmRepl, mm);
    const uint32_t pre_base = wb + (uint32_t)offsetof(WarpSmRepl, pre);
    const uint32_t repl_base = (uint32_t)__cvta_generic_to_shared(s_repl);
    const char* gsrc = A.chars + 64 * (int)lane;
    const int m = A.m, rlen = A.rlen;

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.item_counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= A.nitems) break;
        const int ra = __ldg(A.item_bounds + item), rb = __ldg(A.item_bounds + item + 1);
        const int slot0 = __ldg(A.item_slot + item);
        int byte_a = 0, byte_b = 0;
        if (ra < rb) {
            byte_a = __ldg(A.offsets + ra);
            byte_b = __ldg(A.offsets + rb);
        }
        if (ra >= rb || byte_a >= byte_b) {
            // rows without a byte start where the next slot starts
            if (WRITE)
                for (int j = ra + (int)lane; j < rb; j += 32) A.new_off[j] = (int32_t)__ldg(A.slot_base + slot0);
            continue;
        }
        int ws = byte_a & (MODE == 1 ? ~63 : ~15);
        uint32_t carry_sp = 0;  // MODE 1: top word of S at the end of the previous window's owned bytes
        int kown = ra;  // next row whose new offset has not been written (rows ra .. rb-1 start in [byte_a, byte_b])
        int carry = 0;  // leading bytes of the window still covered by the last occurrence of the previous one
        int stage = 0;
        size_t slot = (size_t)slot0;
        __syncwarp();
        ring_issue(my0, gsrc, A.chars, ws, A.end, lane);

        for (; ws < byte_b; ws += REPL_STRIDE, stage ^= 1, ++slot) {
            const int we = ws + REPL_STRIDE;  // end of the owned bytes; the look-ahead lane covers [we, we + 64)
            const bool more = we < byte_b;
            const uint32_t cur0 = my0 + (uint32_t)stage * WIN64;
            if (more) ring_issue(my0 + (uint32_t)(stage ^ 1) * WIN64, gsrc, A.chars, we, A.end, lane);

            // ---- RS bits of every row that starts inside [ws, ws + 2048); the rows this window owns are kfirst .. kown-1
            asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(rs_base + 8u * lane), "r"(0u) : "memory");
            __syncwarp();
            const int kfirst = kown;
            for (int j0 = kown;; j0 += 32) {
                const int j = j0 + (int)lane;
                const int o = j <= A.n ? __ldg(A.offsets + j) : 0x7fffffff;  // offsets[n] too: the end of a row-slice view is a row start of its parent
                const bool inw = o < ws + WIN64;
                if (MODE == 0 && inw) reds_or(rs_base + 4u * (uint32_t)((o - ws) >> 5), 1u << ((o - ws) & 31));
                kown += __popc(__ballot_sync(FULL, j < rb && (o < we || !more)));
                if (__ballot_sync(FULL, inw) != FULL) break;
            }
            __syncwarp();
            const u64 nrs = ~lds64(rs_base + 8u * lane);

            // ---- bytes -> bit planes -> occurrence starts
            if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            u64 p[8];
            const uint4 v0 = lds128(cur0), v1 = lds128(cur0 ^ 16u), v2 = lds128(cur0 ^ 32u), v3 = lds128(cur0 ^ 48u);
            {
                uint32_t pl[8], ph[8];
                transpose_planes(v0, v1, pl);
                transpose_planes(v2, v3, ph);
#pragma unroll
                for (int b = 0; b < 8; ++b) p[b] = mk64(pl[b], ph[b]);
            }
            const int wp = ws + 64 * (int)lane;
            const int own_lo = byte_a > ws ? byte_a : ws, own_hi = byte_b < we ? byte_b : we;
            u64 own = 0;
            if (wp + 64 > own_lo && wp < own_hi) {
                own = ~0ull;
                if (wp < own_lo) own &= ~0ull << (own_lo - wp);
                if (wp + 64 > own_hi) own &= ~0ull >> (wp + 64 - own_hi);
            }
            u64 M, D;
            if (MODE == 0) {
                u64 X = eq_byte64(p, A.pat[m - 1]);
#pragma unroll 1
                for (int k = m - 2; k >= 0; --k) {
                    const u64 y = X & nrs;  // the byte behind must belong to the same row
                    const uint32_t dn = __shfl_down_sync(FULL, lo32(y), 1);
                    const uint32_t hi = __funnelshift_r(hi32(y), lane == 31 ? 0u : dn, 1), lo = __funnelshift_r(lo32(y), hi32(y), 1);
                    X = mk64(lo, hi) & eq_byte64(p, A.pat[k]);
                }
                M = X & own;
                // ---- DROP: M smeared upwards over m bytes (doubling), plus the tail of the previous window's last occurrence
                D = M;
                for (int cover = 1; cover < m;) {
                    const int sh = cover < m - cover ? cover : m - cover;  // <= 16
                    const uint32_t up = __shfl_up_sync(FULL, hi32(D), 1);
                    const uint32_t lo = __funnelshift_l(lane == 0 ? 0u : up, lo32(D), sh), hi = __funnelshift_l(lo32(D), hi32(D), sh);
                    D |= mk64(lo, hi);
                    cover += sh;
                }
                if (lane == 0 && carry) D |= (1ull << carry) - 1ull;
                {
                    const int top = M ? 64 * (int)lane + 63 - __clzll((long long)M) : -1;
                    const int mx = __reduce_max_sync(FULL, top);
                    carry = mx + m - REPL_STRIDE;
                    if (carry < 0) carry = 0;
                }
            } else {
                const long long wi = ((long long)ws + 64 * (int)lane - A.span_base) >> 6;
                u64 Mw = 0, Kw = 0, Aw = 0;
                if (wi < A.span_words) {
                    Mw = __ldg(A.span_m + wi);
                    Kw = __ldg(A.span_k + wi);
                    Aw = __ldg(A.span_a + wi);
                }
                const u64 cont = p[7] & ~p[6];
                // does the K-run that continues behind the window hold an A?  (warp-uniform scan, nearly always one word)
                bool rc = false;
                {
                    long long w = ((long long)ws + WIN64 - A.span_base) >> 6;
                    for (int it = 0; w < A.span_words; ++it, ++w) {
                        const u64 kw = __ldg(A.span_k + w), aw = __ldg(A.span_a + w);
                        const u64 nk = ~kw;
                        const u64 run = nk ? ((nk & (0ull - nk)) - 1ull) : ~0ull;  // the bits of the run inside this word
                        if (aw & run) { rc = true; break; }
                        if (nk) break;
                        if (it == 64) {  // a run of more than 4 KiB behind the window: not worth a serial scan per window
                            if (lane == 0) atomicOr(A.flags, 1u);
                            break;
                        }
                    }
                }
                const u64 S = spread64(Mw, Kw, carry_sp, L);
                const u64 FM = Mw & ~(adv64(S, carry_sp, L) & Kw);
                const u64 R = rspread64(Aw, Kw, rc, L);
                const u64 FMv = FM & R;
                carry_sp = __shfl_sync(FULL, hi32(S), 30);  // the next window starts behind lane 30
                // the k characters in front of every occurrence's last step (lane 31 is the look-ahead for lane 30)
                u64 X = FMv, acc = 0;
#pragma unroll 1
                for (int c = 0; c < A.k_chars; ++c) {
                    X = shift_down64(X, 0u, L);
                    acc |= X;
#pragma unroll 1
                    for (int r = 0; r < 3; ++r) {  // markers that landed on a continuation byte travel on to the lead byte
                        const u64 T = X & cont;
                        if (!__any_sync(FULL, T != 0)) break;
                        X = (X & ~cont) | shift_down64(T, 0u, L);
                        acc |= X;
                    }
                }
                M = FMv & own;
                D = acc | (S & R);
            }
            const u64 K = own & ~D;
            const uint32_t cnt = (uint32_t)__popcll(K) + (uint32_t)rlen * (uint32_t)__popcll(M);
            if (!WRITE) {
                const uint32_t tot = __reduce_add_sync(FULL, cnt);
                if (lane == 0) A.slot_counts[slot] = tot;
                continue;
            }
            // ---- write pass: exclusive prefix of the output bytes over the lanes
            uint32_t pre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, pre, d);
                if ((int)lane >= d) pre += v;
            }
            const uint32_t total = __shfl_sync(FULL, pre, 31);
            pre -= cnt;
            const long long out_a = (long long)__ldg(A.slot_base + slot);
            const uint32_t phase = (uint32_t)(out_a & 15);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(kk_base + 8u * lane), "r"(lo32(K)), "r"(hi32(K)) : "memory");
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(mm_base + 8u * lane), "r"(lo32(M)), "r"(hi32(M)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(pre_base + 4u * lane), "r"(pre) : "memory");
            // my kept bytes -> tile, in position order (straight line; an occurrence start only moves the cursor on by rlen: a
            // window without any occurrence, warp-uniform, runs the copy of the loop that does not test for them), then the
            // replacements: one short loop over my M bits
            {
                const uint32_t tile = wb + (uint32_t)offsetof(WarpSmRepl, tile);
                const uint32_t o0 = phase + pre;
                const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
                uint32_t o = o0;
                if (__any_sync(FULL, M != 0)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t km = (uint32_t)(K >> (4 * i)) & 15u, mm4 = (uint32_t)(M >> (4 * i)) & 15u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (mm4 & (1u << k)) o += (uint32_t)rlen;
                            if (km & (1u << k)) {
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(tile + tile_pad(o)), "r"(w[i] >> (8 * k)) : "memory");
                                ++o;
                            }
                        }
                    }
                    if (rlen > 0) {
                        int nm = 0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t m32 = h ? hi32(M) : lo32(M), k32 = h ? hi32(K) : lo32(K);
                            const uint32_t ob = o0 + (h ? (uint32_t)__popc(lo32(K)) : 0u);
                            for (uint32_t e = m32; e; e &= e - 1, ++nm) {
                                const int b = __ffs((int)e) - 1;
                                const uint32_t at = ob + (uint32_t)__popc(k32 & ((1u << b) - 1u)) + (uint32_t)(nm * rlen);
                                for (int q = 0; q < rlen; ++q)
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(tile + tile_pad(at + (uint32_t)q)), "r"(lds8(repl_base + (uint32_t)q)) : "memory");
                            }
                        }
                    }
                } else {
                    tile_zero(tile, phase + total, lane);
                    scatter_kept(tile, o0, w, K);
                }
            }
            __syncwarp();
            // ---- new offsets of the rows that start in the owned bytes: output bytes before the row's first byte
            for (int k0 = kfirst; k0 < kown; k0 += 32) {
                const int j = k0 + (int)lane;
                if (j < kown) {
                    const int x = __ldg(A.offsets + j) - ws;
                    uint32_t before;
                    if (x >= own_hi - ws) before = total;
                    else {
                        const uint32_t l = (uint32_t)x >> 6, bit = (uint32_t)x & 63u;
                        const u64 below = (1ull << bit) - 1ull;
                        before = lds32(pre_base + 4u * l) + (uint32_t)__popcll(lds64(kk_base + 8u * l) & below) +
                                 (uint32_t)rlen * (uint32_t)__popcll(lds64(mm_base + 8u * l) & below);
                    }
                    A.new_off[j] = (int32_t)(out_a + before);
                }
            }
            flush_tile(wb + (uint32_t)offsetof(WarpSmRepl, tile), A.out, out_a, (int)total, lane);
            __syncwarp();
        }
    }
}
