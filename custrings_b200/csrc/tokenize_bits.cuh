// NVText::tokenize (text/tokens.cu:41-155) as a bit-stream compaction — included by regex_bits.cu inside custr::bits.
//
// For whitespace (byte <= ' ') or a small set of ASCII delimiter bytes the flat token column is a pure function of three
// per-byte bit streams of the chars buffer:
//     D  the byte is a delimiter                       (boolean formula over the bit planes, like a regex class)
//     T  = ~D            the byte belongs to a token   -> output chars = the T bytes, in order (stream compaction)
//     S  = T & (ROWSTART | advance(D))                  the byte starts a token -> one output offset per S bit
// so tokenize needs no per-row walk at all: token t starts at output offset (number of T bits before its S bit).
// Same window machinery as the chain kernel (2048-byte windows, lane = 64 bytes, cp.async ring, ROWSTART bits scattered
// from `offsets`, 32 KiB row-aligned work items handed out dynamically).  Two passes:
//   count : per (window, owner) slot the number of S and T bits                     -> one exclusive scan on the host side
//   write : recompute the streams; a lane's tokens / bytes land at slot base + warp prefix + rank inside its word; bytes
//           are compacted into a shared-memory tile and written back with coalesced 16-byte stores.
// Slots are numbered item-major (item_slot[item] + window index inside the item, from an exclusive scan over the number of
// windows every item touches), which is position order: a window shared by several items gets one slot per owner.
#pragma once

struct TokArgs {
    const char* chars;
    const int32_t* offsets;
    int32_t n, first, end, nitems;
    unsigned int* item_counter;
    const int32_t* item_bounds;
    uint32_t whitespace;      // 1: byte <= 0x20 delimits; 0: the bytes in delims[0..ndelims)
    uint32_t ndelims;
    uint8_t delims[8];
    const int32_t* item_slot; // first slot of every work item
    unsigned long long* slot_counts;        // count pass: (tokens << 32 | bytes) per slot, atomically accumulated
    const unsigned long long* slot_base;    // write pass: exclusive scan of slot_counts
    int32_t* tok_off;         // write pass: output offsets (token t -> first byte)
    char* out;                // write pass: output chars (16-byte aligned)
};

constexpr int TOK_DELIMS_MAX = 8;

// The staging tile is padded by one word per 64 bytes: lanes that each write ~64 output bytes start 17 words apart instead of
// 16, so their byte stores fall into different banks (unpadded, a window that drops nothing is a 16-way conflict per store:
// profiles/r2_ncu_splice.txt).  tile_pad maps a logical tile offset to its padded one; 16-byte aligned logical chunks stay
// contiguous (they never straddle a 64-byte block) but are only word aligned.
__device__ __forceinline__ uint32_t tile_pad(uint32_t o) { return o + ((o >> 6) << 2); }
constexpr int tile_padded_bytes(int logical) { return (logical + (logical / 64 + 1) * 4 + 15) & ~15; }

struct __align__(64) WarpSmTok {
    char ring[RING_STAGES][WIN64];
    uint32_t rs[64];
    char tile[tile_padded_bytes(WIN64 + 32)];  // compacted bytes of the window, mirrored to the 16-byte phase of the output
};

__device__ __forceinline__ u64 delim_stream(const TokArgs& A, const u64 (&p)[8])
{
    u64 d;
    if (A.whitespace) d = ~p[6] & (~p[5] | ~(p[4] | p[3] | p[2] | p[1] | p[0]));  // 0x00..0x20
    else {
        d = 0;
        for (uint32_t k = 0; k < A.ndelims; ++k) d |= cls_eq(p, A.delims[k]);
    }
    return d & ~p[7];
}

// number of windows every work item touches (0 for an item without bytes)
__global__ void k_tok_item_windows(const int32_t* __restrict__ offsets, const int32_t* __restrict__ item_bounds, int nitems, int32_t* __restrict__ out)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const int ra = item_bounds[item], rb = item_bounds[item + 1];
    int w = 0;
    if (ra < rb) {
        const int a = offsets[ra], b = offsets[rb];
        if (a < b) w = ((b - 1) / WIN64) - (a / WIN64) + 1;
    }
    out[item] = w;
}

// Kept bytes (mask K, one bit per byte of my 16 words) -> padded tile from logical offset o0 on, without inserts.
// Default: one predicated st.shared.u8 per input byte, straight-line.  -DCUSTR_SCATTER_WORDS: every word is compacted with one
// PRMT (selector from a 16-entry table indexed by its 4 keep bits), appended to a 4-byte shift register, and whole aligned
// words are OR-ed into the ZEROED tile (red.shared.or.b32; the partial words at either end of a lane's range combine with the
// neighbours' by the OR) — ~17 shared-memory operations per lane and window instead of 64 and ~200 fewer instructions, but
// measured no faster on C2 (tokenize 1.606 vs 1.610 ms, replace 1.32 vs 1.34 ms, split_record 1.77 vs 1.66 ms): the write
// passes are bound by issue efficiency (IPC 0.44 per scheduler), not by the shared-memory pipe, once the tile is padded.
__constant__ uint16_t c_compact_sel[16] = {0x4444, 0x4440, 0x4441, 0x4410, 0x4442, 0x4420, 0x4421, 0x4210,
                                           0x4443, 0x4430, 0x4431, 0x4310, 0x4432, 0x4320, 0x4321, 0x3210};
__device__ __forceinline__ void tile_zero(uint32_t tile, uint32_t logical_bytes, uint32_t lane)
{
#ifdef CUSTR_SCATTER_WORDS
    const uint32_t end = tile_pad(logical_bytes) + 4u;
    for (uint32_t q = 16u * lane; q < end; q += 512u) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(tile + q), "r"(0u) : "memory");
    __syncwarp();
#endif
}
__device__ __forceinline__ void scatter_kept(uint32_t tile, uint32_t o0, const uint32_t (&w)[16], u64 K)
{
#ifndef CUSTR_SCATTER_WORDS
    uint32_t o = o0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t m = (uint32_t)(K >> (4 * i)) & 15u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (m & (1u << k)) {
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(tile + tile_pad(o)), "r"(w[i] >> (8 * k)) : "memory");
                ++o;
            }
        }
    }
#else
    uint32_t f = o0 & 3u, a = o0 & ~3u, lo = 0;  // f pending bytes in lo, next aligned logical word at a
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t m = (uint32_t)(K >> (4 * i)) & 15u;
        const uint32_t c = __byte_perm(w[i], 0u, (uint32_t)c_compact_sel[m]);
        const uint32_t hi = __funnelshift_l(c, 0u, 8u * f);  // c >> (32 - 8 f)
        lo |= c << (8u * f);
        f += (uint32_t)__popc(m);
        if (f >= 4u) {
            reds_or(tile + tile_pad(a), lo);
            lo = hi;
            a += 4u;
            f -= 4u;
        }
    }
    if (f) reds_or(tile + tile_pad(a), lo);
#endif
}

// tile (shared address; its logical byte 0 = output byte out_a & ~15) -> out[out_a, out_a + nbytes): 16-byte stores on the
// aligned interior; the partial chunk at either end goes out byte-wise, one byte per lane (lanes 0-15 the head, 16-31 the tail)
__device__ __forceinline__ void flush_tile(uint32_t tile, char* __restrict__ out, long long out_a, int nbytes, uint32_t lane)
{
    const long long a0 = out_a & ~15ll, oe = out_a + nbytes;
    for (long long q = a0 + 16 * (int)lane; q + 16 <= oe; q += 16 * 32)
        if (q >= out_a) {
            const uint32_t a = tile + tile_pad((uint32_t)(q - a0));
            *(uint4*)(out + q) = make_uint4(lds32(a), lds32(a + 4u), lds32(a + 8u), lds32(a + 12u));
        }
    const long long qt = oe & ~15ll;
    const int i = (int)lane & 15;
    if (lane < 16) {
        const long long r = a0 + i;
        if (a0 < out_a && r >= out_a && r < oe) out[r] = (char)lds8(tile + tile_pad((uint32_t)(r - a0)));
    } else {
        const long long r = qt + i;
        if ((qt > a0 || a0 == out_a) && r < oe) out[r] = (char)lds8(tile + tile_pad((uint32_t)(r - a0)));
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(THREADS, 3)
k_tokenize64(const __grid_constant__ TokArgs A)
{
    extern __shared__ __align__(64) unsigned char tok_dsm[];  // WARPS x WarpSmTok (50.5 KB: above the static limit)
    WarpSmTok* sm = (WarpSmTok*)tok_dsm;
    LaneCtx L;
    L.lane = lane_id();
    asm volatile("" : "+r"(L.lane));
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    L.m31 = L.lane == 31 ? 1u : 0u;
    asm volatile("" : "+r"(L.m31));
    const uint32_t lane = L.lane;
    WarpSmTok& W = sm[threadIdx.x >> 5];
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(&W);
    uint32_t my0 = wb + ring_lane_offset(lane);
    uint32_t my_w = wb + 8u * lane + RING_STAGES * WIN64;  // my word of the ROWSTART stream
    const uint32_t rs_base = wb + RING_STAGES * WIN64;
    const char* gsrc = A.chars + 64 * (int)lane;
    asm volatile("" : "+r"(wb), "+r"(my0), "+r"(my_w));

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.item_counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= A.nitems) break;
        const int ra = __ldg(A.item_bounds + item), rb = __ldg(A.item_bounds + item + 1);
        if (ra >= rb) continue;
        const int byte_a = __ldg(A.offsets + ra), byte_b = __ldg(A.offsets + rb);
        if (byte_a >= byte_b) continue;
        const int slot0 = __ldg(A.item_slot + item);
        uint32_t last_d = 0x80000000u;  // "previous byte is a delimiter": irrelevant at the item start (it is a row start)
        int ws = byte_a & ~(WIN64 - 1);
        int kcur = ra + 1;
        int pend = byte_a - ws;
        int o_nxt = (kcur + (int)lane <= rb) ? __ldg(A.offsets + kcur + (int)lane) : 0x7fffffff;
        int stage = 0;
        __syncwarp();
        ring_issue(my0, gsrc, A.chars, ws, A.end, lane);

        for (; ws < byte_b; ws += WIN64, stage ^= 1) {
            const int we = ws + WIN64;
            const bool more = we < byte_b;
            const uint32_t cur0 = my0 + (uint32_t)stage * WIN64;
            if (more) ring_issue(my0 + (uint32_t)(stage ^ 1) * WIN64, gsrc, A.chars, we, A.end, lane);

            // ---- ROWSTART bits of the offsets that fall into [ws, we)
            sts64(my_w, (lane == 0 && pend == 0) ? 1u : 0u, 0u);
            __syncwarp();
            if (pend > 0 && lane == 0) reds_or(rs_base + 4u * (uint32_t)(pend >> 5), 1u << (pend & 31));
            bool at_we = false;
            int consumed = 0;
            for (;;) {
                const int o = consumed == 0 ? o_nxt : ((kcur + consumed + (int)lane <= rb) ? __ldg(A.offsets + kcur + consumed + (int)lane) : 0x7fffffff);
                const bool inw = o <= we;
                if (inw && o < we) reds_or(rs_base + 4u * (uint32_t)((o - ws) >> 5), 1u << ((o - ws) & 31));
                const unsigned m_in = __ballot_sync(FULL, inw);
                at_we = at_we || __any_sync(FULL, inw && o == we);
                consumed += __popc(m_in);
                if (m_in != FULL) break;
            }
            kcur += consumed;
            o_nxt = (kcur + (int)lane <= rb) ? __ldg(A.offsets + kcur + (int)lane) : 0x7fffffff;
            pend = at_we ? 0 : -1;
            __syncwarp();
            const u64 rs = lds64(my_w);

            // ---- bytes -> bit planes -> delimiter / token / token-start streams
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
            const u64 D = delim_stream(A, p);
            // this work item owns the bytes of [byte_a, byte_b)
            const int wp = ws + 64 * (int)lane;
            u64 own = 0;
            if (wp + 64 > byte_a && wp < byte_b) {
                own = ~0ull;
                if (wp < byte_a) own &= ~0ull << (byte_a - wp);
                if (wp + 64 > byte_b) own &= ~0ull >> (wp + 64 - byte_b);
            }
            const u64 T = ~D & own;
            const u64 S = T & (rs | adv64(D, last_d, L));
            last_d = hi32(D);

            const uint32_t cnt = ((uint32_t)__popcll(S) << 16) | (uint32_t)__popcll(T);  // <= 64 each
            const size_t slot = (size_t)slot0 + (size_t)((ws - (byte_a & ~(WIN64 - 1))) / WIN64);
            if (!WRITE) {
                const uint32_t tot = __reduce_add_sync(FULL, cnt);
                if (lane == 0 && tot) atomicAdd(A.slot_counts + slot, ((unsigned long long)(tot >> 16) << 32) | (tot & 0xffffu));
                continue;
            }
            // ---- write pass: exclusive prefix of (tokens, bytes) over the lanes
            uint32_t pre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, pre, d);
                if ((int)lane >= d) pre += v;
            }
            const uint32_t total = __shfl_sync(FULL, pre, 31);
            pre -= cnt;
            const unsigned long long base = __ldg(A.slot_base + slot);
            const long long out_a = (long long)(base & 0xffffffffull);  // first output byte of this (window, owner)
            const int tok_a = (int)(base >> 32);
            const int nbytes = (int)(total & 0xffffu);
            const uint32_t phase = (uint32_t)(out_a & 15);
            // tokens of my word.  Their offsets are staged in the ring stage this window came from (its bytes now live in
            // registers) and written back coalesced; a window with more tokens than the stage holds stores directly.
            __syncwarp();  // every lane has read its chunk of this ring stage: it may now be overwritten with token offsets
            const int ntoks = (int)(total >> 16);
            const bool stage_toks = ntoks <= WIN64 / 4;
            const uint32_t tokbuf = wb + (uint32_t)stage * WIN64;
            {   // 32-bit halves: FLO / POPC / shifts on 64-bit values cost 2-4 instructions each
                int t = (int)(pre >> 16);
                int32_t off0 = (int32_t)(out_a + (pre & 0xffffu));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t s32 = h ? hi32(S) : lo32(S);
                    const uint32_t t32 = h ? hi32(T) : lo32(T);
                    while (s32) {
                        const int b = __ffs((int)s32) - 1;
                        s32 &= s32 - 1;
                        const int32_t off = off0 + __popc(t32 & ((1u << b) - 1u));
                        if (stage_toks) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tokbuf + 4u * (uint32_t)t), "r"(off) : "memory");
                        else A.tok_off[tok_a + t] = off;
                        ++t;
                    }
                    off0 += __popc(t32);
                }
            }
            // bytes of my word -> tile (the words are still in registers)
            {
                const uint32_t tile = wb + (uint32_t)offsetof(WarpSmTok, tile);
                const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
                tile_zero(tile, phase + (uint32_t)nbytes, lane);
                scatter_kept(tile, phase + (pre & 0xffffu), w, T);
            }
            __syncwarp();
            if (stage_toks)
                for (int i = (int)lane; i < ntoks; i += 32) A.tok_off[tok_a + i] = (int32_t)lds32(tokbuf + 4u * (uint32_t)i);
            flush_tile(wb + (uint32_t)offsetof(WarpSmTok, tile), A.out, out_a, nbytes, lane);
            __syncwarp();
        }
    }
}
