// NVStrings::split_record (strings/split.cu:125-223) with a single ASCII delimiter byte and no split limit, as a bit-stream
// compaction — included by regex_bits.cu inside custr::bits, after tokenize_bits.cuh whose window machinery it shares.
//
// Per-row semantics (custring_view.inl:1169-1279): a valid row yields 1 + (number of delimiter bytes) tokens, empty ones
// included ("a,,b" -> "a", "", "b"; "" -> ""), a null row yields none.  In stream terms, with D = delimiter bytes and
// RSV = first byte of every valid row:
//     output chars   = the non-delimiter bytes T = ~D in order (stream compaction, exactly tokenize's)
//     token offsets  = one entry per EVENT in position order — a RSV bit (the row's first token) or a D bit (the token behind
//                      it) — whose value is the number of T bytes before that position; a row that starts with the delimiter
//                      has both events on one byte, the row start first (same value: its first token is empty)
//     row_offsets[r] = number of events before the start of row r
// The one thing bits cannot express is several VALID rows starting on the same byte, i.e. an empty valid row: the count pass
// flags it and the caller takes the per-row path for that column.
// Two passes like tokenize: count (events, T bytes) per (item, window) slot -> exclusive scan -> write.
#pragma once

struct SplitArgs {
    const char* chars;
    const int32_t* offsets;
    const uint8_t* validity;  // null: every row valid
    int32_t vbit0;
    int32_t n, first, end, nitems;
    unsigned int* item_counter;
    const int32_t* item_bounds;
    uint32_t delim;
    const int32_t* item_slot;
    unsigned long long* slot_counts;      // count pass: (events << 32 | bytes) per slot
    const unsigned long long* slot_base;  // write pass: exclusive scan of slot_counts
    unsigned int* flags;                  // count pass: bit 0 = an empty valid row exists (not expressible: caller falls back)
    int32_t* tok_off;                     // write pass outputs
    int32_t* row_off;
    char* out;
};

struct __align__(64) WarpSmSplit {
    char ring[RING_STAGES][WIN64];
    uint32_t rs[64];    // ROWSTART of valid rows (RSV), one bit per byte of the window
    uint32_t dd[64];    // D & own: the delimiter bytes this item owns
    uint32_t pre[32];   // exclusive prefix over the lanes of (events << 16 | T bytes)
    char tile[tile_padded_bytes(WIN64 + 32)];
};

__device__ __forceinline__ bool split_row_valid(const SplitArgs& A, int row)
{
    if (!A.validity) return true;
    const int b = A.vbit0 + row;
    return (A.validity[b >> 3] >> (b & 7)) & 1;
}

template <bool WRITE>
__global__ void __launch_bounds__(THREADS, 3)
k_split_record64(const __grid_constant__ SplitArgs A)
{
    extern __shared__ __align__(64) unsigned char split_dsm[];
    WarpSmSplit* sm = (WarpSmSplit*)split_dsm;
    LaneCtx L;
    L.lane = lane_id();
    asm volatile("" : "+r"(L.lane));
    L.src = (L.lane + 31) & 31;
    L.is31 = L.lane == 31;
    L.m31 = L.lane == 31 ? 1u : 0u;
    const uint32_t lane = L.lane;
    WarpSmSplit& W = sm[threadIdx.x >> 5];
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(&W);
    uint32_t my0 = wb + ring_lane_offset(lane);
    const uint32_t rs_base = wb + (uint32_t)offsetof(WarpSmSplit, rs);
    const uint32_t dd_base = wb + (uint32_t)offsetof(WarpSmSplit, dd);
    const uint32_t pre_base = wb + (uint32_t)offsetof(WarpSmSplit, pre);
    const char* gsrc = A.chars + 64 * (int)lane;

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
            // rows without a byte: valid ones are empty rows (flagged: the caller falls back), null ones have no token and
            // start where the next slot starts
            if (ra < rb) {
                for (int j0 = ra; j0 < rb; j0 += 32) {
                    const int j = j0 + (int)lane;
                    if (j < rb) {
                        if (!WRITE) { if (split_row_valid(A, j)) atomicOr(A.flags, 1u); }
                        else A.row_off[j] = (int)(__ldg(A.slot_base + slot0) >> 32);
                    }
                }
            }
            continue;
        }
        int ws = byte_a & ~(WIN64 - 1);
        const int ws0 = ws;
        int kcur = ra;  // next row whose start has not been placed (rows ra .. rb-1 start in [byte_a, byte_b])
        int stage = 0;
        __syncwarp();
        ring_issue(my0, gsrc, A.chars, ws, A.end, lane);

        for (; ws < byte_b; ws += WIN64, stage ^= 1) {
            const int we = ws + WIN64;
            const bool more = we < byte_b;
            const bool last = !more;
            const uint32_t cur0 = my0 + (uint32_t)stage * WIN64;
            if (more) ring_issue(my0 + (uint32_t)(stage ^ 1) * WIN64, gsrc, A.chars, we, A.end, lane);

            // ---- RSV bits: rows of this item that start inside [ws, we) (the last window also takes the rows that start at
            //      its very end: trailing rows without bytes); kfirst = first such row
            asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(rs_base + 8u * lane), "r"(0u) : "memory");
            __syncwarp();
            const int kfirst = kcur;
            for (;;) {
                const int j = kcur + (int)lane;
                const int o = j < rb ? __ldg(A.offsets + j) : 0x7fffffff;
                const bool inw = j < rb && (o < we || (last && o <= we));
                if (inw) {
                    const bool valid = split_row_valid(A, j);
                    if (valid && o < we) reds_or(rs_base + 4u * (uint32_t)((o - ws) >> 5), 1u << ((o - ws) & 31));
                    if (!WRITE && valid && __ldg(A.offsets + j + 1) == o) atomicOr(A.flags, 1u);  // empty valid row
                }
                const unsigned m_in = __ballot_sync(FULL, inw);
                kcur += __popc(m_in);
                if (m_in != FULL) break;
            }
            __syncwarp();
            const u64 rsv = lds64(rs_base + 8u * lane);

            // ---- bytes -> bit planes -> delimiter stream
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
            u64 own = 0;  // this work item owns the bytes of [byte_a, byte_b)
            if (wp + 64 > byte_a && wp < byte_b) {
                own = ~0ull;
                if (wp < byte_a) own &= ~0ull << (byte_a - wp);
                if (wp + 64 > byte_b) own &= ~0ull >> (wp + 64 - byte_b);
            }
            const u64 D = cls_eq(p, A.delim) & ~p[7] & own;
            const u64 T = ~D & own;
            const uint32_t cnt = ((uint32_t)(__popcll(D) + __popcll(rsv)) << 16) | (uint32_t)__popcll(T);  // events <= 128, bytes <= 64
            const size_t slot = (size_t)slot0 + (size_t)((ws - ws0) / WIN64);
            if (!WRITE) {
                const uint32_t tot = __reduce_add_sync(FULL, cnt);
                if (lane == 0 && tot) atomicAdd(A.slot_counts + slot, ((unsigned long long)(tot >> 16) << 32) | (tot & 0xffffu));
                continue;
            }
            // ---- write pass: exclusive prefix of (events, bytes) over the lanes
            uint32_t pre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, pre, d);
                if ((int)lane >= d) pre += v;
            }
            const uint32_t total = __shfl_sync(FULL, pre, 31);
            pre -= cnt;
            const unsigned long long base = __ldg(A.slot_base + slot);
            const long long out_a = (long long)(base & 0xffffffffull);
            const int tok_a = (int)(base >> 32);
            const int nbytes = (int)(total & 0xffffu);
            const uint32_t phase = (uint32_t)(out_a & 15);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dd_base + 8u * lane), "r"(lo32(D)), "r"(hi32(D)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(pre_base + 4u * lane), "r"(pre) : "memory");
            // events of my word, in position order; a byte that is both a row start and a delimiter gives two entries, the row
            // start first.  Values are staged as 16-bit distances from out_a in the ring stage this window came from (its bytes
            // now live in registers) and written back coalesced; a window with more events than the stage holds stores directly.
            __syncwarp();  // every lane has read its chunk of this ring stage
            const int nev = (int)(total >> 16);
            const bool staged = nev <= WIN64 / 2 - 4;
            const uint32_t evbuf = wb + (uint32_t)stage * WIN64 + 2u * ((uint32_t)tok_a & 3u);  // entry 0 = token tok_a
            {
                int rank0 = (int)(pre >> 16);
                int val0 = (int)(pre & 0xffffu);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d32 = h ? hi32(D) : lo32(D), r32 = h ? hi32(rsv) : lo32(rsv), o32 = h ? hi32(own) : lo32(own);
                    // T bytes below bit b of this half = (b - first owned bit) - delimiters below b: own is one run of bits
                    const int c0 = val0 - (o32 ? __ffs((int)o32) - 1 : 0);
                    int nd = 0;
                    for (uint32_t e = d32; e; e &= e - 1, ++nd) {  // the token behind each delimiter
                        const int b = __ffs((int)e) - 1;
                        const int rank = rank0 + nd + __popc(r32 & ((2u << b) - 1u));
                        const int val = c0 + b - nd;
                        if (staged) asm volatile("st.shared.u16 [%0], %1;" ::"r"(evbuf + 2u * (uint32_t)rank), "h"((unsigned short)val) : "memory");
                        else A.tok_off[tok_a + rank] = (int32_t)out_a + val;
                    }
                    int nr = 0;
                    for (uint32_t e = r32; e; e &= e - 1, ++nr) {  // the first token of each row (few)
                        const int b = __ffs((int)e) - 1;
                        const int ndb = __popc(d32 & ((1u << b) - 1u));
                        const int rank = rank0 + nr + ndb;
                        const int val = c0 + b - ndb;
                        if (staged) asm volatile("st.shared.u16 [%0], %1;" ::"r"(evbuf + 2u * (uint32_t)rank), "h"((unsigned short)val) : "memory");
                        else A.tok_off[tok_a + rank] = (int32_t)out_a + val;
                    }
                    rank0 += nd + nr;
                    val0 += __popc(o32) - nd;
                }
            }
            // bytes of my word -> tile (as in tokenize)
            {
                const uint32_t tile = wb + (uint32_t)offsetof(WarpSmSplit, tile);
                const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
                tile_zero(tile, phase + (uint32_t)nbytes, lane);
                scatter_kept(tile, phase + (pre & 0xffffu), w, T);
            }
            __syncwarp();
            // ---- row_offsets of the rows that start in this window: events before the row's first byte
            for (int k0 = kfirst; k0 < kcur; k0 += 32) {
                const int j = k0 + (int)lane;
                if (j < kcur) {
                    const int x = __ldg(A.offsets + j) - ws;  // 0 .. 2048
                    int before;
                    if (x >= WIN64) before = (int)(total >> 16);
                    else {
                        const uint32_t l = (uint32_t)x >> 6, bit = (uint32_t)x & 63u;
                        const u64 below = (1ull << bit) - 1ull;
                        before = (int)(lds32(pre_base + 4u * l) >> 16) + __popcll(lds64(dd_base + 8u * l) & below) + __popcll(lds64(rs_base + 8u * l) & below);
                    }
                    A.row_off[j] = tok_a + before;
                }
            }
            if (staged) {  // staged events -> token offsets, 16-byte stores on the aligned interior
                const int a0 = tok_a & ~3, te = tok_a + nev;
                const uint32_t buf0 = wb + (uint32_t)stage * WIN64;
                const int32_t add = (int32_t)out_a;
                for (int q = a0 + 4 * (int)lane; q < te; q += 128) {
                    const u64 v = lds64(buf0 + 2u * (uint32_t)(q - a0));
                    const int4 o = make_int4(add + (int)(lo32(v) & 0xffffu), add + (int)(lo32(v) >> 16), add + (int)(hi32(v) & 0xffffu), add + (int)(hi32(v) >> 16));
                    if (q >= tok_a && q + 4 <= te) *(int4*)(A.tok_off + q) = o;
                    else {
                        if (q >= tok_a && q < te) A.tok_off[q] = o.x;
                        if (q + 1 >= tok_a && q + 1 < te) A.tok_off[q + 1] = o.y;
                        if (q + 2 >= tok_a && q + 2 < te) A.tok_off[q + 2] = o.z;
                        if (q + 3 >= tok_a && q + 3 < te) A.tok_off[q + 3] = o.w;
                    }
                }
            }
            flush_tile(wb + (uint32_t)offsetof(WarpSmSplit, tile), A.out, out_a, nbytes, lane);
            __syncwarp();
        }
    }
}
