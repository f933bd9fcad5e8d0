// Warp-cooperative staging for "rewrite every row" kernels (replace_re / replace write pass, tokenize, split_record).
//
// A thread-per-row kernel that copies bytes straight from global to global issues, per warp instruction, 32 accesses to 32
// different rows: ~32 sectors per request, and the copy pass ends up 10-20x away from the HBM roofline.  But the chars of a
// block of 32 consecutive rows are ONE contiguous byte range, and so is everything those rows produce (row-major outputs).
// So a warp (1) copies the block's input range into shared memory with coalesced 16-byte loads, (2) lets every lane rewrite
// its own row from shared memory into a shared output buffer, (3) writes the output range back with coalesced 16-byte
// stores.  Blocks whose input or output range does not fit (very long rows) take the direct global path, lane per row.
//
// Shared buffers mirror the 16-byte phase of the global ranges (sm[(g & 15) + k] <-> global[g + k]) so both bulk copies are
// aligned vector accesses.
#pragma once
#include "common.cuh"

namespace custr {
namespace stage {

constexpr int ROWS = 32;     // rows per warp block
constexpr int CAP = 4992;    // staged bytes per warp and direction (C2: 32 rows x 107 B = 3.4 KB average)
constexpr int WARPS = 4;     // warps per CTA: 4 x (2 x (CAP + 32) + bit streams) = 47.0 KB static shared memory
constexpr int STREAM_WORDS = (CAP + 16) / 64 + 2;  // 64-bit words of one staged bit stream
constexpr int THREADS = WARPS * 32;

struct __align__(16) Buffers {
    char in[WARPS][CAP + 32];
    char out[WARPS][CAP + 32];
};
struct __align__(16) StreamBuffers {  // optional: three per-byte bit streams of the staged input range (span_walk.cuh)
    unsigned long long w[WARPS][3][STREAM_WORDS];
};

// global [gb, ge) -> sm[(gb & 15) + k]; `g` must be 16-byte aligned, [0, g_limit) readable
__device__ __forceinline__ void copy_in(char* sm, const char* __restrict__ g, int gb, int ge, int g_limit, int lane)
{
    const int a0 = gb & ~15;
    for (int p = a0 + 16 * lane; p < ge; p += 16 * 32) {
        if (p + 16 <= g_limit) *(uint4*)(sm + (p - a0)) = __ldg((const uint4*)(g + p));
        else
            for (int q = p; q < ge; ++q) sm[q - a0] = g[q];
    }
}
// sm[(gb & 15) + k] -> global [gb, ge); `g` must be 16-byte aligned
__device__ __forceinline__ void copy_out(const char* sm, char* __restrict__ g, long long gb, long long ge, int lane)
{
    const long long a0 = gb & ~15ll;
    for (long long p = a0 + 16 * lane; p < ge; p += 16 * 32) {
        if (p >= gb && p + 16 <= ge) *(uint4*)(g + p) = *(const uint4*)(sm + (p - a0));
        else {
            const long long lo = p < gb ? gb : p, hi = p + 16 < ge ? p + 16 : ge;
            for (long long q = lo; q < hi; ++q) g[q] = sm[q - a0];
        }
    }
}

// Runs body(i, src, dst, staged, in_a) for every row i of the column, where src[p] is the input byte at chars offset p
// (valid for the row's own range) and dst[o] the output byte at output offset o (valid for the row's own output range);
// staged tells whether they are the shared copies, in_a is the first chars offset of the row's block.
// out_begin(r) = output offset at which row r's output starts (r in [0, n]; monotone).
// pre(in_a, in_b, warp, lane) runs before the rows of a staged block (e.g. to stage more data of the same byte range).
template <typename OutBegin, typename Pre, typename Body>
__device__ __forceinline__ void for_each_row_staged(const ColView& col, int chars_limit, char* __restrict__ out_chars, Buffers& sm,
                                                    OutBegin out_begin, Pre pre, Body body)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblocks = (col.n + ROWS - 1) / ROWS;
    for (int blk = blockIdx.x * WARPS + warp; blk < nblocks; blk += gridDim.x * WARPS) {
        const int r0 = blk * ROWS, r1 = r0 + ROWS < col.n ? r0 + ROWS : col.n;
        const int in_a = col.offsets[r0], in_b = col.offsets[r1];
        const long long out_a = out_begin(r0), out_b = out_begin(r1);
        const int i = r0 + lane;
        const bool fits = (in_b - (in_a & ~15)) <= CAP + 16 && (out_b - (out_a & ~15ll)) <= CAP + 16;
        if (!fits) {  // direct path
            if (i < r1) body(i, (const uint8_t*)col.chars, out_chars, false, in_a);
            continue;
        }
        copy_in(sm.in[warp], col.chars, in_a, in_b, chars_limit, lane);
        pre(in_a, in_b, warp, lane);
        __syncwarp();
        if (i < r1) body(i, (const uint8_t*)sm.in[warp] - (in_a & ~15), sm.out[warp] - (out_a & ~15ll), true, in_a);
        __syncwarp();
        copy_out(sm.out[warp], out_chars, out_a, out_b, lane);
        __syncwarp();
    }
}

inline int grid_for(int n)
{
    int want = ((n + ROWS - 1) / ROWS + WARPS - 1) / WARPS;
    int cap = num_sms() * 4;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace stage
}  // namespace custr
