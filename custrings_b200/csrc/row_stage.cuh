// Warp-cooperative staging helpers for "rewrite every row" kernels (the replace_re splice, regex.cu).
//
// The chars of a block of 32 consecutive rows are ONE contiguous byte range, and so is everything those rows produce
// (row-major outputs): a warp copies the input range into shared memory with coalesced 16-byte loads, rewrites it there,
// and writes the output range back with coalesced 16-byte stores.  Shared buffers mirror the 16-byte phase of the global
// ranges (sm[(g & 15) + k] <-> global[g + k]) so both bulk copies are aligned vector accesses.
#pragma once
#include "common.cuh"

namespace custr {
namespace stage {

constexpr int ROWS = 32;     // rows per warp block
constexpr int WARPS = 4;     // warps per CTA
constexpr int THREADS = WARPS * 32;

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

inline int grid_for(int n)
{
    int want = ((n + ROWS - 1) / ROWS + WARPS - 1) / WARPS;
    int cap = num_sms() * 4;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace stage
}  // namespace custr
