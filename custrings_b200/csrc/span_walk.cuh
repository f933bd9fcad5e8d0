// Exact match spans of last-loop chains (x y z+ with optional assertions in front / at the end) from the three bit
// streams the chain kernel leaves in device memory (regex_bits.cu, Args::span_*), one bit per byte of the column:
//   M  the byte is the lead byte of a character at which a match can BEGIN ITS LAST STEP (the k characters before it matched
//      the first k steps, the leading assertion held at the start)
//   K  the match may continue INTO this byte (last step's class — the loop — or a continuation byte of its character)
//   A  a match may END after this byte (last byte of a character, trailing assertion holds behind it)
// The Pike VM's answer for this family is "leftmost start, then the longest end" (see chain_spans.cuh for the argument), and
// count_re / replace_re resume the search at the end of the previous match (count.cu:174-195, replace.cu:50-106).  In stream
// terms: take the first M bit m at or after the cursor, follow the run of K bits behind it, take the last A bit p of
// [m, run end]; the match is [m - k characters, p + 1) and the cursor becomes (p + 1) + k characters.  Each match costs a few
// word scans instead of a character-by-character NFA walk, and rows without any match are skipped by the chain kernel's
// boolean result.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"

namespace custr {
namespace spans {

using u64 = unsigned long long;

struct Streams {
    const u64* m;
    const u64* k;
    const u64* a;
    int base;  // byte offset of bit 0
};

template <bool LDG>
__device__ __forceinline__ u64 word_at(const u64* s, int w)
{
    return LDG ? __ldg(s + w) : s[w];  // LDG = false: the streams are a shared-memory copy (row_stage.cuh)
}

// first set bit of s in [from, lim), -1 when there is none (positions are byte offsets into chars)
template <bool LDG>
__device__ __forceinline__ int next_set(const u64* s, int base, int from, int lim)
{
    if (from >= lim) return -1;
    const int r = from - base, rl = lim - base;
    int w = r >> 6;
    u64 v = word_at<LDG>(s, w) & (~0ull << (r & 63));
    for (;;) {
        if (v) {
            const int p = (w << 6) + __ffsll((long long)v) - 1;
            return p < rl ? p + base : -1;
        }
        ++w;
        if ((w << 6) >= rl) return -1;
        v = word_at<LDG>(s, w);
    }
}
// first CLEAR bit of s in [from, lim), lim when there is none
template <bool LDG>
__device__ __forceinline__ int next_clear(const u64* s, int base, int from, int lim)
{
    if (from >= lim) return lim;
    const int r = from - base, rl = lim - base;
    int w = r >> 6;
    u64 v = ~word_at<LDG>(s, w) & (~0ull << (r & 63));
    for (;;) {
        if (v) {
            const int p = (w << 6) + __ffsll((long long)v) - 1;
            return p < rl ? p + base : lim;
        }
        ++w;
        if ((w << 6) >= rl) return lim;
        v = ~word_at<LDG>(s, w);
    }
}
// last set bit of s in [lo, hi] (both inclusive), -1 when there is none
template <bool LDG>
__device__ __forceinline__ int last_set(const u64* s, int base, int lo, int hi)
{
    if (hi < lo) return -1;
    const int rlo = lo - base, rhi = hi - base;
    int w = rhi >> 6;
    u64 v = word_at<LDG>(s, w) & (~0ull >> (63 - (rhi & 63)));
    for (;;) {
        if (v) {
            const int p = (w << 6) + 63 - __clzll((long long)v);
            return p >= rlo ? p + base : -1;
        }
        --w;
        if ((w << 6) + 63 < rlo) return -1;
        v = word_at<LDG>(s, w);
    }
}

// calls emit(begin, end) (byte offsets into chars) for the first `budget` matches of the row [a, b); returns their number
// (`chars` only has to be valid for [a, b); LDG = false: streams and chars are shared-memory copies)
template <bool LDG = true, typename F>
__device__ __forceinline__ int walk_spans(const Streams& S, const uint8_t* chars, int a, int b, int k_chars, int budget, F emit)
{
    int found = 0;
    int cm = a;  // smallest admissible M position
    while (found < budget) {
        const int m = next_set<LDG>(S.m, S.base, cm, b);
        if (m < 0) break;
        const int run_end = next_clear<LDG>(S.k, S.base, m + 1, b) - 1;
        const int p = last_set<LDG>(S.a, S.base, m, run_end);
        if (p < 0) {  // no admissible end from this start
            cm = m + 1;
            continue;
        }
        // start of the match: k characters before m.  Fast path: the k bytes before m are loaded independently (not as a
        // chain of dependent loads) and none is a continuation byte, i.e. each of them is one character.
        int s = m - k_chars;
        bool fast = k_chars <= 4 && s >= a;
        if (fast) {
            uint32_t any_cont = 0;
#pragma unroll
            for (int j = 1; j <= 4; ++j)
                if (j <= k_chars) any_cont |= (uint32_t)((chars[m - j] & 0xC0u) == 0x80u);
            fast = !any_cont;
        }
        if (!fast) {
            s = m;
            for (int j = 0; j < k_chars; ++j) {
                --s;
                while (s > a && (chars[s] & 0xC0u) == 0x80u) --s;
            }
        }
        emit(s, p + 1);
        ++found;
        cm = p + 1;  // the next match starts at or after the end of this one: its last step begins k characters later
        fast = k_chars <= 4 && cm + k_chars <= b;
        if (fast) {  // the next k bytes are all ASCII: k characters
            uint32_t hi = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < k_chars) hi |= chars[cm + j];
            fast = hi < 0x80u;
        }
        if (fast) cm += k_chars;
        else
            for (int j = 0; j < k_chars && cm < b; ++j) cm += utf8_width(chars[cm]);
    }
    return found;
}

}  // namespace spans
}  // namespace custr
