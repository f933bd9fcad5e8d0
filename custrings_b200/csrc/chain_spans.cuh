// Per-row match SPANS for linear-chain patterns whose only loop is the last step (x y z+ with optional assertions in
// front / at the end): count_re, replace_re, findall without the Pike VM.
//
// Why this is exact (reference semantics regexec.inl:204-442): threads are seeded left to right and an earlier seed
// outranks a later one, and a higher-priority thread is never cut by a lower-priority END — so the reported match starts at
// the LEFTMOST position from which any match exists.  With a fixed-length prefix and at most one greedy loop at the very
// end there is a single NFA path per (start, end) pair, the loop-continue branch outranks the exit branch, and every END
// reached later overrides the earlier one: the reported end is the LONGEST end at which the trailing assertion holds.
// Patterns with a loop before the last step, alternations, optionals or lazy quantifiers never come here.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"
#include "regex_bits_plan.h"

namespace custr {
namespace spans {

using bits::ChainClassD;
using bits::ChainDev;

CUSTR_HD bool class_has_char(const ChainClassD& cc, uint32_t ch, const uint8_t* __restrict__ uflags)
{
    if (ch < 128u) return (cc.ascii[ch >> 5] >> (ch & 31)) & 1u;
    bool in = false;
    switch (cc.na_kind) {
    case bits::NA_ALWAYS: in = true; break;
    case bits::NA_CHAR_EQ: in = ch == cc.na_arg; break;
    case bits::NA_CLASS: case bits::NA_NCLASS: {
        bool m = false;
        for (uint32_t i = 0; i < cc.na_nranges; i += 2)
            if (ch >= cc.na_ranges[i] && ch <= cc.na_ranges[i + 1]) m = true;
        const uint32_t b = cc.na_builtins;
        if (!m && b) {
            const uint32_t cp = packed_to_cp(ch);
            if (cp <= 0xFFFFu) {
                const uint32_t f = CUSTR_LDG(uflags + cp);
                const bool alnum = (f & 15u) != 0, space = (f & 16u) != 0, digit = (f & 4u) != 0;
                m = ((b & 1) && alnum) || ((b & 2) && space) || ((b & 4) && digit) || ((b & 8) && !alnum) || ((b & 16) && !space) ||
                    ((b & 32) && !digit);
            }
        }
        in = (cc.na_kind == bits::NA_CLASS) ? m : !m;
        break;
    }
    default: break;
    }
    return in;
}

// zero-width assertions between `prev` (character before byte offset q, 0 at the start of the row) and `cur` (character
// at q, 0 at the end of the row); the same predicate serves in front of the chain and before END (regexec.inl:308-352)
CUSTR_HD bool assert_chars(uint32_t mask, uint32_t prev, uint32_t cur, int q, int n, const uint8_t* uflags)
{
    if (!mask) return true;
    if (mask & (bits::AS_BOW | bits::AS_NBOW)) {
        const bool bow = is_alnum_packed(cur, uflags) != is_alnum_packed(prev, uflags);
        if ((mask & bits::AS_BOW) && !bow) return false;
        if ((mask & bits::AS_NBOW) && bow) return false;
    }
    if ((mask & bits::AS_BOL_CARET) && !(q == 0 || prev == '\n')) return false;
    if ((mask & bits::AS_BOL_A) && q != 0) return false;
    if ((mask & bits::AS_EOL_DOLLAR) && !(q >= n || cur == '\n')) return false;
    if ((mask & bits::AS_EOL_Z) && q < n) return false;
    return true;
}

// leftmost-longest search from byte offset `begin` in s[0..n); requires that only the last step may loop (greedy) and
// that the row holds no NUL byte.  Every character is decoded once per visit; the previous character is carried along.
CUSTR_HD int chain_find(const ChainDev& cd, const uint8_t* __restrict__ s, int n, int begin, const uint8_t* __restrict__ uflags,
                        int& mbeg, int& mend)
{
    const int ns = (int)cd.nsteps;
    const bool loop_last = cd.steps[ns - 1].loop != 0;
    const uint32_t before = cd.steps[0].before, after = cd.end_mask;
    const ChainClassD& first = cd.classes[cd.steps[0].cls];
    uint32_t prev = (begin > 0 && begin <= n) ? utf8_packed_before(s + begin, s) : 0u;
    for (int pos = begin; pos < n;) {
        int w0;
        const uint32_t c0 = utf8_packed(s + pos, s + n, w0);
        if (class_has_char(first, c0, uflags) && assert_chars(before, prev, c0, pos, n, uflags)) {
            int q = pos + w0;
            uint32_t last = c0;  // last consumed character
            bool ok = true;
            for (int st = 1; st < ns; ++st) {
                if (q >= n) { ok = false; break; }
                int w;
                const uint32_t c = utf8_packed(s + q, s + n, w);
                if (!class_has_char(cd.classes[cd.steps[st].cls], c, uflags)) { ok = false; break; }
                q += w;
                last = c;
            }
            if (ok) {
                int w = 1;
                uint32_t nxt = q < n ? utf8_packed(s + q, s + n, w) : 0u;
                int best = assert_chars(after, last, nxt, q, n, uflags) ? q : -1;
                if (loop_last) {
                    const ChainClassD& lc = cd.classes[cd.steps[ns - 1].cls];
                    while (q < n && class_has_char(lc, nxt, uflags)) {
                        q += w;
                        last = nxt;
                        nxt = q < n ? utf8_packed(s + q, s + n, w) : 0u;
                        if (assert_chars(after, last, nxt, q, n, uflags)) best = q;
                    }
                }
                if (best >= 0) { mbeg = pos; mend = best; return 1; }
            }
        }
        pos += w0;
        prev = c0;
    }
    return 0;
}

// A NUL byte ends the reference's scan in some situations and not in others (it depends on the start-character skip,
// regexec.inl:217-246,434-440).  Rows holding one are not handled here: the kernels raise a flag and the whole call is
// redone by the exact Pike VM.
CUSTR_HD bool has_nul(const uint8_t* s, int n)
{
    for (int i = 0; i < n; ++i)
        if (s[i] == 0) return true;
    return false;
}

// count.cu:168-196 with chain_find as the engine (matches are never empty here)
CUSTR_HD int row_count(const ChainDev& cd, const uint8_t* s, int n, const uint8_t* uflags)
{
    int found = 0, begin = 0, mb, me;
    while (begin <= n && chain_find(cd, s, n, begin, uflags, mb, me)) { ++found; begin = me; }
    return found;
}

// replace.cu:39-107 with chain_find as the engine
CUSTR_HD int row_replace(const ChainDev& cd, const uint8_t* s, int n, const uint8_t* uflags, const char* repl, int repl_len, int maxrepl,
                         char* o)
{
    int budget = maxrepl < 0 ? utf8_count_chars(s, n) : maxrepl;
    int total = n, last = 0, begin = 0, mb, me;
    while (budget > 0 && chain_find(cd, s, n, begin, uflags, mb, me)) {
        total += repl_len - (me - mb);
        if (o) {
            for (int k = last; k < mb; ++k) *o++ = (char)s[k];
            for (int k = 0; k < repl_len; ++k) *o++ = repl[k];
            last = me;
        }
        begin = me;
        --budget;
    }
    if (o) for (int k = last; k < n; ++k) *o++ = (char)s[k];
    return total;
}

}  // namespace spans
}  // namespace custr
