// Per-row literal operations as __host__ __device__ functions: find / rfind / replace / split / tokenize.
// The CUDA kernels in find.cu and split.cu call these per row; tests/sim runs the same code on the host.
//
// Semantics follow the reference's per-string device library (cpp/src/custring_view.inl) over raw bytes:
//   find      :481-514   naive compare at every byte offset of [byte(pos), byte(end)-len]; result = char index
//   rfind     :550-582
//   split     :1169-1279 (+ split.cu:768-806 column pick, :892-941 whitespace walk)
//   tokenize  text/tokens.cu:41-121
// All positions inside these helpers are BYTE offsets; character indices are only materialised where the API
// returns them.  For valid UTF-8 this is equivalent to the reference's char-index arithmetic (a delimiter match
// always starts on a character boundary), including its one quirk: when COUNTING delimiters the reference
// resumes `bytes` CHARACTERS after a hit (custring_view.inl:1243), not `bytes` bytes.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"

namespace custr {
namespace row {

CUSTR_HD bool bytes_equal(const uint8_t* a, const uint8_t* b, int m)
{
    for (int j = 0; j < m; ++j)
        if (a[j] != b[j]) return false;
    return true;
}

// first occurrence of t[0..m) in s[from..to) (m > 0); byte offset or -1
CUSTR_HD int find_bytes(const uint8_t* s, int from, int to, const uint8_t* t, int m)
{
    const uint8_t t0 = t[0];
    for (int i = from; i + m <= to; ++i)
        if (s[i] == t0 && bytes_equal(s + i + 1, t + 1, m - 1)) return i;
    return -1;
}

CUSTR_HD int rfind_bytes(const uint8_t* s, int from, int to, const uint8_t* t, int m)
{
    for (int i = to - m; i >= from; --i)
        if (bytes_equal(s + i, t, m)) return i;
    return -1;
}

// byte offset `k` characters after byte offset `off`
CUSTR_HD int advance_chars(const uint8_t* s, int n, int off, int k)
{
    while (k > 0 && off < n) {
        ++off;
        while (off < n && (s[off] & 0xC0) == 0x80) ++off;
        --k;
    }
    return off;
}

// NVStrings::find / rfind per row: character index of the hit, -1 if none (find.cu:75-120,163-199)
CUSTR_HD int find_chars(const uint8_t* s, int n, const uint8_t* t, int m, int start, int end, bool reverse)
{
    if (m <= 0) return -1;
    int nchars = utf8_count_chars(s, n);
    bool ascii = nchars == n;
    int count = end - start;
    if (count < 0 && !reverse) count = nchars;  // rfind has no such normalisation (custring_view.inl:557-560)
    long long e = (long long)start + count;
    int cend = (e < 0 || e > nchars) ? nchars : (int)e;
    int spos = ascii ? (start < n ? start : n) : utf8_offset_of(s, n, start);
    int epos = ascii ? cend : utf8_offset_of(s, n, cend);
    int hit = reverse ? rfind_bytes(s, spos, epos, t, m) : find_bytes(s, spos, epos, t, m);
    if (hit < 0) return -1;
    return ascii ? hit : utf8_count_chars(s, hit);
}

// NVStrings::replace(str, repl, maxrepl) per row (modify.cu:125-187): returns new length, writes when o != 0
CUSTR_HD int replace_literal(const uint8_t* s, int n, const uint8_t* t, int m, const uint8_t* r, int rl, int maxrepl, char* o)
{
    int budget = maxrepl < 0 ? utf8_count_chars(s, n) : maxrepl;
    int total = n, last = 0;
    int pos = find_bytes(s, 0, n, t, m);
    while (pos >= 0 && budget > 0) {
        total += rl - m;
        if (o) {
            for (int k = last; k < pos; ++k) *o++ = (char)s[k];
            for (int k = 0; k < rl; ++k) *o++ = (char)r[k];
        }
        last = pos + m;
        pos = find_bytes(s, last, n, t, m);
        --budget;
    }
    if (o) for (int k = last; k < n; ++k) *o++ = (char)s[k];
    return total;
}

// NVStrings::replace(targets, repls) per row (modify.cu:196-260): at each byte the first matching target wins
CUSTR_HD int replace_literal_multi(const uint8_t* s, int n, const ColView& tg, const ColView& rp, char* o)
{
    int total = n, last = 0, pos = 0;
    while (pos < n) {
        int step = 1;
        for (int t = 0; t < tg.n; ++t) {
            if (!tg.valid(t)) continue;
            int tb = tg.offsets[t], tl = tg.offsets[t + 1] - tb;
            if (tl == 0 || tl > n - pos) continue;  // the reference never terminates on an empty target; skip it
            if (!bytes_equal(s + pos, (const uint8_t*)tg.chars + tb, tl)) continue;
            int ri = rp.n == 1 ? 0 : t;
            int rb = rp.offsets[ri], rl = rp.valid(ri) ? rp.offsets[ri + 1] - rb : 0;
            total += rl - tl;
            if (o) {
                for (int k = last; k < pos; ++k) *o++ = (char)s[k];
                for (int k = 0; k < rl; ++k) *o++ = rp.chars[rb + k];
            }
            last = pos + tl;
            step = tl;
            break;
        }
        pos += step;
    }
    if (o) for (int k = last; k < n; ++k) *o++ = (char)s[k];
    return total;
}

// ---- split -------------------------------------------------------------------------------------------------
enum TokKind { TOK_BYTES = 0, TOK_EMPTY = 1, TOK_NULL = 2 };

// number of tokens of split(delimiter) for one non-null row (custring_view.inl:1223-1250)
CUSTR_HD int split_count(const uint8_t* s, int n, const uint8_t* d, int m, int limit)
{
    if (n == 0) return 1;
    int hits = 0;
    if (m > 0) {
        int p = find_bytes(s, 0, n, d, m);
        while (p >= 0) {
            ++hits;
            p = find_bytes(s, advance_chars(s, n, p, m), n, d, m);
        }
    }
    int r = hits + 1;
    return (limit > 0 && r > limit) ? limit : r;
}

// walks the `dcount` tokens of a row in order (split.cu:768-806 evaluated for every column at once)
struct SplitWalk {
    const uint8_t* s; int n; const uint8_t* d; int m; int dcount;
    int c, spos; bool failed;
    CUSTR_HD void init(const uint8_t* s_, int n_, const uint8_t* d_, int m_, int dcount_)
    { s = s_; n = n_; d = d_; m = m_; dcount = dcount_; c = 0; spos = 0; failed = false; }
    CUSTR_HD bool next(int& b, int& e)
    {
        if (c >= dcount) return false;
        b = spos; e = n;
        if (!failed && c < dcount - 1) {
            int hit = m > 0 ? find_bytes(s, spos, n, d, m) : -1;
            if (hit < 0) failed = true;
            else { e = hit; spos = hit + m; }
        }
        ++c;
        if (b >= e) { b = 0; e = 0; }  // empty (not null) token
        return true;
    }
};

// whitespace split: number of columns the reference reports (split.cu:52-87): capped by limit, never 0
CUSTR_HD int wsplit_count(const uint8_t* s, int n, int limit)
{
    int cnt = 0;
    bool in_tok = false;
    for (int i = 0; i < n; ++i) {
        bool sp = s[i] <= ' ';
        if (!sp && !in_tok) ++cnt;
        in_tok = !sp;
    }
    if (limit > 0 && cnt > limit) cnt = limit;
    return cnt == 0 ? 1 : cnt;
}

// walks whitespace-separated tokens; with limit L the L-th token is the untrimmed remainder (split.cu:892-941).
// A row without any token yields ONE entry: null for split() columns, "" for split_record (split.cu:375-386).
struct WsWalk {
    const uint8_t* s; int n; int limit; int c, pos; bool done;
    CUSTR_HD void init(const uint8_t* s_, int n_, int limit_) { s = s_; n = n_; limit = limit_; c = 0; pos = 0; done = false; }
    // returns false when exhausted; kind tells whether [b,e) is a real token or the "no token" placeholder
    CUSTR_HD bool next(int& b, int& e, bool& placeholder)
    {
        placeholder = false;
        if (done) return false;
        while (pos < n && s[pos] <= ' ') ++pos;
        if (pos >= n) {
            done = true;
            if (c == 0) { b = e = 0; placeholder = true; ++c; return true; }
            return false;
        }
        b = pos;
        if (limit > 0 && c + 1 == limit) { e = n; pos = n; done = true; ++c; return true; }
        while (pos < n && s[pos] > ' ') ++pos;
        e = pos;
        ++c;
        return true;
    }
};

// ---- right-to-left splits ------------------------------------------------------------------------------------
// rsplit(delimiter): tokens numbered left to right, found from the right (split.cu:1003-1030 column pick,
// custring_view.inl:1281-1336 record form; both produce the same pieces).  f(k, begin, end).
template <typename F>
CUSTR_HD void rsplit_walk(const uint8_t* s, int n, const uint8_t* d, int m, int dcount, F f)
{
    int epos = n;
    for (int c = dcount - 1; c > 0; --c) {
        const int p = m > 0 ? rfind_bytes(s, 0, epos, d, m) : -1;
        if (p < 0) {  // search failed early: every remaining column repeats the remainder (split.cu:1008-1012)
            for (int k = c; k > 0; --k) f(k, 0, epos);
            break;
        }
        f(c, p + m, epos);
        epos = p;
    }
    f(0, 0, epos);
}

// rsplit_record(whitespace) (split.cu:563-596,655-686): the scan runs right to left over whitespace (<= ' ') runs; with
// a token limit the leftmost piece is the untrimmed remainder [0, epos).  A row without tokens yields one EMPTY string.
template <typename F>
CUSTR_HD void rwsplit_record_walk(const uint8_t* s, int n, int dcount, int limit, F f)
{
    int sidx = dcount - 1, epos = n;
    bool spaces = true;
    for (int pos = n; pos > 0 && sidx >= 0; --pos) {
        const bool sp = s[pos - 1] <= ' ';
        if (spaces == sp) {
            if (spaces) epos = pos - 1;
            continue;
        }
        if (!spaces) {
            if (dcount - sidx == limit) break;
            f(sidx--, pos, epos);
            epos = pos - 1;
        }
        spaces = !spaces;
    }
    if (sidx >= 0) {
        if (epos > 0) f(sidx, 0, epos);
        else f(sidx, 0, 0);
        --sidx;
    }
    for (; sidx >= 0; --sidx) f(sidx, 0, 0);
}

// rsplit(whitespace) column pick (split.cu:1098-1137), evaluated for one column `col` of `ncols`: false = null entry.
// (The limit test uses the column count of the whole call, not the row's own token count — kept as is.)
CUSTR_HD bool rwsplit_column(const uint8_t* s, int n, int dcount, int ncols, int limit, int col, int& b, int& e)
{
    int c = dcount - 1, spos = 0, epos = n;
    bool spaces = true;
    for (int pos = n; pos > 0; --pos) {
        const bool sp = s[pos - 1] <= ' ';
        if (spaces == sp) {
            if (spaces) epos = pos - 1;
            else spos = pos - 1;
            continue;
        }
        if (!spaces) {
            spos = 0;
            if (ncols - c == limit) break;
            spos = pos;
            if (c == col) break;
            epos = pos - 1;
            spos = 0;
            --c;
        }
        spaces = !spaces;
    }
    if (spos < epos) { b = spos; e = epos; return true; }
    return false;
}

// ---- tokenize (text/tokens.cu:41-121): delimiter = whitespace (null set) or any character of a set -----------
struct DelimSet {
    const uint32_t* chars;  // packed chars; nullptr => whitespace (byte <= ' ')
    int count;
};
CUSTR_HD bool is_delim(const DelimSet& ds, const uint8_t* p, const uint8_t* end, int& width)
{
    if (!ds.chars) { width = 1; return *p <= ' '; }
    uint32_t c = utf8_packed(p, end, width);
    for (int i = 0; i < ds.count; ++i)
        if (ds.chars[i] == c) return true;
    return false;
}
struct TokenWalk {
    const uint8_t* s; int n; DelimSet ds; int pos;
    CUSTR_HD void init(const uint8_t* s_, int n_, const DelimSet& ds_) { s = s_; n = n_; ds = ds_; pos = 0; }
    CUSTR_HD bool next(int& b, int& e)
    {
        int w = 1;
        while (pos < n && is_delim(ds, s + pos, s + n, w)) pos += w;
        if (pos >= n) return false;
        b = pos;
        while (pos < n && !is_delim(ds, s + pos, s + n, w)) pos += w;
        e = pos < n ? pos : n;
        return true;
    }
};

// byte-lexicographic comparison, shorter first (custring.inl:240-261)
CUSTR_HD int compare_bytes(const uint8_t* a, int an, const uint8_t* b, int bn)
{
    int m = an < bn ? an : bn;
    for (int i = 0; i < m; ++i)
        if (a[i] != b[i]) return (int)a[i] - (int)b[i];
    return an - bn;
}

}  // namespace row
}  // namespace custr
