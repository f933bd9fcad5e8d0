// Exact per-row Pike VM over the compiled program (regex_prog.h).  This is the engine of record: it
// reproduces the reference's dreprog::regexec (cpp/src/regex/regexec.inl:204-442) step for step — seed order,
// breadth-wise epsilon rounds, first-activation-wins de-duplication, END cutting lower-priority threads — so
// that match existence AND match spans are bit-identical.  The faster bit-parallel tier (regex_bits.cu) is only
// used for patterns it can prove equivalent; everything else runs here.
//
// Differences in mechanism (not in results): positions are byte offsets (the reference walks char indices and
// re-derives byte offsets with O(n) scans, custring_view.inl:261-281); lists live in per-thread local arrays
// sized by instruction-count tier; resetting a list clears only the bits it set; ASCII class tests are a
// 128-bit bitmap lookup.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"
#include "regex_prog.h"

namespace custr {
namespace rxdev {

struct DevProg {
    const rx::DevHeader* h;
    const rx::Inst* insts;
    const int32_t* starts;
    const rx::DevClass* classes;
    const uint32_t* ranges;
    const uint8_t* uflags;
};

CUSTR_HD DevProg bind_program(const uint8_t* img, const uint8_t* uflags)
{
    DevProg p;
    p.h = (const rx::DevHeader*)img;
    p.insts = (const rx::Inst*)(img + sizeof(rx::DevHeader));
    p.starts = (const int32_t*)(p.insts + p.h->ninsts);
    p.classes = (const rx::DevClass*)(p.starts + p.h->nstarts);
    p.ranges = (const uint32_t*)(p.classes + p.h->nclasses);
    p.uflags = uflags;
    return p;
}

CUSTR_HD bool class_match(const DevProg& P, int cls, uint32_t c)
{
    const rx::DevClass& k = P.classes[cls];
    if (c < 128u) return (k.ascii[c >> 5] >> (c & 31)) & 1u;
    const uint32_t* r = P.ranges + k.range_begin;
    for (int i = 0; i < k.range_count; i += 2)
        if (c >= r[i] && c <= r[i + 1]) return true;
    int b = k.builtins;
    if (!b) return false;
    uint32_t cp = packed_to_cp(c);
    if (cp > 0xFFFFu) return false;
    uint32_t f = CUSTR_LDG(P.uflags + cp);
    bool alnum = (f & 15u) != 0, space = (f & 16u) != 0, digit = (f & 4u) != 0;
    if ((b & rx::CB_W) && alnum) return true;
    if ((b & rx::CB_S) && space) return true;
    if ((b & rx::CB_D) && digit) return true;
    if ((b & rx::CB_NW) && !alnum) return true;   // c is not '\n' / '_' here (non-ASCII)
    if ((b & rx::CB_NS) && !space) return true;
    if ((b & rx::CB_ND) && !digit) return true;
    return false;
}

// GROUPS: threads also carry the end of the tracked capture group (reference Relist ranges are int2, regexec.inl:26-108)
template <int CAP, bool GROUPS = false>
struct Lists {
    uint16_t id[2][CAP];
    int32_t beg[2][CAP];
    int32_t endp[2][GROUPS ? CAP : 1];
    uint32_t mask[2][(CAP + 31) / 32];
    int size[2];

    CUSTR_HD void init()
    {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            for (int w = 0; w < (CAP + 31) / 32; ++w) mask[k][w] = 0;
            size[k] = 0;
        }
    }
    CUSTR_HD void reset(int k)
    {
        for (int i = 0; i < size[k]; ++i) {
            int v = id[k][i];
            mask[k][v >> 5] &= ~(1u << (v & 31));
        }
        size[k] = 0;
    }
    CUSTR_HD void activate(int k, int inst, int b, int e = -1)
    {
        uint32_t bit = 1u << (inst & 31);
        if (mask[k][inst >> 5] & bit) return;
        mask[k][inst >> 5] |= bit;
        id[k][size[k]] = (uint16_t)inst;
        beg[k][size[k]] = b;
        if (GROUPS) endp[k][size[k]] = e;
        ++size[k];
    }
};

// CAP == 0: programs with more instructions than the largest in-thread tier (1024).  The lists live in a global scratch
// arena — one slice per RESIDENT thread, bound by init() from the arena registered for the launch (the reference allocates
// 2 x Relist::alloc_size(insts) per ROW instead, regexec.cpp:81-95, count.cu:92-100).  Same operations, pointers instead of
// arrays.  Device only.
struct VmArena {
    uint8_t* base;
    unsigned long long stride;  // bytes per thread
    int ninsts;
};
#ifdef __CUDACC__
static __device__ VmArena g_vm_arena;  // per translation unit; only regex.cu's kernels (and its host code) use it
#endif
CUSTR_HD size_t vm_arena_bytes_per_thread(int ninsts)
{
    const size_t n = (size_t)((ninsts + 31) & ~31);
    // per list: id u16[n] + beg i32[n] + endp i32[n] + mask u32[n/32]; two lists, two Lists objects (plain + group tracking)
    return 2 * 2 * (2 * n + 4 * n + 4 * n + n / 8) + 64;
}
template <bool GROUPS>
struct Lists<0, GROUPS> {
    uint16_t* id[2];
    int32_t* beg[2];
    int32_t* endp[2];
    uint32_t* mask[2];
    int size[2];
    int nwords;

    CUSTR_HD void init()
    {
#ifdef __CUDA_ARCH__
        const VmArena a = g_vm_arena;
        const size_t n = (size_t)((a.ninsts + 31) & ~31);
        const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint8_t* p = a.base + tid * a.stride + (GROUPS ? a.stride / 2 : 0);  // the two Lists objects of a thread get one half each
        p = (uint8_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
        nwords = (int)(n / 32);
        for (int k = 0; k < 2; ++k) {
            beg[k] = (int32_t*)p; p += 4 * n;
            endp[k] = (int32_t*)p; p += 4 * n;
            mask[k] = (uint32_t*)p; p += n / 8;
            id[k] = (uint16_t*)p; p += 2 * n;
            for (int w = 0; w < nwords; ++w) mask[k][w] = 0;
            size[k] = 0;
        }
#endif
    }
    CUSTR_HD void reset(int k)
    {
        for (int i = 0; i < size[k]; ++i) {
            int v = id[k][i];
            mask[k][v >> 5] &= ~(1u << (v & 31));
        }
        size[k] = 0;
    }
    CUSTR_HD void activate(int k, int inst, int b, int e = -1)
    {
        uint32_t bit = 1u << (inst & 31);
        if (mask[k][inst >> 5] & bit) return;
        mask[k][inst >> 5] |= bit;
        id[k][size[k]] = (uint16_t)inst;
        beg[k][size[k]] = b;
        if (GROUPS) endp[k][size[k]] = e;
        ++size[k];
    }
};

// Search row bytes s[0..n) starting at byte offset `begin`; new threads are seeded at offset `off` only while
// off < seed_limit (n for an unanchored search; begin+1 for "match only here": reference `end = begin+1`).
// On success mbeg/mend are the byte offsets of the winning match.
// GROUPS / gid: track capture group `gid` (>= 1) instead of the whole match: threads are seeded with (-1,-1), LBRA / RBRA
// of that group record the position, END reports the group's span (reference regexec.inl:296-307,422-425 with groupId).
template <int CAP, bool GROUPS = false>
__host__ __device__ int vm_find(const DevProg& P, const uint8_t* __restrict__ s, int n, int begin, int seed_limit, int& mbeg,
                       int& mend, Lists<CAP, GROUPS>& L, int gid = 0)
{
    int match = 0;
    int off = begin;
    int cur = 0;
    L.reset(0);
    L.reset(1);
    const int ninsts = P.h->ninsts;
    const int nstarts = P.h->nstarts;
    const int start_op = P.h->start_op;
    const uint32_t start_arg = P.h->start_arg;
    uint32_t prev = (begin > 0 && begin <= n) ? utf8_packed_before(s + begin, s) : 0;
    uint32_t c;
    do {
        if (L.size[cur] == 0) {  // start-skip (regexec.inl:217-246); result-neutral, saves work
            if (start_op == rx::OP_CHAR && start_arg != 0 && start_arg < 0x80u) {
                int o = off;
                while (o < n && s[o] != (uint8_t)start_arg) ++o;
                if (o >= n) return match;
                if (o != off) { off = o; prev = utf8_packed_before(s + off, s); }
            } else if (start_op == rx::OP_BOL && off != 0 && start_arg != '^')
                return match;
        }
        if (off < seed_limit && !match)
            for (int i = 0; i < nstarts; ++i) L.activate(cur, P.starts[i], GROUPS ? -1 : off, -1);

        int w = 1;
        c = off < n ? utf8_packed(s + off, s + n, w) : 0;

        // ---- epsilon rounds: rebuild the list until nothing expands
        bool expanded;
        int rounds = 0;
        do {
            const int nxt = cur ^ 1;
            L.reset(nxt);
            expanded = false;
            for (int i = 0; i < L.size[cur]; ++i) {
                const int id = L.id[cur][i];
                int b = L.beg[cur][i];
                int e = GROUPS ? L.endp[cur][i] : -1;
                const rx::Inst in = P.insts[id];
                int go = -1;
                switch (in.op) {
                case rx::OP_CHAR: case rx::OP_ANY: case rx::OP_ANYNL: case rx::OP_CLASS: case rx::OP_NCLASS: case rx::OP_END:
                    go = id;
                    break;
                case rx::OP_LBRA:
                    if (GROUPS && (int)in.arg == gid) b = off;
                    go = in.next; expanded = true;
                    break;
                case rx::OP_RBRA:
                    if (GROUPS && (int)in.arg == gid) e = off;
                    go = in.next; expanded = true;
                    break;
                case rx::OP_BOL:
                    if (off == 0 || (in.arg == '^' && prev == '\n')) { go = in.next; expanded = true; }
                    break;
                case rx::OP_EOL:
                    if (c == 0 || (in.arg == '$' && c == '\n')) { go = in.next; expanded = true; }
                    break;
                case rx::OP_BOW: case rx::OP_NBOW: {
                    bool ca = is_alnum_packed(c, P.uflags);
                    bool pa = is_alnum_packed(off ? prev : 0, P.uflags);
                    if ((ca != pa) == (in.op == rx::OP_BOW)) { go = in.next; expanded = true; }
                    break;
                }
                case rx::OP_SPLIT:
                    L.activate(nxt, in.other, b, e);
                    go = in.next; expanded = true;
                    break;
                default: break;  // OP_BAD: thread dies
                }
                if (go >= 0) L.activate(nxt, go, b, e);
            }
            cur = nxt;
        } while (expanded && ++rounds <= ninsts + 1);

        // ---- consume c
        {
            const int nxt = cur ^ 1;
            L.reset(nxt);
            for (int i = 0; i < L.size[cur]; ++i) {
                const int id = L.id[cur][i];
                const rx::Inst in = P.insts[id];
                bool take = false;
                switch (in.op) {
                case rx::OP_CHAR: take = in.arg == c; break;
                case rx::OP_ANY: take = c != '\n'; break;
                case rx::OP_ANYNL: take = true; break;
                case rx::OP_CLASS: take = class_match(P, (int)in.arg, c); break;
                case rx::OP_NCLASS: take = !class_match(P, (int)in.arg, c); break;
                case rx::OP_END:
                    match = 1;
                    mbeg = L.beg[cur][i];
                    mend = GROUPS ? L.endp[cur][i] : off;
                    i = L.size[cur];  // cut every lower-priority thread
                    break;
                default: break;
                }
                if (take) L.activate(nxt, in.next, L.beg[cur][i], GROUPS ? L.endp[cur][i] : -1);
            }
            cur = nxt;
        }
        off += w;
        prev = c;
    } while (c && (L.size[cur] > 0 || !match));
    return match;
}

// ---- per-row drivers (shared by the kernels in regex.cu and by the host simulation in tests/sim) ----------

// count.cu:168-196: number of non-overlapping matches
template <int CAP>
__host__ __device__ int row_count(const DevProg& P, const uint8_t* __restrict__ s, int n, Lists<CAP>& L)
{
    int found = 0, begin = 0;
    while (begin <= n) {
        int mb = 0, me = 0;
        if (!vm_find<CAP>(P, s, n, begin, n, mb, me, L)) break;
        ++found;
        if (me > mb) begin = me;
        else begin = mb + (mb < n ? utf8_width(s[mb]) : 1);
    }
    return found;
}

// replace.cu:39-107: returns the new byte length; writes the new bytes when o != nullptr
template <int CAP>
__host__ __device__ int row_replace(const DevProg& P, const uint8_t* __restrict__ s, int n, const char* __restrict__ repl,
                                    int repl_len, int maxrepl, char* o, Lists<CAP>& L)
{
    int budget = maxrepl < 0 ? utf8_count_chars(s, n) : maxrepl;
    int total = n, last = 0, begin = 0;
    while (budget > 0) {
        int mb = 0, me = 0;
        if (!vm_find<CAP>(P, s, n, begin, n, mb, me, L)) break;
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

// replace_multi.cu:40-106: at every character position try each program anchored there, first hit wins
template <int CAP>
__host__ __device__ int row_replace_multi(const uint8_t* const* images, int nprogs, const uint8_t* uflags, const ColView& repls,
                                          const uint8_t* __restrict__ s, int n, char* o, Lists<CAP>& L)
{
    int total = n, last = 0, pos = 0;
    while (pos < n) {
        int adv = utf8_width(s[pos]);
        for (int t = 0; t < nprogs; ++t) {
            DevProg P = bind_program(images[t], uflags);
            int mb = 0, me = 0;
            if (!vm_find<CAP>(P, s, n, pos, pos + 1, mb, me, L)) continue;
            int r = repls.n == 1 ? 0 : t;
            int rb = repls.offsets[r], rl = repls.valid(r) ? repls.offsets[r + 1] - rb : 0;
            total += rl - (me - mb);
            if (o) {
                for (int k = last; k < mb; ++k) *o++ = (char)s[k];
                for (int k = 0; k < rl; ++k) *o++ = repls.chars[rb + k];
                last = me;
            }
            if (me > pos) adv = me - pos;  // the reference spins forever on an empty match; we step on
            break;
        }
        pos += adv;
    }
    if (o) for (int k = last; k < n; ++k) *o++ = (char)s[k];
    return total;
}

}  // namespace rxdev
}  // namespace custr
