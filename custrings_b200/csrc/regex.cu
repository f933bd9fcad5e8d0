// Regex entry points of the C-ABI: contains_re / match / count_re / replace_re (+ multi-pattern form).
// Replaces NVStrings::contains_re count.cu:59-110, match :113-165, count_re :199-250, replace_re
// replace.cu:110-189 and replace_multi.cu:110-197 of the reference.
//
// Execution tiers:
//   bitstream  - regex_bits.cu: whole-buffer bit-parallel evaluation for boolean results, when the pattern
//                lowers to a bitstream program that is provably equivalent to the reference's NFA semantics
//   pikevm     - this file + regex_vm.cuh: exact thread-per-row Pike VM, any pattern, also yields spans
#include "common.cuh"
#include "regex_vm.cuh"
#include "regex_bits.h"
#include "chain_spans.cuh"
#include "span_walk.cuh"
#include "row_stage.cuh"
#include <cub/cub.cuh>
#include <list>
#include <map>
#include <mutex>

namespace custr {

static const uint8_t k_unicode_flags_host[65536] = {
#include "unicode_flags.inc"
};
const uint8_t* host_unicode_flags() { return k_unicode_flags_host; }
static int current_device()
{
    int dev = 0;
    CUSTR_CUDA(cudaGetDevice(&dev));
    return dev;
}
// one table per device: after custr_set_device() the kernels must not dereference the first device's copy
const uint8_t* device_unicode_flags()
{
    static std::mutex mu;
    static std::map<int, uint8_t*> tables;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    auto it = tables.find(dev);
    if (it != tables.end()) return it->second;
    uint8_t* d = nullptr;
    CUSTR_CUDA(cudaMalloc(&d, 65536));
    CUSTR_CUDA(cudaMemcpy(d, k_unicode_flags_host, 65536, cudaMemcpyHostToDevice));
    tables[dev] = d;
    return d;
}

thread_local const char* g_last_tier = "none";
thread_local int g_forced_tier = 0;

// optional device-side timing of the dominant kernel of a call (bench.py's roofline leg)
thread_local int g_profile = 0;
thread_local float g_last_kernel_ms = -1.f;
struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool armed = false;
    void start()
    {
        if (!g_profile) return;
        if (!a) { CUSTR_CUDA(cudaEventCreate(&a)); CUSTR_CUDA(cudaEventCreate(&b)); }
        CUSTR_CUDA(cudaEventRecord(a, g_stream));
        armed = true;
    }
    void stop() { if (armed) CUSTR_CUDA(cudaEventRecord(b, g_stream)); }
    void collect()
    {
        if (!armed) return;
        CUSTR_CUDA(cudaEventSynchronize(b));
        CUSTR_CUDA(cudaEventElapsedTime(&g_last_kernel_ms, a, b));
        armed = false;
    }
};
static thread_local KernelTimer g_timer;

// ---- compiled-program cache ------------------------------------------------------------------------------
struct Compiled {
    rx::Program prog;
    std::vector<uint8_t> image;
    BufPtr dev_image;
    std::shared_ptr<bits::Plan> plan_contains, plan_match;  // bitstream lowering (may be null = not eligible)
    bool plans_built = false;
};
using CompiledPtr = std::shared_ptr<Compiled>;

// Entries are keyed by (device, pattern): the program image lives in the memory of the device that was current when it
// was uploaded.  Both bit-stream plans are built here, under the lock, so a cached entry is immutable once visible.
static CompiledPtr get_compiled(const char* pattern)
{
    static std::mutex mu;
    static std::list<std::pair<std::string, CompiledPtr>> lru;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    std::string key = std::to_string(dev) + ":" + pattern;
    for (auto it = lru.begin(); it != lru.end(); ++it)
        if (it->first == key) {
            lru.splice(lru.begin(), lru, it);
            return lru.front().second;
        }
    CompiledPtr c = std::make_shared<Compiled>();
    c->prog = rx::compile(pattern);
    c->image = rx::serialize(c->prog, host_unicode_flags());
    c->dev_image = upload(c->image.data(), c->image.size());
    c->plan_contains = bits::lower(c->prog, false, host_unicode_flags());
    c->plan_match = bits::lower(c->prog, true, host_unicode_flags());
    c->plans_built = true;
    // the image must be resident before another stream/thread uses the cached entry
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    lru.emplace_front(key, c);
    if (lru.size() > 64) lru.pop_back();
    return c;
}

// A program that is nothing but CHAR, CHAR, ..., END (e.g. the README day-of-week chain: replace("Sun","0") with the
// default regex=True) matches exactly like the literal: same leftmost, non-overlapping spans.  Returns its UTF-8 bytes.
static bool pure_literal(const rx::Program& p, std::string& lit)
{
    if (p.malformed || p.ngroups != 0 || p.insts.empty()) return false;
    lit.clear();
    int id = p.start_inst;
    for (size_t guard = 0; guard <= p.insts.size(); ++guard) {
        const rx::Inst& in = p.insts[id];
        if (in.op == rx::OP_END) return !lit.empty();
        if (in.op != rx::OP_CHAR || in.arg == 0) return false;
        for (int shift = 24; shift >= 0; shift -= 8) {
            unsigned b = (in.arg >> shift) & 0xFFu;
            if (b || shift == 0) { if (b) lit.push_back((char)b); }
        }
        id = in.next;
    }
    return false;
}

static int cap_tier(int ninsts);
static const bits::ChainDev* span_plan(Compiled& c)
{
    if (g_forced_tier == 1) return nullptr;
    if (!c.plans_built) {
        c.plan_contains = bits::lower(c.prog, false, host_unicode_flags());
        c.plan_match = bits::lower(c.prog, true, host_unicode_flags());
        c.plans_built = true;
    }
    return c.plan_contains ? bits::span_chain(*c.plan_contains) : nullptr;
}

// list capacity tier of the Pike-VM kernels; 0 = no in-thread tier fits: lists in the global scratch arena (Lists<0>)
static int cap_tier(int ninsts)
{
    if (ninsts <= 32) return 32;
    if (ninsts <= 256) return 256;
    if (ninsts <= 1024) return 1024;
    return 0;
}
constexpr int VM_MAX_INSTS = 65535;  // instruction ids are 16 bits in the lists

// ---- kernels ---------------------------------------------------------------------------------------------
constexpr int VM_THREADS = 128;
constexpr int SMEM_PROG_MAX = 40 * 1024;

__device__ __forceinline__ const uint8_t* stage_program(const uint8_t* img, int img_bytes, uint8_t* smem)
{
    if (img_bytes > SMEM_PROG_MAX) return img;
    for (int i = threadIdx.x * 4; i < img_bytes; i += blockDim.x * 4) *(uint32_t*)(smem + i) = *(const uint32_t*)(img + i);
    __syncthreads();
    return smem;
}

template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_bool(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags, int anchored,
          uint8_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    L.init();
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        int hit = 0;
        if (i < col.n) {
            if (col.valid(i)) {
                int b = col.offsets[i], n = col.offsets[i + 1] - b;
                int mb, me;
                hit = rxdev::vm_find<CAP>(P, (const uint8_t*)col.chars + b, n, 0, anchored ? 1 : n, mb, me, L);
            }
            out[i] = (uint8_t)hit;
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
    }
}

// exact decision for the rows the bitstream tier could not decide (non-ASCII / NUL bytes); the list length is read
// from device memory so no host round trip sits between the two kernels
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_bool_rows(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags, int anchored,
               const int32_t* __restrict__ rows, const unsigned int* __restrict__ nrows_ptr, uint8_t* __restrict__ out,
               unsigned long long* __restrict__ total)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    L.init();
    const int nrows = (int)*nrows_ptr;
    for (int base = blockIdx.x * blockDim.x; base < nrows; base += gridDim.x * blockDim.x) {
        int k = base + threadIdx.x;
        int hit = 0;
        if (k < nrows) {
            int i = rows[k];
            int b = col.offsets[i], n = col.offsets[i + 1] - b;
            int mb, me;
            hit = rxdev::vm_find<CAP>(P, (const uint8_t*)col.chars + b, n, 0, anchored ? 1 : n, mb, me, L);
            out[i] = (uint8_t)hit;
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
    }
}

template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_count(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags,
           int32_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    L.init();
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        int found = 0;
        if (i < col.n) {
            if (col.valid(i)) {
                int b = col.offsets[i], n = col.offsets[i + 1] - b;
                found = rxdev::row_count<CAP>(P, (const uint8_t*)col.chars + b, n, L);
            }
            out[i] = found;
        }
        unsigned m = __ballot_sync(0xffffffffu, found != 0);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
    }
}

// replace.cu:39-107 — pass 1 (out_chars == nullptr) writes the new byte length per row, pass 2 writes bytes
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_replace(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags,
             const char* __restrict__ repl, int repl_len, int maxrepl, int32_t* __restrict__ out_len,
             const int32_t* __restrict__ out_off, char* __restrict__ out_chars)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    L.init();
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        if (i >= col.n) continue;
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        int b = col.offsets[i], n = col.offsets[i + 1] - b;
        const uint8_t* s = (const uint8_t*)col.chars + b;
        int total = rxdev::row_replace<CAP>(P, s, n, repl, repl_len, maxrepl, out_chars ? out_chars + out_off[i] : nullptr, L);
        if (!out_chars) out_len[i] = total;
    }
}

// replace_backref.cu:36-118 — every match is replaced by the template with its back-references filled in.  refs[j] =
// (group number, byte position inside the template with the \N markers removed); group 0 is the whole match, a group that
// did not take part (or does not exist) contributes nothing.  Pass 1 (out_chars == nullptr) sizes, pass 2 writes.
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_backrefs(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags, const char* __restrict__ tmpl,
              int tmpl_len, const int2* __restrict__ refs, int nrefs, int32_t* __restrict__ out_len, const int32_t* __restrict__ out_off,
              char* __restrict__ out_chars)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    rxdev::Lists<CAP, true> LG;
    L.init();
    LG.init();
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        if (i >= col.n) continue;
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        const int b = col.offsets[i], n = col.offsets[i + 1] - b;
        const uint8_t* s = (const uint8_t*)col.chars + b;
        char* o = out_chars ? out_chars + out_off[i] : nullptr;
        int total = 0, lpos = 0, begin = 0;
        auto put = [&](const char* src, int len) {
            total += len;
            if (o) for (int k = 0; k < len; ++k) *o++ = src[k];
        };
        while (begin <= n) {
            int mb = 0, me = 0;
            if (!rxdev::vm_find<CAP>(P, s, n, begin, n, mb, me, L)) break;
            put((const char*)s + lpos, mb - lpos);
            int ilpos = 0;
            for (int j = 0; j < nrefs; ++j) {
                put(tmpl + ilpos, refs[j].y - ilpos);
                ilpos = refs[j].y;
                int gb = mb, ge = me;
                if (refs[j].x > 0) {
                    gb = ge = -1;
                    if (!rxdev::vm_find<CAP, true>(P, s, n, mb, mb + 1, gb, ge, LG, refs[j].x)) continue;
                }
                if (gb >= 0 && ge > gb) put((const char*)s + gb, ge - gb);
            }
            put(tmpl + ilpos, tmpl_len - ilpos);
            lpos = me;
            // (the reference never terminates on an empty match, replace_backref.cu:101; step over one character instead)
            begin = me > mb ? me : me + (me < n ? utf8_width(s[me]) : 1);
        }
        put((const char*)s + lpos, n - lpos);
        if (!out_chars) out_len[i] = total;
    }
}

// replace_multi.cu:40-106 — at every character position try each program anchored there, first hit wins
struct MultiProgs {
    const uint8_t* const* images;  // device array of device images
    int count;
};
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_replace_multi(ColView col, MultiProgs progs, const uint8_t* __restrict__ uflags, ColView repls, int32_t* __restrict__ out_len,
                   const int32_t* __restrict__ out_off, char* __restrict__ out_chars)
{
    rxdev::Lists<CAP> L;
    L.init();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        int b = col.offsets[i], n = col.offsets[i + 1] - b;
        const uint8_t* s = (const uint8_t*)col.chars + b;
        int total = rxdev::row_replace_multi<CAP>(progs.images, progs.count, uflags, repls, s, n,
                                                  out_chars ? out_chars + out_off[i] : nullptr, L);
        if (!out_chars) out_len[i] = total;
    }
}



// ---- span fast path for last-loop chains (chain_spans.cuh): thread-per-row scalar scan, no NFA lists -------------------
__global__ void __launch_bounds__(256)
k_chain_count(ColView col, const __grid_constant__ bits::ChainDev cd, const uint8_t* __restrict__ uflags, int32_t* __restrict__ out,
              unsigned long long* __restrict__ total, int* __restrict__ nul_seen)
{
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        int found = 0;
        if (i < col.n) {
            if (col.valid(i)) {
                const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
                const int n = col.offsets[i + 1] - col.offsets[i];
                if (spans::has_nul(s, n)) *nul_seen = 1;
                else found = spans::row_count(cd, s, n, uflags);
            }
            out[i] = found;
        }
        unsigned m = __ballot_sync(0xffffffffu, found != 0);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
    }
}

__global__ void __launch_bounds__(256)
k_chain_replace(ColView col, const __grid_constant__ bits::ChainDev cd, const uint8_t* __restrict__ uflags, const char* __restrict__ repl,
                int repl_len, int maxrepl, int32_t* __restrict__ out_len, const int32_t* __restrict__ out_off, char* __restrict__ out_chars,
                int* __restrict__ nul_seen)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        int n = col.offsets[i + 1] - col.offsets[i];
        if (!out_chars && spans::has_nul(s, n)) { *nul_seen = 1; out_len[i] = 0; continue; }
        int total = spans::row_replace(cd, s, n, uflags, repl, repl_len, maxrepl, out_chars ? out_chars + out_off[i] : nullptr);
        if (!out_chars) out_len[i] = total;
    }
}

// ---- span path over the chain kernel's bit streams (span_walk.cuh): a few word scans per match ------------------------
// Rows the chain kernel's boolean result rejected have no match: only the others are walked.  A warp compacts the hit rows
// of 128 consecutive rows into a dense list first, so every lane of the (divergent, loop-heavy) walk has a row to work on.
// (Staging the stream slices in shared memory as the splice does was measured slower here: these walks live on occupancy.)
template <typename Miss, typename Hit>
__device__ __forceinline__ void for_hit_rows(int n, const uint8_t* __restrict__ hits, Miss on_miss, Hit on_hit)
{
    constexpr int CHUNK = 128;
    __shared__ int32_t lists[8][CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    int32_t* list = lists[warp];
    for (int base = (blockIdx.x * warps + warp) * CHUNK; base < n; base += gridDim.x * warps * CHUNK) {
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < CHUNK / 32; ++k) {
            const int i = base + 32 * k + lane;
            const bool h = i < n && hits[i];
            if (i < n && !h) on_miss(i);
            const unsigned m = __ballot_sync(0xffffffffu, h);
            if (h) list[cnt + __popc(m & ((1u << lane) - 1u))] = i;
            cnt += __popc(m);
        }
        __syncwarp();
        for (int idx = lane; idx < cnt; idx += 32) on_hit(list[idx]);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
k_span_count(ColView col, spans::Streams S, const uint8_t* __restrict__ hits, int k_chars, int32_t* __restrict__ out,
             unsigned long long* __restrict__ total)
{
    unsigned long long mine = 0;
    for_hit_rows(
        col.n, hits, [&](int i) { out[i] = 0; },
        [&](int i) {
            const int found = spans::walk_spans(S, (const uint8_t*)col.chars, col.offsets[i], col.offsets[i + 1], k_chars, 0x7fffffff, [](int, int) {});
            out[i] = found;
            mine += found != 0;
        });
    if (mine) atomicAdd(total, mine);
}

// size pass of replace_re: new byte length of every row
__global__ void __launch_bounds__(256)
k_span_replace(ColView col, spans::Streams S, const uint8_t* __restrict__ hits, int k_chars, int repl_len, int maxrepl,
               int32_t* __restrict__ out_len)
{
    const uint8_t* chars = (const uint8_t*)col.chars;
    const int budget = maxrepl < 0 ? 0x7fffffff : maxrepl;
    for_hit_rows(
        col.n, hits, [&](int i) { out_len[i] = col.offsets[i + 1] - col.offsets[i]; },
        [&](int i) {
            const int a = col.offsets[i], b = col.offsets[i + 1];
            int total = b - a;
            spans::walk_spans(S, chars, a, b, k_chars, budget, [&](int s, int e) { total += repl_len - (e - s); });
            out_len[i] = total;
        });
}

// write pass as a warp-cooperative SPLICE.  A warp takes 32 consecutive rows: their chars are one contiguous byte range and
// so is their output.  (1) The range and its slices of the three bit streams are staged in shared memory with coalesced
// loads.  (2) Every lane walks the matches of its own row (word scans over the shared streams) and records them in two
// shared bit maps over the same byte range: DROP (bytes inside a match) and INS (first byte of a match).  (3) The warp
// then rewrites the whole range cooperatively — lane = one 64-byte word: kept bytes and replacement strings land at
// (warp prefix of popc(keep) + repl_len * popc(ins)) in a shared output tile, no per-row control flow — and (4) the
// tile leaves with coalesced 16-byte stores.  Blocks that do not fit (very long rows / much longer output) take the
// direct lane-per-row path.
namespace splice {
using u64 = unsigned long long;
constexpr int ROWS = 32, CAP = 4352, WARPS = 4, THREADS = WARPS * 32;
constexpr int WORDS = (CAP + 64) / 64 + 2;
struct __align__(16) Sm {
    char in[WARPS][CAP + 64 + 32];   // from the 64-byte aligned start of the block: in[p - a64]
    char out[WARPS][CAP + 32];       // out[(out_a & 15) + k]
    u64 m[WARPS][WORDS], k[WARPS][WORDS], a[WARPS][WORDS], drop[WARPS][WORDS], ins[WARPS][WORDS];
};
// relative positions [lo, hi), hi > lo; 32-bit shared-memory atomics (ATOMS.OR) on the halves of the 64-bit words
__device__ __forceinline__ void set_bits(uint32_t* arr, int lo, int hi)
{
    const int wl = lo >> 5, wh = (hi - 1) >> 5;
    for (int w = wl; w <= wh; ++w) {
        uint32_t mask = ~0u;
        if (w == wl) mask &= ~0u << (lo & 31);
        if (w == wh) mask &= ~0u >> (31 - ((hi - 1) & 31));
        atomicOr(arr + w, mask);
    }
}
}  // namespace splice

__global__ void __launch_bounds__(splice::THREADS)
k_span_replace_splice(ColView col, int chars_limit, spans::Streams S, const uint8_t* __restrict__ hits, int k_chars,
                      const char* __restrict__ repl, int repl_len, int maxrepl, const int32_t* __restrict__ out_off, char* __restrict__ out_chars)
{
    using namespace splice;
    __shared__ Sm sm;
    const uint8_t* chars = (const uint8_t*)col.chars;
    const int budget = maxrepl < 0 ? 0x7fffffff : maxrepl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblocks = (col.n + ROWS - 1) / ROWS;
    for (int blk = blockIdx.x * WARPS + warp; blk < nblocks; blk += gridDim.x * WARPS) {
        const int r0 = blk * ROWS, r1 = r0 + ROWS < col.n ? r0 + ROWS : col.n;
        const int in_a = col.offsets[r0], in_b = col.offsets[r1];
        const long long out_a = out_off[r0], out_b = out_off[r1];
        const int i = r0 + lane;
        const int a64 = in_a & ~63;
        const bool fits = in_b > in_a && (in_b - a64) <= CAP + 64 && (out_b - (out_a & ~15ll)) <= CAP + 16;
        if (!fits) {  // direct path, lane per row
            if (i < r1) {
                const int a = col.offsets[i], b = col.offsets[i + 1];
                char* o = out_chars + out_off[i];
                int last = a;
                if (hits[i])
                    spans::walk_spans<true>(S, chars, a, b, k_chars, budget, [&](int s, int e) {
                        for (int q = last; q < s; ++q) *o++ = (char)chars[q];
                        for (int q = 0; q < repl_len; ++q) *o++ = repl[q];
                        last = e;
                    });
                for (int q = last; q < b; ++q) *o++ = (char)chars[q];
            }
            continue;
        }
        // (1) stage chars + stream slices, clear the bit maps; the range of this warp's NEXT block is pulled towards L2 now
        {
            const int nb = blk + gridDim.x * WARPS;
            if (nb < nblocks) {
                const int na = col.offsets[nb * ROWS];
                const char* pf = col.chars + na + 128 * lane;  // 32 lanes x 128 B = 4 KiB
                if (na + 128 * lane < chars_limit) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
        }
        stage::copy_in(sm.in[warp], col.chars, a64, in_b, chars_limit, lane);
        const int w0 = (a64 - S.base) >> 6, nwords = ((in_b - 1 - a64) >> 6) + 1;
        for (int w = lane; w < nwords; w += 32) {
            sm.m[warp][w] = __ldg(S.m + w0 + w);
            sm.k[warp][w] = __ldg(S.k + w0 + w);
            sm.a[warp][w] = __ldg(S.a + w0 + w);
            sm.drop[warp][w] = 0;
            sm.ins[warp][w] = 0;
        }
        __syncwarp();
        // (2) matches of my row -> DROP / INS bits
        if (i < r1 && hits[i]) {
            const spans::Streams T{sm.m[warp] - w0, sm.k[warp] - w0, sm.a[warp] - w0, S.base};
            spans::walk_spans<false>(T, (const uint8_t*)sm.in[warp] - a64, col.offsets[i], col.offsets[i + 1], k_chars, budget, [&](int s, int e) {
                set_bits((uint32_t*)sm.drop[warp], s - a64, e - a64);
                atomicOr((uint32_t*)sm.ins[warp] + ((s - a64) >> 5), 1u << ((s - a64) & 31));
            });
        }
        __syncwarp();
        // (3) cooperative rewrite, lane = one 64-byte word
        const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(sm.out[warp]) + (uint32_t)(out_a & 15);
        int running = 0;
        for (int base_w = 0; base_w < nwords; base_w += 32) {
            const int j = base_w + lane;
            u64 keep = 0, insb = 0;
            if (j < nwords) {
                const int wp = a64 + 64 * j;  // chars offset of bit 0 of this word
                u64 own = ~0ull;
                if (wp < in_a) own &= ~0ull << (in_a - wp);
                if (wp + 64 > in_b) own &= ~0ull >> (wp + 64 - in_b);
                keep = own & ~sm.drop[warp][j];
                insb = own & sm.ins[warp][j];
            }
            const int cnt = __popcll(keep) + repl_len * __popcll(insb);
            int pre = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += v;
            }
            const int tot = __shfl_sync(0xffffffffu, pre, 31);
            const uint32_t o0 = tile_s + (uint32_t)(running + pre - cnt);  // shared address of this word's first output byte
            if (keep) {  // kept bytes: straight-line, one predicated byte store per input byte; insertions only move the cursor
                const uint32_t* w32 = (const uint32_t*)(sm.in[warp] + 64 * j);
                uint32_t o = o0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t kq = (uint32_t)(keep >> (4 * q)) & 15u, iq = (uint32_t)(insb >> (4 * q)) & 15u;
                    const uint32_t v = w32[q];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        o += ((iq >> b) & 1u) * (uint32_t)repl_len;
                        if (kq & (1u << b)) {
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(o), "r"(v >> (8 * b)) : "memory");
                            ++o;
                        }
                    }
                }
            }
            for (u64 t = insb; t;) {  // the replacement strings (a few per word)
                const int bit = __ffsll((long long)t) - 1;
                t &= t - 1;
                const u64 below = (1ull << bit) - 1ull;
                uint32_t o = o0 + (uint32_t)(__popcll(keep & below) + repl_len * __popcll(insb & below));
                for (int r = 0; r < repl_len; ++r) asm volatile("st.shared.u8 [%0], %1;" ::"r"(o + r), "r"((uint32_t)(uint8_t)repl[r]) : "memory");
            }
            running += tot;
        }
        __syncwarp();
        // (4) tile -> output
        stage::copy_out(sm.out[warp], out_chars, out_a, out_b, lane);
        __syncwarp();
    }
}

// Runs the chain kernel with span streams for `c` over `col`.  False: not applicable (caller uses the scalar / VM path).
struct SpanRun {
    bits::SpanStreams ss;
    BufPtr hits, keep_rows, keep_count;
    unsigned int* dirty_count = nullptr;
    spans::Streams view() const { return spans::Streams{ss.m, ss.k, ss.a, ss.base}; }
};
// (r.ss.counts_out set: count mode — per-row match counts straight from the chain kernel, `total` = rows with a match)
static bool run_span_streams(Compiled& c, const custr_column* col, SpanRun& r, unsigned long long* total = nullptr)
{
    if (!span_plan(c) || !cap_tier((int)c.prog.insts.size())) return false;
    if (r.ss.counts_out && !bits::count_in_kernel_ok(*c.plan_contains)) return false;
    r.hits = dev_alloc((size_t)col->n);
    Scratch<unsigned long long> unused(1);
    CUSTR_CUDA(cudaMemsetAsync(unused.get(), 0, 8, g_stream));
    int32_t* dirty_rows = nullptr;
    const bool ok = bits::run(*c.plan_contains, col, (const uint8_t*)c.dev_image->ptr, device_unicode_flags(), (uint8_t*)r.hits->ptr,
                              total ? total : unused.get(), &dirty_rows, &r.dirty_count, r.keep_rows, r.keep_count, &r.ss);
    return ok;  // `unused` is freed stream-ordered
}
static unsigned int read_dirty(const SpanRun& r)
{
    unsigned int h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, r.dirty_count, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h;
}

// ---- findall / extract (capture-span callers, reference findall.cu:36-98, extract.cu:36-68) ---------------------------
// walks the non-overlapping matches of a row exactly like count_re does and calls f(k, begin, end)
template <int CAP, typename F>
__device__ __forceinline__ int walk_matches(const rxdev::DevProg& P, const uint8_t* s, int n, rxdev::Lists<CAP>& L, int limit, F f)
{
    int k = 0, begin = 0;
    while (begin <= n && k < limit) {
        int mb = 0, me = 0;
        if (!rxdev::vm_find<CAP>(P, s, n, begin, n, mb, me, L)) break;
        f(k, mb, me);
        ++k;
        if (me > mb) begin = me;
        else begin = mb + (mb < n ? utf8_width(s[mb]) : 1);
    }
    return k;
}

// pass 0: column-major lengths; pass 1: column-major copy; pass 2: row-major lengths; pass 3: row-major copy
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_findall(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags, int pass,
             const int32_t* __restrict__ counts_or_rowoff, int ncols, int32_t* __restrict__ lens, uint8_t* __restrict__ valid,
             const ColumnOut* __restrict__ outs, const int32_t* __restrict__ tok_off, char* __restrict__ flat_out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    L.init();
    const size_t n = col.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const int len = col.offsets[i + 1] - col.offsets[i];
        if (pass < 2) {
            const int cnt = counts_or_rowoff[i];
            if (pass == 0) {
                for (int k = 0; k < ncols; ++k) { lens[(size_t)k * (n + 1) + i] = 0; valid[(size_t)k * n + i] = k < cnt; }
                if (cnt > 0) walk_matches<CAP>(P, s, len, L, cnt, [&](int k, int b, int e) { lens[(size_t)k * (n + 1) + i] = e - b; });
            } else if (cnt > 0)
                walk_matches<CAP>(P, s, len, L, cnt, [&](int k, int b, int e) {
                    char* o = outs[k].chars + outs[k].offsets[i];
                    for (int j = b; j < e; ++j) *o++ = (char)s[j];
                });
        } else {
            const int first = counts_or_rowoff[i], cnt = counts_or_rowoff[i + 1] - first;
            if (cnt <= 0) continue;
            if (pass == 2) walk_matches<CAP>(P, s, len, L, cnt, [&](int k, int b, int e) { lens[first + k] = e - b; });
            else
                walk_matches<CAP>(P, s, len, L, cnt, [&](int k, int b, int e) {
                    char* o = flat_out + tok_off[first + k];
                    for (int j = b; j < e; ++j) *o++ = (char)s[j];
                });
        }
    }
}

// extract: first match of the row, then one anchored re-run per capture group (reference extract.cu:50-58)
template <int CAP>
__global__ void __launch_bounds__(VM_THREADS)
k_vm_extract(ColView col, const uint8_t* __restrict__ img, int img_bytes, const uint8_t* __restrict__ uflags, int pass, int ngroups,
             int32_t* __restrict__ lens, uint8_t* __restrict__ valid, const ColumnOut* __restrict__ outs)
{
    extern __shared__ __align__(16) uint8_t smem[];
    rxdev::DevProg P = rxdev::bind_program(stage_program(img, img_bytes, smem), uflags);
    rxdev::Lists<CAP> L;
    rxdev::Lists<CAP, true> LG;
    L.init();
    LG.init();
    const size_t n = col.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const int len = col.offsets[i + 1] - col.offsets[i];
        int mb = 0, me = 0;
        const bool hit = col.valid(i) && rxdev::vm_find<CAP>(P, s, len, 0, len, mb, me, L);
        for (int g = 0; g < ngroups; ++g) {
            int gb = -1, ge = -1;
            bool ok = hit && rxdev::vm_find<CAP, true>(P, s, len, mb, mb + 1, gb, ge, LG, g + 1) && ge > gb && gb >= 0;
            if (pass == 0) {
                lens[(size_t)g * (n + 1) + i] = ok ? ge - gb : 0;
                valid[(size_t)g * n + i] = ok;
            } else if (ok) {
                char* o = outs[g].chars + outs[g].offsets[i];
                for (int j = gb; j < ge; ++j) *o++ = (char)s[j];
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------
static inline int vm_grid(int n)
{
    int want = (n + VM_THREADS - 1) / VM_THREADS;
    int cap = num_sms() * 16;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

// Programs beyond the largest in-thread tier: the launch is bracketed by an arena of global scratch (one slice per resident
// thread), the grid is capped so that the arena stays below 4 GiB, and the call is serialised (the arena descriptor is a
// device global).  The reference sizes the same scratch per ROW (regexec.cpp:81-95).
struct VmArenaLaunch {
    BufPtr buf;
    std::unique_lock<std::mutex> lock;
    int nblocks = 0;
};
static VmArenaLaunch vm_arena_for(int ninsts, int want_blocks)
{
    static std::mutex mu;
    VmArenaLaunch a;
    a.lock = std::unique_lock<std::mutex>(mu);
    const size_t per = (rxdev::vm_arena_bytes_per_thread(ninsts) + 15) & ~(size_t)15;
    size_t threads = (size_t)want_blocks * VM_THREADS;
    const size_t cap = ((size_t)4 << 30) / per;
    if (threads > cap) threads = cap < VM_THREADS ? VM_THREADS : cap / VM_THREADS * VM_THREADS;
    a.nblocks = (int)(threads / VM_THREADS);
    a.buf = dev_alloc(per * threads);
    rxdev::VmArena h{(uint8_t*)a.buf->ptr, per, ninsts};
    CUSTR_CUDA(cudaMemcpyToSymbolAsync(rxdev::g_vm_arena, &h, sizeof(h), 0, cudaMemcpyHostToDevice, g_stream));
    return a;
}
#define DISPATCH_CAP(cap, KERNEL, grid, smem, ...)                                         \
    do {                                                                                   \
        if ((cap) == 32) LAUNCH(KERNEL<32>, grid, VM_THREADS, smem, __VA_ARGS__);          \
        else if ((cap) == 256) LAUNCH(KERNEL<256>, grid, VM_THREADS, smem, __VA_ARGS__);   \
        else if ((cap) == 1024) LAUNCH(KERNEL<1024>, grid, VM_THREADS, smem, __VA_ARGS__); \
        else {                                                                             \
            VmArenaLaunch arena__ = vm_arena_for(g_dispatch_ninsts, grid);                 \
            LAUNCH(KERNEL<0>, arena__.nblocks, VM_THREADS, smem, __VA_ARGS__);                \
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));                                   \
        }                                                                                  \
    } while (0)
static thread_local int g_dispatch_ninsts = 0;  // instruction count of the program being dispatched (set by check_cap)

static int check_cap(const Compiled& c, const char* who)
{
    const int n = (int)c.prog.insts.size();
    if (n > VM_MAX_INSTS)  // the reference's wording (count.cu:76-84), at a far larger limit
        throw ArgError{fail(CUSTR_ERR_INVALID, std::string(who) + ": number of instructions (" + std::to_string(n) + ") exceeds available memory")};
    g_dispatch_ninsts = n;
    return cap_tier(n);
}

static int smem_for(const Compiled& c) { return (int)c.image.size() <= SMEM_PROG_MAX ? (int)((c.image.size() + 15) & ~15ull) : 0; }

static unsigned long long read_counter(unsigned long long* d)
{
    unsigned long long h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h;
}

// ---- top-level alternation of chains --------------------------------------------------------------------------------
// `a|b|c` exists in a row iff one of the alternatives does (existence does not depend on thread priority), so when the
// whole pattern is not a chain but every top-level alternative is, each alternative runs on the chain kernel (~0.45 ms per
// GiB) and the boolean results are OR-ed — instead of the DAG interpreter kernel (~3.4 ms per GiB).
static bool split_top_level(const char* pattern, std::vector<std::string>& parts)
{
    parts.clear();
    std::string cur;
    int depth = 0;
    for (const char* p = pattern; *p; ++p) {
        const char ch = *p;
        if (ch == '\\') {
            if (p[1] >= '0' && p[1] <= '9') return false;  // octal escapes swallow the following character: do not reason about them
            cur.push_back(ch);
            if (p[1]) cur.push_back(*++p);
            continue;
        }
        if (ch == '[') {  // copy the class verbatim
            cur.push_back(ch);
            ++p;
            if (*p == '^') cur.push_back(*p++);
            if (*p == ']') cur.push_back(*p++);
            for (; *p && *p != ']'; ++p) {
                cur.push_back(*p);
                if (*p == '\\' && p[1]) cur.push_back(*++p);
            }
            if (!*p) return false;
            cur.push_back(']');
            continue;
        }
        if (ch == '(') ++depth;
        if (ch == ')') --depth;
        if (depth < 0) return false;
        if (ch == '|' && depth == 0) {
            if (cur.empty()) return false;
            parts.push_back(cur);
            cur.clear();
            continue;
        }
        cur.push_back(ch);
    }
    if (depth != 0 || cur.empty()) return false;
    parts.push_back(cur);
    return parts.size() >= 2 && parts.size() <= 8;
}

__global__ void k_or_results(const uint8_t* const* __restrict__ parts, int nparts, int n, uint8_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    unsigned cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint8_t v = 0;
        for (int k = 0; k < nparts; ++k) v |= parts[k][i];
        out[i] = v;
        cnt += v;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

static const std::shared_ptr<bits::Plan>& plan_of(Compiled& c, bool anchored)
{
    if (!c.plans_built) {
        c.plan_contains = bits::lower(c.prog, false, host_unicode_flags());
        c.plan_match = bits::lower(c.prog, true, host_unicode_flags());
        c.plans_built = true;
    }
    return anchored ? c.plan_match : c.plan_contains;
}

// true: out_dev / total hold the answer for every row without a NUL byte; the rows with one are in dirty_rows (VM, full pattern)
static bool alternation_of_chains(const custr_column* col, const char* pattern, bool anchored, uint8_t* out_dev, unsigned long long* total,
                                  int32_t** dirty_rows, unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count)
{
    std::vector<std::string> alts;
    if (bits::g_force_generic || !split_top_level(pattern, alts)) return false;
    std::vector<CompiledPtr> progs;
    for (const std::string& a : alts) {
        progs.push_back(get_compiled(a.c_str()));
        const std::shared_ptr<bits::Plan>& pl = plan_of(*progs.back(), anchored);
        if (!pl || !bits::plan_is_chain(*pl) || !cap_tier((int)progs.back()->prog.insts.size())) return false;
    }
    const int32_t n = col->n;
    std::vector<BufPtr> outs;
    std::vector<const uint8_t*> ptrs;
    Scratch<unsigned long long> unused(1);
    CUSTR_CUDA(cudaMemsetAsync(unused.get(), 0, 8, g_stream));
    for (size_t k = 0; k < progs.size(); ++k) {
        outs.push_back(dev_alloc((size_t)n));
        ptrs.push_back((const uint8_t*)outs.back()->ptr);
        int32_t* dr = nullptr;
        unsigned int* dc = nullptr;
        BufPtr kr, kc;
        if (!bits::run(*plan_of(*progs[k], anchored), col, (const uint8_t*)progs[k]->dev_image->ptr, device_unicode_flags(), (uint8_t*)outs.back()->ptr,
                       unused.get(), &dr, &dc, kr, kc))
            return false;
        if (k == 0) { *dirty_rows = dr; *dirty_count = dc; keep_rows = kr; keep_count = kc; }  // the NUL rows are the same for every alternative
    }
    BufPtr d_ptrs = upload(ptrs.data(), sizeof(void*) * ptrs.size());
    LAUNCH(k_or_results, num_sms() * 8, 256, 0, (const uint8_t* const*)d_ptrs->ptr, (int)ptrs.size(), n, out_dev, total);
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));  // the per-alternative buffers die with this scope
    return true;
}

// Literal `contains` through the chain kernel (find.cu): the literal is escaped into a pattern; when it lowers to a chain the
// boolean result of every row WITHOUT a NUL byte is left in out_dev (*total += hits) and the rows holding one come back in
// dirty_rows for the caller's own byte-compare kernel (a literal search has plain byte semantics there, unlike the regex VM).
bool literal_contains_chain(const custr_column* col, const char* literal, uint8_t* out_dev, unsigned long long* total, int32_t** dirty_rows,
                            unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count)
{
    if (g_forced_tier == 1 || bits::g_force_generic) return false;
    std::string pat;
    for (const char* p = literal; *p; ++p) {
        if (strchr("\\.[](){}*+?|^$", *p)) pat.push_back('\\');
        pat.push_back(*p);
    }
    if (pat.empty()) return false;
    CompiledPtr c = get_compiled(pat.c_str());
    const std::shared_ptr<bits::Plan>& plan = plan_of(*c, false);
    if (!plan || !bits::plan_is_chain(*plan) || !cap_tier((int)c->prog.insts.size())) return false;
    return bits::run(*plan, col, (const uint8_t*)c->dev_image->ptr, device_unicode_flags(), out_dev, total, dirty_rows, dirty_count, keep_rows,
                     keep_count);
}

static int bool_search(const custr_column* col, const char* pattern, uint8_t* results, int devmem, bool anchored, const char* who)
{
    if (!col || !pattern || !results) return fail(CUSTR_ERR_ARG, std::string(who) + ": null argument");
    int32_t n = col->n;
    if (n == 0) return 0;
    CompiledPtr c = get_compiled(pattern);
    ResultBuf<uint8_t> out(results, n, devmem);
    // match total of the paths that do not bring their own counter block (allocated and cleared on first use)
    struct LazyTotal {
        std::unique_ptr<Scratch<unsigned long long>> s;
        unsigned long long* get()
        {
            if (!s) {
                s.reset(new Scratch<unsigned long long>(1));
                CUSTR_CUDA(cudaMemsetAsync(s->get(), 0, 8, g_stream));
            }
            return s->get();
        }
    } total;
    bool done = false;
    if (g_forced_tier != 1) {
        if (!c->plans_built) {
            c->plan_contains = bits::lower(c->prog, false, host_unicode_flags());
            c->plan_match = bits::lower(c->prog, true, host_unicode_flags());
            c->plans_built = true;
        }
        const std::shared_ptr<bits::Plan>& plan = anchored ? c->plan_match : c->plan_contains;
        int cap0 = cap_tier((int)c->prog.insts.size());
        if (cap0 && (!plan || !bits::plan_is_chain(*plan))) {  // not a chain: maybe a top-level alternation of chains
            int32_t* dirty_rows = nullptr;
            unsigned int* dirty_count = nullptr;
            BufPtr keep_rows, keep_count;
            g_timer.start();
            if (alternation_of_chains(col, pattern, anchored, out.dev, total.get(), &dirty_rows, &dirty_count, keep_rows, keep_count)) {
                int grid = vm_grid(n < 1 << 20 ? n : 1 << 20);
                DISPATCH_CAP(cap0, k_vm_bool_rows, grid, smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr, (int)c->image.size(),
                             device_unicode_flags(), anchored ? 1 : 0, (const int32_t*)dirty_rows, (const unsigned int*)dirty_count, out.dev,
                             total.get());
                g_timer.stop();
                g_last_tier = "bitstream";
                int matches = (int)read_counter(total.get());
                g_timer.collect();
                out.finish();
                return matches;
            }
            g_timer.armed = false;
        }
        if (plan) {
            int cap = cap_tier((int)c->prog.insts.size());
            int32_t* dirty_rows = nullptr;
            unsigned int* dirty_count = nullptr;
            // one zeroed 16-byte block: [match total u64][dirty-row count u32][work-item counter u32]
            BufPtr keep_rows, keep_count = dev_alloc(16);
            CUSTR_CUDA(cudaMemsetAsync(keep_count->ptr, 0, 16, g_stream));
            unsigned long long* d_total = (unsigned long long*)keep_count->ptr;
            g_timer.start();
            if (cap && bits::run(*plan, col, (const uint8_t*)c->dev_image->ptr, device_unicode_flags(), out.dev, d_total, &dirty_rows,
                                 &dirty_count, keep_rows, keep_count)) {
                g_timer.stop();
                g_last_tier = "bitstream";
                done = true;
                struct Counters { unsigned long long total; unsigned int dirty, items; } h{};
                {   // the one read-back of the call lands in page-locked memory (a pageable target makes the driver stage the copy)
                    static thread_local Counters* pinned = nullptr;
                    if (!pinned && cudaHostAlloc((void**)&pinned, sizeof(Counters), cudaHostAllocDefault) != cudaSuccess) {
                        cudaGetLastError();
                        pinned = nullptr;
                    }
                    Counters* dst = pinned ? pinned : &h;
                    CUSTR_CUDA(cudaMemcpyAsync(dst, keep_count->ptr, 16, cudaMemcpyDeviceToHost, g_stream));
                    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                    h = *dst;
                }
                if (h.dirty) {  // rows holding a NUL byte: exact VM over the work list (rare: no launch at all otherwise)
                    int grid = vm_grid(n < 1 << 20 ? n : 1 << 20);
                    DISPATCH_CAP(cap, k_vm_bool_rows, grid, smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                                 (int)c->image.size(), device_unicode_flags(), anchored ? 1 : 0, (const int32_t*)dirty_rows,
                                 (const unsigned int*)dirty_count, out.dev, d_total);
                    h.total = read_counter(d_total);
                }
                g_timer.collect();
                out.finish();
                return (int)h.total;
            }
            g_timer.armed = false;
        }
    }
    if (!done) {
        int cap = check_cap(*c, who);
        g_timer.start();
        DISPATCH_CAP(cap, k_vm_bool, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                     (int)c->image.size(), device_unicode_flags(), anchored ? 1 : 0, out.dev, total.get());
        g_timer.stop();
        g_last_tier = "pikevm";
    }
    int matches = (int)read_counter(total.get());
    g_timer.collect();
    out.finish();
    return matches;
}

static custr_column* finish_replace(const custr_column* col, Scratch<int32_t>& lens, BufPtr& off, int64_t& total)
{
    off = dev_alloc(sizeof(int32_t) * (size_t)(col->n + 1));
    total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, col->n);
    if (total > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "replace: result exceeds 2 GiB of chars (int32 offsets)")};
    return nullptr;
}

static BufPtr copy_validity(const custr_column* col)
{
    if (!col->validity) return nullptr;
    BufPtr v = dev_alloc((col->n + 7) / 8);
    // re-aligned copy through the export helper
    custr_create_offsets(col, nullptr, nullptr, (uint8_t*)v->ptr, 1);
    return v;
}


struct FindallCtx { const custr_column* col; Compiled* c; int cap; const int32_t* counts; int ncols; };
static void findall_copy(const ColumnOut* d_outs, void* vctx)
{
    FindallCtx* x = (FindallCtx*)vctx;
    DISPATCH_CAP(x->cap, k_vm_findall, vm_grid(x->col->n), smem_for(*x->c), view_of(x->col), (const uint8_t*)x->c->dev_image->ptr,
                 (int)x->c->image.size(), device_unicode_flags(), 1, x->counts, x->ncols, (int32_t*)nullptr, (uint8_t*)nullptr, d_outs,
                 (const int32_t*)nullptr, (char*)nullptr);
}
struct ExtractCtx { const custr_column* col; Compiled* c; int cap; int ngroups; };
static void extract_copy(const ColumnOut* d_outs, void* vctx)
{
    ExtractCtx* x = (ExtractCtx*)vctx;
    DISPATCH_CAP(x->cap, k_vm_extract, vm_grid(x->col->n), smem_for(*x->c), view_of(x->col), (const uint8_t*)x->c->dev_image->ptr,
                 (int)x->c->image.size(), device_unicode_flags(), 1, x->ngroups, (int32_t*)nullptr, (uint8_t*)nullptr, d_outs);
}

static int max_count(const int32_t* d, int n)
{
    Scratch<int32_t> out(1);
    size_t tmp_bytes = 0;
    cub::DeviceReduce::Max(nullptr, tmp_bytes, d, out.get(), n, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceReduce::Max(tmp->ptr, tmp_bytes, d, out.get(), n, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    int32_t h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, out.get(), 4, cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h;
}

}  // namespace custr

using namespace custr;

extern "C" {

const char* custr_last_regex_tier(void) { return g_last_tier; }
void custr_set_profiling(int on) { g_profile = on; }
float custr_last_kernel_ms(void) { return g_last_kernel_ms; }
// A/B: size of a work item of the chain / tokenize kernels in KiB (default 32)
void custr_set_item_kib(int kib)
{
    bits::g_item_bytes = (kib >= 4 && kib <= 32) ? kib * 1024 : 0;
    bits::g_item_stagger = kib != -1;  // -1: default sizes without the graded first round (A/B)
}
// run-time compiled plan kernels (regex_jit.cu): mode 0 never, 1 = large columns whose plan has no ahead-of-time shape
// (default), 2 = always; min_bytes > 0 also sets the column size from which mode 1 compiles
void custr_set_jit(int mode, long long min_bytes)
{
    bits::g_jit_mode = mode < 0 || mode > 2 ? 1 : mode;
    if (min_bytes > 0) bits::g_jit_min_bytes = min_bytes;
}
long long custr_jit_launch_count(void) { return bits::g_jit_launches.load(); }
const char* custr_jit_note(void) { return bits::jit_last_note(); }
void custr_set_regex_tier(int tier)
{
    g_forced_tier = tier == 1 ? 1 : 0;
    bits::g_force_generic = tier == 2;  // 2: bitstream tier, generic DAG interpreter even for chain-shaped plans
    bits::g_chain_win = tier == 3;      // 3: bitstream tier, window-at-a-time chain kernel (k_chain64) also for boolean results
    bits::g_no_spec = tier == 4 || tier == 5;  // 4: no ahead-of-time shape specialisation (the plan's run-time compiled kernel instead, when available)
    bits::g_no_jit = tier == 5;                // 5: neither: the generic ahead-of-time kernels that interpret the plan
}

int custr_regex_describe(const char* pattern, char* buf, size_t buflen)
{
    if (!pattern) return CUSTR_ERR_ARG;
    rx::Program p = rx::compile(pattern);
    std::string d = p.describe();
    std::shared_ptr<bits::Plan> plan = bits::lower(p, false, host_unicode_flags());
    d += plan ? "bitstream: " + bits::describe(*plan) + "\n" : "bitstream: not eligible\n";
    if (buf && buflen) {
        size_t k = d.size() < buflen - 1 ? d.size() : buflen - 1;
        memcpy(buf, d.data(), k);
        buf[k] = 0;
    }
    return (int)p.insts.size();
}

int custr_contains_re(const custr_column* col, const char* pattern, uint8_t* results, int devmem)
{
    return guarded([&] { return bool_search(col, pattern, results, devmem, false, "contains_re"); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}

int custr_match(const custr_column* col, const char* pattern, uint8_t* results, int devmem)
{
    return guarded([&] { return bool_search(col, pattern, results, devmem, true, "match"); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}

int custr_count_re(const custr_column* col, const char* pattern, int32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !pattern || !results) return fail(CUSTR_ERR_ARG, "count_re: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            CompiledPtr c = get_compiled(pattern);
            int cap = check_cap(*c, "count_re");
            ResultBuf<int32_t> out(results, n, devmem);
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            if (const bits::ChainDev* cd = span_plan(*c)) {  // last-loop chain
                {   // every step on the loop's class: the chain kernel counts by itself (k_chain64 count mode)
                    SpanRun cr;
                    cr.ss.counts_out = out.dev;
                    if (run_span_streams(*c, col, cr, total.get())) {
                        int matches = (int)read_counter(total.get());
                        if (read_dirty(cr) == 0) {
                            g_last_tier = "bitcount";
                            out.finish();
                            return matches;
                        }
                        CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
                    }
                }
                SpanRun sr;
                if (run_span_streams(*c, col, sr)) {  // bit streams + word scans (rows holding NUL: fall through)
                    LAUNCH(k_span_count, vm_grid(n) * 2, 256, 0, view_of(col), sr.view(), (const uint8_t*)sr.hits->ptr, (int)cd->nsteps - 1, out.dev,
                           total.get());
                    int matches = (int)read_counter(total.get());
                    if (read_dirty(sr) == 0) {
                        g_last_tier = "bitspans";
                        out.finish();
                        return matches;
                    }
                    CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
                }
                Scratch<int> nul_seen(1);
                CUSTR_CUDA(cudaMemsetAsync(nul_seen.get(), 0, sizeof(int), g_stream));
                LAUNCH(k_chain_count, vm_grid(n) * 2, 256, 0, view_of(col), *cd, device_unicode_flags(), out.dev, total.get(), nul_seen.get());
                int matches = (int)read_counter(total.get());
                int h_nul = 0;
                CUSTR_CUDA(cudaMemcpyAsync(&h_nul, nul_seen.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
                CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                if (!h_nul) {
                    g_last_tier = "chainspan";
                    out.finish();
                    return matches;
                }
                CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));  // a row holds a NUL byte: redo with the exact VM
            }
            DISPATCH_CAP(cap, k_vm_count, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                         (int)c->image.size(), device_unicode_flags(), out.dev, total.get());
            g_last_tier = "pikevm";
            int matches = (int)read_counter(total.get());
            out.finish();
            return matches;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}


int custr_findall(const custr_column* col, const char* pattern, custr_column** out, int32_t cap)
{
    return guarded(
        [&]() -> int {
            if (!col || !pattern) return fail(CUSTR_ERR_ARG, "findall: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            CompiledPtr c = get_compiled(pattern);
            int vcap = check_cap(*c, "findall");
            Scratch<int32_t> counts((size_t)n + 1);
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            DISPATCH_CAP(vcap, k_vm_count, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr, (int)c->image.size(),
                         device_unicode_flags(), counts.get(), total.get());
            int ncols = max_count(counts.get(), n);
            if (ncols == 0) {  // no match anywhere: one all-null column (findall.cu:131-132)
                if (cap > 0) out[0] = all_null_column(n);
                return 1;
            }
            Scratch<int32_t> lens((size_t)ncols * (n + 1));
            Scratch<uint8_t> valid((size_t)ncols * n);
            DISPATCH_CAP(vcap, k_vm_findall, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                         (int)c->image.size(), device_unicode_flags(), 0, (const int32_t*)counts.get(), ncols, lens.get(), valid.get(),
                         (const ColumnOut*)nullptr, (const int32_t*)nullptr, (char*)nullptr);
            FindallCtx ctx{col, c.get(), vcap, counts.get(), ncols};
            std::vector<custr_column*> cols = assemble_columns(n, ncols, lens.get(), valid.get(), findall_copy, &ctx);
            for (int k = 0; k < ncols; ++k) { if (k < cap) out[k] = cols[k]; else custr_column_free(cols[k]); }
            g_last_tier = "pikevm";
            return ncols;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_findall_record(const custr_column* col, const char* pattern, custr_column** tokens, int32_t* row_offsets, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !pattern || !tokens) return fail(CUSTR_ERR_ARG, "findall_record: null argument");
            int32_t n = col->n;
            CompiledPtr c = get_compiled(pattern);
            int vcap = check_cap(*c, "findall_record");
            Scratch<int32_t> counts((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(counts.get() + n, 0, 4, g_stream));
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            if (n) DISPATCH_CAP(vcap, k_vm_count, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                                (int)c->image.size(), device_unicode_flags(), counts.get(), total.get());
            Scratch<int32_t> row_off((size_t)n + 1);
            int64_t ntok = scan_lengths_to_offsets(counts.get(), row_off.get(), n);
            Scratch<int32_t> tlens((size_t)ntok + 1);
            CUSTR_CUDA(cudaMemsetAsync(tlens.get() + ntok, 0, 4, g_stream));
            if (ntok) DISPATCH_CAP(vcap, k_vm_findall, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                                   (int)c->image.size(), device_unicode_flags(), 2, (const int32_t*)row_off.get(), 0, tlens.get(),
                                   (uint8_t*)nullptr, (const ColumnOut*)nullptr, (const int32_t*)nullptr, (char*)nullptr);
            BufPtr tok_off = dev_alloc(sizeof(int32_t) * (size_t)(ntok + 1));
            int64_t bytes = scan_lengths_to_offsets(tlens.get(), (int32_t*)tok_off->ptr, (int32_t)ntok);
            BufPtr chars = dev_alloc((size_t)bytes);
            if (ntok) DISPATCH_CAP(vcap, k_vm_findall, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                                   (int)c->image.size(), device_unicode_flags(), 3, (const int32_t*)row_off.get(), 0, (int32_t*)nullptr,
                                   (uint8_t*)nullptr, (const ColumnOut*)nullptr, (const int32_t*)tok_off->ptr, (char*)chars->ptr);
            if (row_offsets) {
                cudaMemcpyKind kind = devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
                CUSTR_CUDA(cudaMemcpyAsync(row_offsets, row_off.get(), sizeof(int32_t) * (size_t)(n + 1), kind, g_stream));
            }
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            *tokens = make_column(chars, tok_off, nullptr, (int32_t)ntok, 0, bytes);
            g_last_tier = "pikevm";
            return (int)ntok;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_extract(const custr_column* col, const char* pattern, custr_column** out, int32_t cap)
{
    return guarded(
        [&]() -> int {
            if (!col || !pattern) return fail(CUSTR_ERR_ARG, "extract: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            CompiledPtr c = get_compiled(pattern);
            int vcap = check_cap(*c, "extract");
            int groups = c->prog.ngroups;
            if (groups == 0) return 0;
            Scratch<int32_t> lens((size_t)groups * (n + 1));
            Scratch<uint8_t> valid((size_t)groups * n);
            DISPATCH_CAP(vcap, k_vm_extract, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                         (int)c->image.size(), device_unicode_flags(), 0, groups, lens.get(), valid.get(), (const ColumnOut*)nullptr);
            ExtractCtx ctx{col, c.get(), vcap, groups};
            std::vector<custr_column*> cols = assemble_columns(n, groups, lens.get(), valid.get(), extract_copy, &ctx);
            for (int k = 0; k < groups; ++k) { if (k < cap) out[k] = cols[k]; else custr_column_free(cols[k]); }
            g_last_tier = "pikevm";
            return groups;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_replace_re(const custr_column* col, const char* pattern, const char* repl, int32_t maxrepl)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "replace_re: null column")};
            if (!pattern || !*pattern) throw ArgError{fail(CUSTR_ERR_INVALID, "nvstrings::replace_re: pattern parameter cannot be null or empty")};
            if (!repl) repl = "";
            int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            CompiledPtr c = get_compiled(pattern);
            std::string lit;
            if (g_forced_tier != 1 && pure_literal(c->prog, lit)) {  // literal pattern: the find-based kernel is exact
                custr_column* r = custr_replace(col, lit.c_str(), repl, maxrepl);
                if (!r) throw ArgError{CUSTR_ERR_INVALID};
                g_last_tier = "literal";
                return r;
            }
            int cap = check_cap(*c, "replace_re");
            int repl_len = (int)strlen(repl);
            BufPtr d_repl = upload(repl, repl_len ? repl_len : 1);
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            const bits::ChainDev* cd = span_plan(*c);
            Scratch<int> nul_seen(1);
            int h_nul = 0;
            if (cd) {  // last-loop chain, bit streams + word scans
                SpanRun sr;
                trace_point("replace_re: entry");
                if (run_span_streams(*c, col, sr)) {
                    trace_point("replace_re: chain + spans");
                    const int k_chars = (int)cd->nsteps - 1;
                    if (maxrepl < 0 && !bits::g_chain_win && bits::replace_spans_ok(*cd) && read_dirty(sr) == 0) {
                        // single-class chain, every match: streaming splice over the span streams, no per-row walk
                        BufPtr chars2, off2;
                        int64_t total2 = 0;
                        if (bits::replace_spans_flat(col, *cd, sr.ss, repl, repl_len, chars2, off2, total2)) {
                            g_last_tier = "bitsplice";
                            trace_point("replace_re: splice done.");
                            return make_column(chars2, off2, copy_validity(col), n, col->nulls, total2);
                        }
                    }
                    LAUNCH(k_span_replace, vm_grid(n) * 2, 256, 0, view_of(col), sr.view(), (const uint8_t*)sr.hits->ptr, k_chars, repl_len, maxrepl,
                           lens.get());
                    if (read_dirty(sr) == 0) {
                        BufPtr off2;
                        int64_t total2 = 0;
                        finish_replace(col, lens, off2, total2);
                        BufPtr chars2 = dev_alloc((size_t)total2);
                        LAUNCH(k_span_replace_splice, stage::grid_for(n), splice::THREADS, 0, view_of(col), col->first_off + (int)col->nbytes, sr.view(),
                               (const uint8_t*)sr.hits->ptr, k_chars, (const char*)d_repl->ptr, repl_len, maxrepl, (const int32_t*)off2->ptr,
                               (char*)chars2->ptr);
                        g_last_tier = "bitspans";
                        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                        return make_column(chars2, off2, copy_validity(col), n, col->nulls, total2);
                    }
                }
            }
            if (cd) {  // last-loop chain: scalar leftmost-longest scan is exact (rows holding NUL: whole call goes to the VM)
                CUSTR_CUDA(cudaMemsetAsync(nul_seen.get(), 0, sizeof(int), g_stream));
                LAUNCH(k_chain_replace, vm_grid(n) * 2, 256, 0, view_of(col), *cd, device_unicode_flags(), (const char*)d_repl->ptr, repl_len,
                       maxrepl, lens.get(), (const int32_t*)nullptr, (char*)nullptr, nul_seen.get());
                CUSTR_CUDA(cudaMemcpyAsync(&h_nul, nul_seen.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
                CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            }
            if (cd && !h_nul) {
                BufPtr off2;
                int64_t total2 = 0;
                finish_replace(col, lens, off2, total2);
                BufPtr chars2 = dev_alloc((size_t)total2);
                LAUNCH(k_chain_replace, vm_grid(n) * 2, 256, 0, view_of(col), *cd, device_unicode_flags(), (const char*)d_repl->ptr, repl_len,
                       maxrepl, (int32_t*)nullptr, (const int32_t*)off2->ptr, (char*)chars2->ptr, nul_seen.get());
                g_last_tier = "chainspan";
                CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                return make_column(chars2, off2, copy_validity(col), n, col->nulls, total2);
            }
            DISPATCH_CAP(cap, k_vm_replace, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                         (int)c->image.size(), device_unicode_flags(), (const char*)d_repl->ptr, repl_len, maxrepl, lens.get(),
                         (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off;
            int64_t total = 0;
            finish_replace(col, lens, off, total);
            BufPtr chars = dev_alloc((size_t)total);
            DISPATCH_CAP(cap, k_vm_replace, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr,
                         (int)c->image.size(), device_unicode_flags(), (const char*)d_repl->ptr, repl_len, maxrepl,
                         (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            g_last_tier = "pikevm";
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));  // d_repl / lens die with this scope
            return make_column(chars, off, copy_validity(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_replace_with_backrefs(const custr_column* col, const char* pattern, const char* repl)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "replace_with_backrefs: null column")};
            if (!pattern || !*pattern) throw ArgError{fail(CUSTR_ERR_INVALID, "nvstrings::replace_with_backrefs parameter cannot be null or empty")};
            const int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            if (!repl) return all_null_column(n);  // replace_backref.cu:127-128
            // template parse (regex/backref.h:31-57): every backslash followed by digits is a reference
            std::string tmpl;
            std::vector<int2> refs;
            for (const char* p = repl; *p;) {
                if (*p == '\\' && p[1] >= '0' && p[1] <= '9') {
                    char* endp = nullptr;
                    long idx = strtol(p + 1, &endp, 10);
                    refs.push_back(make_int2((int)(idx > 0x7fffffffL ? 0x7fffffff : idx), (int)tmpl.size()));
                    p = endp;
                } else
                    tmpl.push_back(*p++);
            }
            CompiledPtr c = get_compiled(pattern);
            const int cap = check_cap(*c, "replace_with_backrefs");
            BufPtr d_tmpl = upload(tmpl.data(), tmpl.size() ? tmpl.size() : 1);
            int2 none = make_int2(0, 0);
            BufPtr d_refs = upload(refs.empty() ? &none : refs.data(), sizeof(int2) * (refs.empty() ? 1 : refs.size()));
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            DISPATCH_CAP(cap, k_vm_backrefs, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr, (int)c->image.size(),
                         device_unicode_flags(), (const char*)d_tmpl->ptr, (int)tmpl.size(), (const int2*)d_refs->ptr, (int)refs.size(), lens.get(),
                         (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off;
            int64_t total = 0;
            finish_replace(col, lens, off, total);
            BufPtr chars = dev_alloc((size_t)total);
            DISPATCH_CAP(cap, k_vm_backrefs, vm_grid(n), smem_for(*c), view_of(col), (const uint8_t*)c->dev_image->ptr, (int)c->image.size(),
                         device_unicode_flags(), (const char*)d_tmpl->ptr, (int)tmpl.size(), (const int2*)d_refs->ptr, (int)refs.size(),
                         (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            g_last_tier = "pikevm";
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, copy_validity(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_replace_re_multi(const custr_column* col, const char* const* patterns, int32_t npatterns,
                                     const custr_column* repls)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col || !repls) throw ArgError{fail(CUSTR_ERR_ARG, "replace_re: null column")};
            if (npatterns <= 0 || !patterns) throw ArgError{fail(CUSTR_ERR_INVALID, "replace_re patterns is empty")};
            if (repls->n != npatterns && repls->n != 1)
                throw ArgError{fail(CUSTR_ERR_INVALID, "replace_re patterns and repls must have the same number of strings")};
            int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            std::vector<CompiledPtr> progs;
            std::vector<const uint8_t*> imgs;
            int cap = 32, max_insts = 0;
            for (int t = 0; t < npatterns; ++t) {
                if (!patterns[t]) throw ArgError{fail(CUSTR_ERR_INVALID, "replace_re: null pattern")};
                progs.push_back(get_compiled(patterns[t]));
                check_cap(*progs.back(), "replace_re");
                max_insts = std::max(max_insts, (int)progs.back()->prog.insts.size());
                imgs.push_back((const uint8_t*)progs.back()->dev_image->ptr);
            }
            cap = cap_tier(max_insts);  // one list size for all programs: the largest one's tier
            g_dispatch_ninsts = max_insts;
            BufPtr d_imgs = upload(imgs.data(), imgs.size() * sizeof(void*));
            MultiProgs mp{(const uint8_t* const*)d_imgs->ptr, npatterns};
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            DISPATCH_CAP(cap, k_vm_replace_multi, vm_grid(n), 0, view_of(col), mp, device_unicode_flags(), view_of(repls),
                         lens.get(), (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off;
            int64_t total = 0;
            finish_replace(col, lens, off, total);
            BufPtr chars = dev_alloc((size_t)total);
            DISPATCH_CAP(cap, k_vm_replace_multi, vm_grid(n), 0, view_of(col), mp, device_unicode_flags(), view_of(repls),
                         (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            g_last_tier = "pikevm";
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, copy_validity(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

}  // extern "C"
