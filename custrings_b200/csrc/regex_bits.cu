// Bitstream regex executor (device).  Model, plan format and the equivalence argument: regex_bits_plan.h;
// host lowering: regex_bits_lower.cpp.  DESIGN.md §4.2 has the roofline accounting.
//
// Mapping to the machine
//   * the flat chars buffer is cut into 1024-byte WINDOWS; one warp evaluates one window at a time, lane L owning the
//     32 bytes [32L, 32L+32) = one 32-bit word of every stream.  Loads are two LDG.128 per lane, fully coalesced,
//     software-prefetched one window ahead; no byte is read twice except the two boundary windows of a work item.
//   * the 32 bytes of a lane are bit-transposed in registers (16 PRMT + 48 SHF/LOP3) into 8 bit planes; character
//     classes are boolean formulas over the planes, so classification costs a few LOP3 per 32 bytes instead of a
//     table lookup per byte.
//   * advance / look-ahead are a warp shuffle + funnel shift; `spread` (x+, x*, and the per-row sticky OR) is a
//     1024-bit carry-propagating add done with two ballots; inter-window state is one carry bit per stream.
//   * rows never interact: ROWSTART bits are scattered from the offsets array (shared-memory atomicOr); a work item is a
//     contiguous ROW range whose bytes start near a 32 KiB boundary (warp-cooperative 32-ary search in offsets), so
//     warps never exchange state.
//   * rows holding a byte >= 0x80 or a NUL are appended to a work list and decided by the exact Pike-VM kernel.
#include "regex_bits.h"
#include "regex_bits_plan.h"
#include "regex_vm.cuh"
#include <cub/cub.cuh>
#include <cstddef>
#include <mutex>

namespace custr {
namespace bits {

struct Plan {
    PlanDev dev;
    bool is_chain = false;
    bool span_ok = false;
    bool chain_has_opt = false;
    ChainDev chain;
    std::string text;
};

thread_local bool g_no_spec = false;        // A/B switch: never use the shape-specialised chain kernels
thread_local bool g_force_generic = false;  // A/B switch: run chain-shaped plans on the generic interpreter kernel
bool g_item_stagger = true;               // graded first-round items on large columns (A/B: custr_set_item_kib(-1) turns it off)
int g_item_bytes = 0;                      // forced size of a work item of the chain / tokenize kernels (custr_set_item_kib; 0 = item_bytes_for)
thread_local bool g_chain_win = false;      // A/B switch: boolean results from k_chain64 (window at a time) instead of k_chain_item

#include "regex_bits_dev.cuh"

__device__ __forceinline__ void load_window(const char* __restrict__ chars, int ws, int end, uint4& lo, uint4& hi)
{
    int b = ws + 32 * (int)lane_id();
    if (b + 32 <= end) {
        const uint4* q = (const uint4*)(chars + b);
        lo = __ldg(q);
        hi = __ldg(q + 1);
    } else {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t v = 0;
            for (int k = 0; k < 4; ++k) {
                int pos = b + 4 * j + k;
                if (pos < end && pos >= 0) v |= (uint32_t)(uint8_t)chars[pos] << (8 * k);
            }
            w[j] = v;
        }
        lo = make_uint4(w[0], w[1], w[2], w[3]);
        hi = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

__global__ void __launch_bounds__(THREADS)
k_bitstream(const __grid_constant__ PlanDev plan, const Args A)
{
    __shared__ WarpSmem smem[WARPS];
    WarpSmem& S = smem[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const int warps_total = gridDim.x * WARPS;
    unsigned long long my_matches = 0;

    for (int item = blockIdx.x * WARPS + (threadIdx.x >> 5); item < A.nitems; item += warps_total) {
        // ---- rows of this item: first row starting at or after the item's byte boundary
        const int lo_byte = A.first + item * ITEM_BYTES;
        int ra = item == 0 ? 0 : warp_lower_bound(A.offsets, A.n, lo_byte);
        int rb = item == A.nitems - 1 ? A.n : warp_lower_bound(A.offsets, A.n, lo_byte + ITEM_BYTES);
        if (ra >= rb) continue;
        const int byte_a = __ldg(A.offsets + ra), byte_b = __ldg(A.offsets + rb);
        if (byte_a >= byte_b) continue;  // only empty rows: results stay 0 (pre-cleared)
        int krs = ra;       // next offset index whose ROWSTART bit is not set yet
        int kfin = ra + 1;  // next offset index j whose row j-1 is not finalised yet
        uint32_t carry = 0;
        int ws = byte_a & ~(WIN - 1);
        uint4 cur_lo, cur_hi, nxt_lo, nxt_hi;
        load_window(A.chars, ws, A.end, cur_lo, cur_hi);

        for (; ws < byte_b; ws += WIN) {
            const int we = ws + WIN;
            if (we < byte_b) load_window(A.chars, we, A.end, nxt_lo, nxt_hi);  // prefetch

            // ---- ROWSTART bits from the offsets array
            S.rs[lane] = 0;
            __syncwarp();
            for (;;) {
                int j = krs + (int)lane;
                int o = j <= rb ? __ldg(A.offsets + j) : 0x7fffffff;
                bool in = o < we;
                if (in && o >= ws) atomicOr(&S.rs[(o - ws) >> 5], 1u << ((o - ws) & 31));
                unsigned m = __ballot_sync(FULL, in);
                krs += __popc(m);
                if (m != FULL) break;
            }
            __syncwarp();
            Assertions as;
            as.rs = S.rs[lane];
            const uint32_t nrs = ~as.rs;
            const uint32_t rs_next = (krs <= rb && __ldg(A.offsets + krs) == we) || we >= A.end;
            const uint32_t next_byte = (!rs_next && we < A.end) ? (uint8_t)A.chars[we] : 0;

            // ---- bit planes, dirty bytes
            uint32_t p[8];
            transpose_planes(cur_lo, cur_hi, p);
            const uint32_t dirty_bits = p[7] | ~(p[0] | p[1] | p[2] | p[3] | p[4] | p[5] | p[6] | p[7]);

            // ---- class streams
            for (int k = 0; k < plan.nclasses; ++k) {
                uint32_t v = 0;
                for (int a = 0; a < plan.classes[k].natoms; ++a) v |= cls_atom(p, plan.classes[k].atoms[a]);
                if (plan.classes[k].negate) v = ~v;
                S.cls[k][lane] = v;
            }
            // ---- assertion streams (only what the plan uses)
            uint32_t carry_out = 0;
            as.bow_b = as.bow_a = as.bolc_b = as.nl = as.lb = as.eold_a = 0;
            const uint32_t needs = plan.before_needs | plan.after_needs;
            if (needs & (AS_BOW | AS_NBOW)) {
                uint32_t al = cls_alnum(p);
                as.bow_b = al ^ (advance(al, (carry >> CY_A) & 1u) & nrs);
                carry_out |= top_bit(al) << CY_A;
                uint32_t nb = next_byte | 0x20u;
                uint32_t a_next = (next_byte - '0' < 10u) || (nb - 'a' < 26u);
                as.bow_a = al ^ shift_down(al & nrs, a_next);
            }
            if (needs & (AS_BOL_CARET | AS_EOL_DOLLAR | AS_EOL_Z)) {
                as.nl = cls_eq(p, '\n');
                as.bolc_b = as.rs | (advance(as.nl, (carry >> CY_NL) & 1u) & nrs);
                carry_out |= top_bit(as.nl) << CY_NL;
                as.lb = shift_down(as.rs, rs_next);
                as.eold_a = as.lb | shift_down(as.nl & nrs, next_byte == '\n');
            }
            __syncwarp();

            // ---- marker steps
            uint32_t E = 0;
            const uint32_t start = plan.anchored ? as.rs : 0xffffffffu;
            int end_i = 0;
            for (int s = 0; s < plan.nsteps; ++s) {
                const StepD st = plan.steps[s];
                const uint32_t c = S.cls[st.cls][lane];
                uint32_t entry = 0;
                for (int j = 0; j < st.npreds; ++j) {
                    const PredD pr = st.preds[j];
                    uint32_t t = pr.src == SRC_START ? start : S.slot[pr.src][lane];
                    entry |= apply_before(t, pr.mask, as);
                }
                uint32_t P = entry & c;
                if (st.self_loop) P = spread(P, apply_before(c & nrs, st.self_mask, as), (carry >> s) & 1u);
                while (end_i < plan.nends && plan.ends[end_i].src == s) E |= apply_after(P, plan.ends[end_i++].mask, as);
                S.slot[s][lane] = advance(P, (carry >> s) & 1u) & nrs;
                carry_out |= top_bit(P) << s;
            }

            // ---- sticky per-row OR of matches (and of dirty bytes when there are any)
            uint32_t F = spread(E, nrs, (carry >> CY_F) & 1u);
            carry_out |= top_bit(F) << CY_F;
            S.f[lane] = F;
            const bool any_dirty = __any_sync(FULL, dirty_bits != 0) || ((carry >> CY_D) & 1u);
            if (any_dirty) {
                uint32_t D = spread(dirty_bits, nrs, (carry >> CY_D) & 1u);
                carry_out |= top_bit(D) << CY_D;
                S.d[lane] = D;
            }
            __syncwarp();

            // ---- finalise the rows whose last byte lies in this window
            for (;;) {
                int j = kfin + (int)lane;
                int o = j <= rb ? __ldg(A.offsets + j) : 0x7fffffff;
                bool in = o <= we;
                bool hit = false, dirty = false;
                if (in) {
                    int o_prev = __ldg(A.offsets + j - 1);
                    if (o > o_prev) {  // non-empty row j-1, last byte o-1 >= ws
                        int b = o - 1 - ws;
                        hit = (S.f[b >> 5] >> (b & 31)) & 1u;
                        dirty = any_dirty && ((S.d[b >> 5] >> (b & 31)) & 1u);
                        if (!dirty) A.out[j - 1] = hit;
                    }
                }
                unsigned dm = __ballot_sync(FULL, dirty);
                if (dm) {
                    unsigned basei = 0;
                    if (lane == 0) basei = atomicAdd(A.dirty_count, __popc(dm));
                    basei = __shfl_sync(FULL, basei, 0);
                    if (dirty) A.dirty_rows[basei + __popc(dm & ((1u << lane) - 1))] = j - 1;
                }
                my_matches += __popc(__ballot_sync(FULL, hit && !dirty));
                unsigned m = __ballot_sync(FULL, in);
                kfin += __popc(m);
                if (m != FULL) break;
            }
            __syncwarp();
            carry = carry_out;
            cur_lo = nxt_lo;
            cur_hi = nxt_hi;
        }
    }
    if (lane == 0 && my_matches) atomicAdd(A.total, my_matches);
}


// Byte position (relative to the first char) at which work item t begins.  Plain: t * ib.  Staggered (large columns): the first W
// items — the first item of every resident warp — have graded sizes ib/8, 2 ib/8, .. ib, so that the warps, which all start at
// the same moment, finish their first item at different times and never run their phases in lock-step (row bookkeeping of
// some warps then overlaps the window loop of others from the first round on); behind them every item has ib bytes.
constexpr int ITEM_STAGGER = 8;
__host__ __device__ inline long long item_start(long long t, long long W, long long ib, bool staggered)
{
    if (!staggered) return t * ib;
    const long long unit = ib / ITEM_STAGGER, group = unit * (ITEM_STAGGER * (ITEM_STAGGER + 1) / 2);
    if (t <= W) {
        const long long q = t / ITEM_STAGGER, r = t % ITEM_STAGGER;
        return q * group + unit * (r * (r + 1) / 2);
    }
    return (W / ITEM_STAGGER) * group + (t - W) * ib;
}
// item_bounds[t] = first row r with offsets[r] >= first + item_start(t), for t = 0..nitems (bounds[nitems] = n): one thread per
// boundary, a binary search over the offsets (once per column)
__global__ void k_item_bounds(const int32_t* __restrict__ offsets, int n, int first, int nitems, int item_bytes, int W, int staggered,
                              int32_t* __restrict__ bounds)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nitems) return;
    if (t == nitems) { bounds[t] = n; return; }
    const long long target = (long long)first + item_start(t, W, item_bytes, staggered != 0);
    int lo = 0, hi = n;  // first r in [0, n] with offsets[r] >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)offsets[mid] >= target) hi = mid;
        else lo = mid + 1;
    }
    bounds[t] = lo;
}

// Size of a work item for a column of `nbytes` chars: 32 KiB, except for columns too small to give every resident warp (W = SMs x
// 3 CTAs x 8 warps) one such item — there the item shrinks (down to 4 KiB) so that all warps share the one round: a warp working
// alone is latency-bound and the first item of a launch runs cold (tools/size_sweep.py: 1 MB 0.119 -> 0.035 ms, 34 MB 0.068 ->
// 0.049 ms).  Balancing the rounds of LARGER columns the same way (item = nbytes / (W x rounds)) was measured and is worse
// (537 MB: 0.307 vs 0.223 ms): with every warp finishing its item at the same moment the phases of all warps run in lock-step.
static int item_bytes_for(int64_t nbytes)
{
    if (g_item_bytes) return g_item_bytes;
    const int64_t W = (int64_t)num_sms() * 3 * WARPS;
    if (nbytes >= W * 32768) return 32768;
    int64_t ib = ((nbytes + W - 1) / W + 255) & ~255ll;
    if (ib < 4096) ib = 4096;
    if (ib > 32768) ib = 32768;
    return (int)ib;
}
// the column's work-item index, built on first use; columns are shared read-only between threads, so the lazy build is
// serialised (the kernel is stream-ordered before every consumer on this thread's stream; a second thread waits for it)
static const int32_t* ensure_item_bounds(const custr_column* col, const int32_t* offsets, int first, int& nitems)
{
    const int ib = item_bytes_for(col->nbytes);
    const long long W = (long long)num_sms() * 3 * WARPS;
    const bool staggered = g_item_stagger && !g_item_bytes && col->nbytes >= W * 32768;  // at least one full round of 32 KiB items
    if (!staggered) nitems = (int)((col->nbytes + ib - 1) / ib);
    else nitems = (int)(W + (col->nbytes - item_start(W, W, ib, true) + ib - 1) / ib);
    const int key = staggered ? -ib : ib;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!col->item_bounds || col->item_bounds_count != nitems || col->item_bounds_bytes != key) {
        BufPtr b = dev_alloc(sizeof(int32_t) * (size_t)(nitems + 2));
        LAUNCH(k_item_bounds, (nitems + 1 + 255) / 256, 256, 0, offsets, col->n, first, nitems, ib, (int)W, staggered ? 1 : 0, (int32_t*)b->ptr);
        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
        col->item_bounds = b;
        col->item_bounds_count = nitems;
        col->item_bounds_bytes = key;
    }
    return (const int32_t*)col->item_bounds->ptr;
}

#include "regex_chain.cuh"
#include "regex_chain64.cuh"

// item-buffered chain kernel for boolean results (regex_chain_item.cuh), built in four translation units (regex_item.cu)
int chain_item_ctas_per_sm();
void launch_chain_item_g0(const ChainDev& cd, const Args& a, int blocks);
void launch_chain_item_g1(const ChainDev& cd, const Args& a, int blocks);
void launch_chain_item_g2(const ChainDev& cd, const Args& a, int blocks);
void launch_chain_item_g3(const ChainDev& cd, const Args& a, int blocks);
bool jit_launch_chain_item(const ChainDev& cd, const Args& a, int blocks);  // regex_jit.cu
std::atomic<long long> g_jit_launches{0};
thread_local bool g_no_jit = false;        // A/B switch (tier 5): ahead-of-time kernels only
long long g_jit_min_bytes = 64ll << 20;    // smallest column (bytes of chars) that is worth a run-time compilation
static void launch_chain_item(const ChainDev& cd, const Args& a, int blocks)
{
    if (cd.nsteps <= 2) launch_chain_item_g0(cd, a, blocks);
    else if (cd.nsteps <= 4) launch_chain_item_g1(cd, a, blocks);
    else if (cd.nsteps <= 6) launch_chain_item_g2(cd, a, blocks);
    else launch_chain_item_g3(cd, a, blocks);
}

const PlanDev& device_plan(const Plan& plan);

bool plan_is_chain(const Plan& plan) { return plan.is_chain; }

// count_re inside the chain kernel (see k_chain64's count mode): the last step loops and every step uses its class
bool count_in_kernel_ok(const Plan& plan)
{
    if (!plan.is_chain || !plan.span_ok) return false;
    const ChainDev& cd = plan.chain;
    if (cd.nsteps == 0 || !cd.steps[cd.nsteps - 1].loop) return false;
    for (uint32_t s = 0; s < cd.nsteps; ++s)
        if (cd.steps[s].cls != cd.steps[cd.nsteps - 1].cls) return false;
    return true;
}

bool run(const Plan& plan, const custr_column* col, const uint8_t* prog_img, const uint8_t* uflags, uint8_t* out,
         unsigned long long* total, int32_t** dirty_rows, unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count,
         SpanStreams* spans)
{
    if (spans && !(plan.is_chain && plan.span_ok && !g_force_generic)) return false;
    int32_t* counts = spans ? spans->counts_out : nullptr;
    if (counts && !count_in_kernel_ok(plan)) return false;

    const int32_t n = col->n;
    if (((uintptr_t)col->chars & 15) != 0) return false;  // vector loads need a 16-byte aligned base
    // boolean results of chain-shaped plans come from the item-buffered kernel, which writes every row itself
    const bool use_item = plan.is_chain && !g_force_generic && !g_chain_win && !spans && col->nbytes != 0;
    if (counts) CUSTR_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n, g_stream));
    else if (!use_item) CUSTR_CUDA(cudaMemsetAsync(out, 0, (size_t)n, g_stream));
    keep_rows = dev_alloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    // counters: [0] dirty-row count, [1] work-item counter.  A caller may hand in a zeroed 16-byte block whose first 8 bytes
    // are its match total (one allocation, one memset and one read-back per call instead of two of each)
    unsigned int* counters;
    if (keep_count && keep_count->bytes >= 16) counters = (unsigned int*)keep_count->ptr + 2;
    else {
        keep_count = dev_alloc(2 * sizeof(unsigned int));
        CUSTR_CUDA(cudaMemsetAsync(keep_count->ptr, 0, 2 * sizeof(unsigned int), g_stream));
        counters = (unsigned int*)keep_count->ptr;
    }
    *dirty_rows = (int32_t*)keep_rows->ptr;
    *dirty_count = counters;
    if (col->nbytes == 0) return true;
    Args a;
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    // (the DAG interpreter searches its own 32 KiB items; the chain kernels take theirs from the column's item index)
    a.nitems = (int)((col->nbytes + ITEM_BYTES - 1) / ITEM_BYTES);  // the chain kernels take theirs from ensure_item_bounds below
    a.out = out;
    a.total = total;
    a.dirty_rows = *dirty_rows;
    a.dirty_count = *dirty_count;
    a.item_counter = counters + 1;
    a.item_bounds = nullptr;
    a.prog_img = prog_img;
    a.uflags = uflags;
    a.span_m = a.span_k = a.span_a = nullptr;
    a.span_base = 0;
    a.counts = counts;
    if (spans && !counts) {
        a.span_base = a.first & ~(WIN64 - 1);
        const size_t words = (((size_t)(a.end - a.span_base) + WIN64 - 1) / WIN64) * (WIN64 / 64);
        spans->keep = dev_alloc(3 * words * sizeof(unsigned long long));
        CUSTR_CUDA(cudaMemsetAsync(spans->keep->ptr, 0, 3 * words * sizeof(unsigned long long), g_stream));
        a.span_m = (unsigned long long*)spans->keep->ptr;
        a.span_k = a.span_m + words;
        a.span_a = a.span_k + words;
        spans->m = a.span_m;
        spans->k = a.span_k;
        spans->a = a.span_a;
        spans->base = a.span_base;
        spans->words = (long long)words;
    }
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (plan.is_chain && !g_force_generic) {
        // 2048-byte windows, 64-bit streams, cp.async ring; grid = resident set (3 CTAs per SM), dynamic items
        a.item_bounds = ensure_item_bounds(col, a.offsets, a.first, a.nitems);  // once per column (it is immutable)
        blocks = (a.nitems + WARPS - 1) / WARPS;
        const int resident = num_sms() * (use_item ? chain_item_ctas_per_sm() : 3);
        if (blocks > resident) blocks = resident;
        if (use_item) {
            // the plan's own run-time compiled kernel (regex_jit.cu) when no ahead-of-time shape covers the plan and the
            // column is large enough to repay ~0.6 s of NVRTC on first use; else / on any failure the ahead-of-time kernels
            const ChainDev& cd = plan.chain;
            bool plain = true;
            for (uint32_t s = 0; s < cd.nsteps; ++s) plain = plain && !cd.steps[s].opt && ((cd.steps[s].exit != 0) == (s + 1 == cd.nsteps));
            const bool aot_shape = !g_no_spec && plain && cd.nsteps <= 4 && chain_spec_of(cd) != 0;
            const bool want_jit = !g_no_jit && (g_jit_mode == 2 || (g_jit_mode == 1 && !aot_shape && col->nbytes >= (int64_t)g_jit_min_bytes));
            if (want_jit && jit_launch_chain_item(cd, a, blocks)) g_jit_launches.fetch_add(1, std::memory_order_relaxed);
            else launch_chain_item(cd, a, blocks);
        } else
            launch_chain64(plan.chain, a, blocks);
        return true;
    }
#ifndef CUSTR_EXPERIMENT_ONLY_4_1
    LAUNCH(k_bitstream, blocks, THREADS, 0, device_plan(plan), a);
#endif
    return true;
}

#include "tokenize_bits.cuh"

// NVText::tokenize through the bit-stream compaction kernels.  delims == nullptr: whitespace.  False = not applicable
// (unaligned chars base, empty column): the caller uses the per-row path.
bool tokenize_flat(const custr_column* col, const uint8_t* delims, int ndelims, BufPtr& out_chars, BufPtr& out_off, int64_t& ntok, int64_t& nbytes)
{
    if (((uintptr_t)col->chars & 15) != 0 || col->nbytes == 0 || col->n == 0 || ndelims > TOK_DELIMS_MAX) return false;
    const int32_t n = col->n;
    TokArgs a{};
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    a.whitespace = delims ? 0u : 1u;
    a.ndelims = delims ? (uint32_t)ndelims : 0u;
    for (int k = 0; k < ndelims && delims; ++k) a.delims[k] = delims[k];
    a.item_bounds = ensure_item_bounds(col, a.offsets, a.first, a.nitems);
    const int win_base = a.first & ~(WIN64 - 1);
    const size_t nwin = ((size_t)(a.end - win_base) + WIN64 - 1) / WIN64;
    // every item adds at most one shared window; + one empty slot whose exclusive sum is the grand total
    const size_t nslots = nwin + (size_t)a.nitems + 1;
    Scratch<int32_t> item_w((size_t)a.nitems + 1), item_slot((size_t)a.nitems + 1);
    CUSTR_CUDA(cudaMemsetAsync(item_w.get() + a.nitems, 0, sizeof(int32_t), g_stream));
    LAUNCH(k_tok_item_windows, (a.nitems + 255) / 256, 256, 0, a.offsets, a.item_bounds, a.nitems, item_w.get());
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream);
        BufPtr t = dev_alloc(tb);
        CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(t->ptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    a.item_slot = item_slot.get();
    BufPtr counts = dev_alloc(sizeof(unsigned long long) * nslots), base = dev_alloc(sizeof(unsigned long long) * nslots);
    BufPtr counter = dev_alloc(2 * sizeof(unsigned int));
    CUSTR_CUDA(cudaMemsetAsync(counts->ptr, 0, sizeof(unsigned long long) * nslots, g_stream));
    CUSTR_CUDA(cudaMemsetAsync(counter->ptr, 0, 2 * sizeof(unsigned int), g_stream));
    a.slot_counts = (unsigned long long*)counts->ptr;
    a.slot_base = (const unsigned long long*)base->ptr;
    const int smem = WARPS * (int)sizeof(WarpSmTok);
    // (per device and cheap: set on every call rather than caching a flag that would be wrong after custr_set_device)
    CUSTR_CUDA(cudaFuncSetAttribute(k_tokenize64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUSTR_CUDA(cudaFuncSetAttribute(k_tokenize64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    const int resident = num_sms() * 3;
    if (blocks > resident) blocks = resident;
    a.item_counter = (unsigned int*)counter->ptr;
    auto kc = k_tokenize64<false>;
    LAUNCH(kc, blocks, THREADS, smem, a);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(tmp->ptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long total = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&total, (unsigned long long*)base->ptr + (nslots - 1), sizeof(total), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    ntok = (int64_t)(total >> 32);
    nbytes = (int64_t)(total & 0xffffffffull);
    out_chars = dev_alloc((size_t)nbytes);
    out_off = dev_alloc(sizeof(int32_t) * (size_t)(ntok + 1));
    const int32_t last = (int32_t)nbytes;
    CUSTR_CUDA(cudaMemcpyAsync((int32_t*)out_off->ptr + ntok, &last, sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    if (ntok) {
        a.tok_off = (int32_t*)out_off->ptr;
        a.out = (char*)out_chars->ptr;
        a.item_counter = (unsigned int*)counter->ptr + 1;
        auto kw = k_tokenize64<true>;
        LAUNCH(kw, blocks, THREADS, smem, a);
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));  // `last` and the scratch buffers die with this scope
    return true;
}

#include "split_bits.cuh"
#include "replace_bits.cuh"

// NVStrings::split_record(delimiter) for ONE ASCII delimiter byte and no split limit through the bit-stream kernels
// (split_bits.cuh): flat token column (chars + int32 offsets[ntok + 1]) and row_off[n + 1] (device).  False = not applicable
// (unaligned chars base, empty column, or the column holds an empty valid row): the caller uses the per-row path.
// The splice kernels pay per byte AND per row start (ROWSTART scatter, one output offset per row); the per-row kernels pay per
// row.  Below ~16 bytes per row (the README's day-of-week column: 3.3) the per-row walk wins — C3 chain 1.9 ms vs 5.7 ms.
constexpr int64_t STREAM_MIN_ROW_BYTES = 16;

bool split_record_flat(const custr_column* col, uint8_t delim, BufPtr& out_chars, BufPtr& out_off, BufPtr& row_off, int64_t& ntok, int64_t& nbytes)
{
    if (((uintptr_t)col->chars & 15) != 0 || col->nbytes == 0 || col->n == 0 || delim >= 0x80) return false;
    if (col->nbytes < STREAM_MIN_ROW_BYTES * (int64_t)col->n) return false;
    const int32_t n = col->n;
    SplitArgs a{};
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.validity = col->validity;
    a.vbit0 = col->vbit0;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    a.delim = delim;
    a.item_bounds = ensure_item_bounds(col, a.offsets, a.first, a.nitems);
    const int win_base = a.first & ~(WIN64 - 1);
    const size_t nwin = ((size_t)(a.end - win_base) + WIN64 - 1) / WIN64;
    const size_t nslots = nwin + (size_t)a.nitems + 1;
    Scratch<int32_t> item_w((size_t)a.nitems + 1), item_slot((size_t)a.nitems + 1);
    CUSTR_CUDA(cudaMemsetAsync(item_w.get() + a.nitems, 0, sizeof(int32_t), g_stream));
    LAUNCH(k_tok_item_windows, (a.nitems + 255) / 256, 256, 0, a.offsets, a.item_bounds, a.nitems, item_w.get());
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream);
        BufPtr t = dev_alloc(tb);
        CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(t->ptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    a.item_slot = item_slot.get();
    BufPtr counts = dev_alloc(sizeof(unsigned long long) * nslots), base = dev_alloc(sizeof(unsigned long long) * nslots);
    BufPtr counter = dev_alloc(4 * sizeof(unsigned int));  // [0], [1] work-item counters of the two passes, [2] flags
    CUSTR_CUDA(cudaMemsetAsync(counts->ptr, 0, sizeof(unsigned long long) * nslots, g_stream));
    CUSTR_CUDA(cudaMemsetAsync(counter->ptr, 0, 4 * sizeof(unsigned int), g_stream));
    a.slot_counts = (unsigned long long*)counts->ptr;
    a.slot_base = (const unsigned long long*)base->ptr;
    a.flags = (unsigned int*)counter->ptr + 2;
    const int smem = WARPS * (int)sizeof(WarpSmSplit);
    CUSTR_CUDA(cudaFuncSetAttribute(k_split_record64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUSTR_CUDA(cudaFuncSetAttribute(k_split_record64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    const int resident = num_sms() * 3;
    if (blocks > resident) blocks = resident;
    a.item_counter = (unsigned int*)counter->ptr;
    auto kc = k_split_record64<false>;
    LAUNCH(kc, blocks, THREADS, smem, a);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(tmp->ptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long total = 0;
    unsigned int flags = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&total, (unsigned long long*)base->ptr + (nslots - 1), sizeof(total), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaMemcpyAsync(&flags, a.flags, sizeof(flags), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    if (flags & 1u) return false;  // an empty valid row: several rows' first tokens on one byte
    ntok = (int64_t)(total >> 32);
    nbytes = (int64_t)(total & 0xffffffffull);
    out_chars = dev_alloc((size_t)nbytes);
    out_off = dev_alloc(sizeof(int32_t) * (size_t)(ntok + 1));
    row_off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
    const int32_t last_off = (int32_t)nbytes, last_row = (int32_t)ntok;
    CUSTR_CUDA(cudaMemcpyAsync((int32_t*)out_off->ptr + ntok, &last_off, sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    CUSTR_CUDA(cudaMemcpyAsync((int32_t*)row_off->ptr + n, &last_row, sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    a.tok_off = (int32_t*)out_off->ptr;
    a.row_off = (int32_t*)row_off->ptr;
    a.out = (char*)out_chars->ptr;
    a.item_counter = (unsigned int*)counter->ptr + 1;
    auto kw = k_split_record64<true>;
    LAUNCH(kw, blocks, THREADS, smem, a);
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return true;
}

// Both splice modes of replace_bits.cuh: count pass -> scan -> write pass.  `a` holds the mode's matcher arguments.
static bool splice_run(const custr_column* col, ReplArgs& a, int mode, BufPtr& out_chars, BufPtr& out_off, int64_t& nbytes)
{
    const int32_t n = col->n;
    a.chars = col->chars;
    a.offsets = col->offsets;
    a.n = n;
    a.first = col->first_off;
    a.end = col->first_off + (int32_t)col->nbytes;
    a.item_bounds = ensure_item_bounds(col, a.offsets, a.first, a.nitems);
    const size_t nslots = (size_t)col->nbytes / REPL_STRIDE + 2 * (size_t)a.nitems + 2;
    Scratch<int32_t> item_w((size_t)a.nitems + 1), item_slot((size_t)a.nitems + 1);
    CUSTR_CUDA(cudaMemsetAsync(item_w.get() + a.nitems, 0, sizeof(int32_t), g_stream));
    LAUNCH(k_repl_item_windows, (a.nitems + 255) / 256, 256, 0, a.offsets, a.item_bounds, a.nitems, mode == 1 ? ~63 : ~15, item_w.get());
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream);
        BufPtr t = dev_alloc(tb);
        CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(t->ptr, tb, item_w.get(), item_slot.get(), a.nitems + 1, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    a.item_slot = item_slot.get();
    BufPtr counts = dev_alloc(sizeof(unsigned long long) * nslots), base = dev_alloc(sizeof(unsigned long long) * nslots);
    BufPtr counter = dev_alloc(4 * sizeof(unsigned int));  // [0], [1] work-item counters of the two passes, [2] flags
    CUSTR_CUDA(cudaMemsetAsync(counts->ptr, 0, sizeof(unsigned long long) * nslots, g_stream));
    CUSTR_CUDA(cudaMemsetAsync(counter->ptr, 0, 4 * sizeof(unsigned int), g_stream));
    a.slot_counts = (unsigned long long*)counts->ptr;
    a.slot_base = (const unsigned long long*)base->ptr;
    a.flags = (unsigned int*)counter->ptr + 2;
    const int smem = WARPS * (int)sizeof(WarpSmRepl);
    auto kc = mode == 1 ? k_replace_splice64<1, false> : k_replace_splice64<0, false>;
    auto kw = mode == 1 ? k_replace_splice64<1, true> : k_replace_splice64<0, true>;
    CUSTR_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUSTR_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int blocks = (a.nitems + WARPS - 1) / WARPS;
    const int resident = num_sms() * 3;
    if (blocks > resident) blocks = resident;
    a.item_counter = (unsigned int*)counter->ptr;
    trace_point("splice: set-up");
    LAUNCH(kc, blocks, THREADS, smem, a);
    trace_point("splice: count pass");
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(tmp->ptr, tmp_bytes, a.slot_counts, (unsigned long long*)base->ptr, (int)nslots, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long total = 0;
    unsigned int flags = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&total, (unsigned long long*)base->ptr + (nslots - 1), sizeof(total), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaMemcpyAsync(&flags, a.flags, sizeof(flags), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    if (flags & 1u) return false;
    if (total > 0x7fffffffull) throw ArgError{fail(CUSTR_ERR_INVALID, "replace: result exceeds 2 GiB of chars")};
    nbytes = (int64_t)total;
    out_chars = dev_alloc((size_t)nbytes);
    out_off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
    const int32_t last_off = (int32_t)nbytes;
    CUSTR_CUDA(cudaMemcpyAsync((int32_t*)out_off->ptr + n, &last_off, sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    a.new_off = (int32_t*)out_off->ptr;
    a.out = (char*)out_chars->ptr;
    a.item_counter = (unsigned int*)counter->ptr + 1;
    trace_point("splice: scan + allocs");
    LAUNCH(kw, blocks, THREADS, smem, a);
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    trace_point("splice: write pass");
    return true;
}
// every occurrence adds rlen bytes and removes at least `shortest`: the most a 1984-byte window can emit must fit the tile
static bool splice_fits_tile(int shortest, int rlen)
{
    return rlen <= shortest || REPL_STRIDE + ((REPL_STRIDE + shortest - 1) / shortest) * (rlen - shortest) + shortest <= REPL_TILE;
}

// NVStrings::replace, literal target, every occurrence (replace_bits.cuh MODE 0).  false = not expressible there (bordered or
// long target, replacement that could overflow the staging tile, unaligned view, very short rows): the caller takes the per-row path.
bool replace_literal_flat(const custr_column* col, const char* pat, int m, const char* repl, int rlen, BufPtr& out_chars, BufPtr& out_off,
                          int64_t& nbytes)
{
    if (((uintptr_t)col->chars & 15) != 0 || col->nbytes == 0 || col->n == 0) return false;
    if (m < 1 || m > REPL_PAT_MAX || rlen > REPL_REPL_MAX) return false;
    if (col->nbytes < STREAM_MIN_ROW_BYTES * (int64_t)col->n) return false;
    for (int b = 1; b < m; ++b)  // a border: occurrences could overlap and the leftmost scan would skip some
        if (memcmp(pat, pat + (m - b), (size_t)b) == 0) return false;
    if (!splice_fits_tile(m, rlen)) return false;
    ReplArgs a{};
    a.m = m;
    a.rlen = rlen;
    memcpy(a.pat, pat, (size_t)m);
    memcpy(a.repl, repl, (size_t)rlen);
    return splice_run(col, a, 0, out_chars, out_off, nbytes);
}

// replace_re, every match, for single-class chains (replace_bits.cuh MODE 1) from the span streams a run() over the same column
// left in `ss`.  false = not this family / not expressible: the caller walks the streams per row (span_walk.cuh).
bool replace_spans_ok(const ChainDev& cd)
{
    if (cd.nsteps < 1 || cd.nsteps > CHAIN_MAX_STEPS) return false;
    for (uint32_t s = 0; s < cd.nsteps; ++s) {
        const ChainStepD& st = cd.steps[s];
        if (st.cls != cd.steps[cd.nsteps - 1].cls || st.opt || (s > 0 && st.before)) return false;
        if ((st.loop != 0) != (s + 1 == cd.nsteps) || (st.exit != 0) != (s + 1 == cd.nsteps)) return false;
    }
    return true;
}
bool replace_spans_flat(const custr_column* col, const ChainDev& cd, const SpanStreams& ss, const char* repl, int rlen, BufPtr& out_chars,
                        BufPtr& out_off, int64_t& nbytes)
{
    if (!replace_spans_ok(cd) || !ss.m || ss.words <= 0) return false;
    if (((uintptr_t)col->chars & 15) != 0 || col->nbytes == 0 || col->n == 0 || rlen > REPL_REPL_MAX) return false;
    if (col->nbytes < STREAM_MIN_ROW_BYTES * (int64_t)col->n) return false;
    if (!splice_fits_tile((int)cd.nsteps, rlen)) return false;
    ReplArgs a{};
    a.m = (int)cd.nsteps;
    a.rlen = rlen;
    memcpy(a.repl, repl, (size_t)rlen);
    a.span_m = ss.m;
    a.span_k = ss.k;
    a.span_a = ss.a;
    a.span_words = ss.words;
    a.span_base = ss.base;
    a.k_chars = (int)cd.nsteps - 1;
    return splice_run(col, a, 1, out_chars, out_off, nbytes);
}

}  // namespace bits
}  // namespace custr
