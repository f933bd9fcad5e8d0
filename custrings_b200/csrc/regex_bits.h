// Bit-parallel ("bitstream") regex tier: model in regex_bits_plan.h, lowering in regex_bits_lower.cpp, device
// executor in regex_bits.cu.
#pragma once
#include "common.cuh"
#include "regex_prog.h"
#include <memory>
#include <string>

namespace custr {
namespace bits {

struct Plan;  // lowered program

// Try to lower a compiled pattern to a bitstream plan for a boolean search (anchored = `match`, else
// `contains_re`).  Returns null when the pattern is outside the provably-equivalent subset.
std::shared_ptr<Plan> lower(const rx::Program& prog, bool anchored, const uint8_t* unicode_flags);
std::string describe(const Plan& plan);

// Device run.  out[i] = 1 for matching rows that are pure ASCII; *total += their number.  Rows with a non-ASCII or NUL
// byte are NOT decided: their indices are appended to *dirty_rows (count in *dirty_count, both device memory owned by the
// returned buffers) for the exact Pike-VM kernel.  Returns false (nothing launched) if the column layout is not
// supported (unaligned chars base).  Enqueued on g_stream.
//
// With `spans` (only for plans where span_chain() is non-null) the chain kernel also leaves three bit streams of the last
// step in device memory — M (a match can begin its last step here), K (the match may continue into this byte), A (a match
// may end after this byte) — from which span_walk.cuh derives the exact match spans of every row.
struct SpanStreams {
    const unsigned long long* m = nullptr;
    const unsigned long long* k = nullptr;
    const unsigned long long* a = nullptr;
    int32_t base = 0;   // byte offset (into chars) of bit 0
    long long words = 0;  // 64-bit words per stream
    BufPtr keep;
    // count mode: when set (and count_in_kernel_ok(plan)) the kernel writes the number of matches of every row here and
    // leaves no streams; `out` of run() is then unused
    int32_t* counts_out = nullptr;
};
bool count_in_kernel_ok(const Plan& plan);
bool plan_is_chain(const Plan& plan);  // linear chain (k_chain64) rather than a DAG (k_bitstream interpreter)
bool run(const Plan& plan, const custr_column* col, const uint8_t* prog_img, const uint8_t* uflags, uint8_t* out,
         unsigned long long* total, int32_t** dirty_rows, unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count,
         SpanStreams* spans = nullptr);
// NVText::tokenize as a bit-stream compaction (tokenize_bits.cuh): whitespace (delims == nullptr) or up to 8 ASCII
// delimiter bytes.  Produces the flat token column (chars + int32 offsets[ntok + 1]).  False = not applicable.
bool tokenize_flat(const custr_column* col, const uint8_t* delims, int ndelims, BufPtr& out_chars, BufPtr& out_off, int64_t& ntok,
                   int64_t& nbytes);
// NVStrings::split_record with one ASCII delimiter byte and no split limit as a bit-stream compaction (split_bits.cuh): flat
// token column + row_off[n + 1] on the device.  False = not applicable (e.g. the column holds an empty valid row).
bool split_record_flat(const custr_column* col, uint8_t delim, BufPtr& out_chars, BufPtr& out_off, BufPtr& row_off, int64_t& ntok, int64_t& nbytes);
// replace with a literal target, every occurrence (replace_bits.cuh); false = not expressible, take the per-row path
bool replace_literal_flat(const custr_column* col, const char* pat, int m, const char* repl, int rlen, BufPtr& out_chars, BufPtr& out_off,
                          int64_t& nbytes);
struct ChainDev;
// replace_re, every match, for single-class chains (x{n,} / x+ with assertions) from the span streams (replace_bits.cuh MODE 1)
bool replace_spans_ok(const ChainDev& cd);
bool replace_spans_flat(const custr_column* col, const ChainDev& cd, const SpanStreams& ss, const char* repl, int rlen, BufPtr& out_chars,
                        BufPtr& out_off, int64_t& nbytes);
extern thread_local bool g_force_generic;
extern thread_local bool g_no_spec;
extern thread_local bool g_chain_win;
extern int g_item_bytes;
extern bool g_item_stagger;
extern thread_local bool g_no_jit;
extern int g_jit_mode;               // 0 never, 1 large columns without an ahead-of-time shape (default), 2 always
extern long long g_jit_min_bytes;
extern std::atomic<long long> g_jit_launches;
const char* jit_last_note();

// Non-null when the plan is a linear chain whose only loop is its last step: for those patterns the match SPAN the Pike VM
// reports (leftmost start, then thread priority) is "leftmost start, longest end that satisfies the trailing assertion",
// which chain_spans.cuh computes with a scalar per-row scan (count_re / replace_re / findall fast path).
struct ChainDev;
const ChainDev* span_chain(const Plan& plan);

// Plain host executor of the chain model of a plan (tests/sim only); false when the plan is not a chain
bool reference_execute_chain(const Plan& plan, const char* chars, const int32_t* offsets, int32_t n, uint8_t* out, uint8_t* dirty);
// Plain host executor of a plan (tests/sim only): out[i] for every row, dirty[i] = row needs the exact path.
void reference_execute(const Plan& plan, const char* chars, const int32_t* offsets, const uint8_t* validity, int32_t n,
                       uint8_t* out, uint8_t* dirty);

}  // namespace bits
}  // namespace custr
