// Bit-parallel ("bitstream") regex tier: see regex_bits.cu.
#pragma once
#include "common.cuh"
#include "regex_prog.h"
#include <memory>
#include <string>

namespace custr {
namespace bits {

struct Plan;  // lowered program (host description + device image)

// Try to lower a compiled pattern to a bitstream plan for a boolean search (anchored = `match`, else
// `contains_re`).  Returns null when the pattern is outside the provably-equivalent subset.
std::shared_ptr<Plan> lower(const rx::Program& prog, bool anchored, const uint8_t* unicode_flags);
std::string describe(const Plan& plan);
// out[i] = 1 if row i matches; *total += number of matching rows.  Enqueued on g_stream.
void run(const Plan& plan, const custr_column* col, uint8_t* out, unsigned long long* total);

}  // namespace bits
}  // namespace custr
