#include "regex_bits.h"
namespace custr { namespace bits {
struct Plan { int dummy; };
std::shared_ptr<Plan> lower(const rx::Program&, bool, const uint8_t*) { return nullptr; }
std::string describe(const Plan&) { return "stub"; }
void run(const Plan&, const custr_column*, uint8_t*, unsigned long long*) {}
}}
