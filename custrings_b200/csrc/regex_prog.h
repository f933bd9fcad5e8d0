// Compiled-regex program shared by the host compiler (regex_compile.cpp) and the device engines.
// It is this repo's own encoding of the reference's Reprog (cpp/src/regex/regcomp.h:51-103): same
// instruction graph (so thread priority, de-duplication and therefore match spans are identical), different
// layout: a SPLIT keeps the reference's union trick (the continuation is always `next`), classes carry a
// precomputed 128-bit ASCII membership bitmap next to the exact (ranges + builtin flags) definition.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace custr {
namespace rx {

enum Op : int32_t {
    OP_CHAR = 1,   // arg = UTF-8 packed char (bytes big-endian in a u32, reference `Char`)
    OP_ANY,        // any char except '\n'
    OP_ANYNL,
    OP_CLASS,      // arg = class id
    OP_NCLASS,
    OP_END,
    OP_LBRA,       // arg = capture group id (>=1)
    OP_RBRA,
    OP_BOL,        // arg = '^' (also after '\n') or 'A' (\A: only at 0)
    OP_EOL,        // arg = '$' (also before '\n') or 'Z'
    OP_BOW,
    OP_NBOW,
    OP_SPLIT,      // try `other` first (higher priority), then `next`
    OP_NOP,        // compile time only
    OP_BAD         // malformed construct in the pattern: never matches, kills the thread
};

struct Inst {
    int32_t op;
    uint32_t arg;
    int32_t next;
    int32_t other;
};

// builtin flags inside a class (reference regcomp.cpp:53-58)
enum : int32_t { CB_W = 1, CB_S = 2, CB_D = 4, CB_NW = 8, CB_NS = 16, CB_ND = 32 };

struct Class {
    int32_t builtins = 0;
    std::vector<uint32_t> ranges;  // pairs lo,hi of packed chars, sorted & merged
};

struct Program {
    std::vector<Inst> insts;
    std::vector<int32_t> starts;  // seed order = priority order; no terminator
    std::vector<Class> classes;
    int32_t start_inst = 0;
    int32_t ngroups = 0;
    bool malformed = false;
    std::string describe() const;
};

// UTF-8 pattern -> program (never throws; malformed constructs become OP_BAD/no-ops like the reference,
// which "silently inserts NOPs", regcomp.cpp:616-621)
Program compile(const char* pattern_utf8);

// ---- flat device image ------------------------------------------------------------------------------
struct DevClass {
    uint32_t ascii[4];   // bit c set <=> class matches (positive sense, before NCLASS negation) ASCII char c
    int32_t builtins;
    int32_t range_begin; // index into the ranges array (in u32 units, pairs)
    int32_t range_count; // number of u32 entries (2 per range)
    int32_t pad;
};
struct DevHeader {
    int32_t ninsts, nstarts, nclasses, nranges;
    int32_t start_inst, ngroups;
    int32_t start_op;     // OP_CHAR / OP_BOL when the start-skip of regexec.inl:217-246 applies, else 0
    uint32_t start_arg;
};
// image = DevHeader | Inst[ninsts] | int32 starts[nstarts] | DevClass[nclasses] | uint32 ranges[nranges]
std::vector<uint8_t> serialize(const Program& p, const uint8_t* unicode_flags);

// exact class membership on the host (used to build the ASCII bitmaps and by the bitstream lowering)
bool class_matches(const Class& c, uint32_t packed_char, const uint8_t* unicode_flags);
uint32_t packed_to_codepoint(uint32_t packed);

}  // namespace rx
}  // namespace custr
