// Bitstream ("Parabix-style") regex plan: the data handed from the host lowering (regex_bits_lower.cpp) to the
// device executor (regex_bits.cu).  See DESIGN.md §4.2.
//
// Model.  Every byte position of the flat chars buffer owns one bit in each stream.  For a compiled program whose
// consuming instructions form a DAG (self-loops on one instruction allowed) we keep, per consuming instruction j,
// the marker stream  P_j[p] = "some NFA thread has just consumed byte p through instruction j".  Then
//     P_j = CLASS_j & ( OR_{i -> j} before(mask_ij, advance(P_i) & ~ROWSTART)  |  before(mask_sj, START) )
//     P_j = spread(P_j, CLASS_j & ~ROWSTART & before(self_mask))              when j loops on itself
//     MATCH = OR_{i -> END} after(mask_i, P_i)
// where `before(mask, .)` applies zero-width assertions evaluated between p-1 and p, `after(mask, .)` the same
// between p and p+1 (p+1 possibly being the end of the row), and ROWSTART marks the first byte of every row so that
// nothing flows between rows.  Match EXISTENCE does not depend on thread priority, so this equals the reference
// Pike VM's answer for contains_re / match whenever the lowering accepts the pattern.  Rows containing a byte
// >= 0x80 or == 0 are not decided here: they are queued for the exact Pike-VM kernel.
#pragma once
#include <cstdint>

namespace custr {
namespace bits {

enum AtomKind : uint8_t { AK_EQ = 0, AK_RANGE, AK_WORD, AK_ALNUM, AK_DIGIT, AK_SPACE, AK_LOWER, AK_UPPER, AK_ANY };
// zero-width assertion mask bits
enum : uint8_t { AS_BOW = 1, AS_NBOW = 2, AS_BOL_CARET = 4, AS_BOL_A = 8, AS_EOL_DOLLAR = 16, AS_EOL_Z = 32 };

constexpr int MAX_CLASSES = 8;
constexpr int MAX_ATOMS = 6;
constexpr int MAX_STEPS = 24;
constexpr int MAX_PREDS = 6;
constexpr int MAX_ENDS = 8;
constexpr int SRC_START = 255;

// how a class treats a NON-ASCII character (decided by decoding it, only where such bytes occur)
enum NaKind : uint32_t { NA_NEVER = 0, NA_ALWAYS, NA_CHAR_EQ, NA_CLASS, NA_NCLASS };
struct AtomD { uint8_t kind, lo, hi, pad; };
struct ClassD { uint8_t natoms, negate, pad0, pad1; AtomD atoms[MAX_ATOMS]; uint32_t na_kind, na_arg; };
struct PredD { uint8_t src, mask; };                       // src = step index or SRC_START
struct StepD { uint8_t cls, npreds, self_loop, self_mask; PredD preds[MAX_PREDS]; };
struct EndD { uint8_t src, mask; };

struct PlanDev {
    uint8_t nclasses, nsteps, nends, anchored;
    uint8_t before_needs, after_needs, pad0, pad1;  // union of assertion bits used in each context
    ClassD classes[MAX_CLASSES];
    StepD steps[MAX_STEPS];
    EndD ends[MAX_ENDS];
};

// Linear-chain specialisation (regex_bits.cu k_chain): step s is fed only by step s-1 (step 0 by START), END hangs
// off the last step.  Covers literals, class sequences, x+, x*, x{n,m}-free tails, leading/trailing assertions.
constexpr int CHAIN_MAX_STEPS = 8;
constexpr int CHAIN_MAX_CLASSES = 8;
struct ChainStepD { uint32_t cls, before, loop, opt, exit; };  // opt: the step may be skipped (x?, x*); exit: edge into END
struct ChainClassD {
    uint32_t builtins;   // OR of (1 << AtomKind) for the builtin atoms (AK_WORD .. AK_ANY)
    uint32_t natoms;     // remaining EQ / RANGE atoms
    AtomD atoms[MAX_ATOMS];
    uint32_t negate, na_kind, na_arg;
    // exact definition of the source class for non-ASCII characters, kept in the kernel parameter block so the rare
    // decode path does not chase the program image through global memory (na_inline == 0: use the image)
    uint32_t na_inline, na_builtins, na_nranges;
    uint32_t na_ranges[8];
    uint32_t na2[64];    // membership bitmap of the 2-byte characters U+0080..U+07FF (bit = code point), exact
    uint32_t ascii[4];   // exact ASCII membership bitmap (used by the per-row span matcher, chain_spans.cuh)
};
struct ChainDev {
    uint32_t nsteps, nclasses, anchored, end_mask;
    uint32_t needs;           // union of all assertion bits
    uint32_t builtin_union;   // union of ChainClassD::builtins (which builtin streams to compute per window)
    uint32_t na2_alnum[64];   // \\b-alphanumeric bitmap of U+0080..U+07FF
    ChainStepD steps[CHAIN_MAX_STEPS];
    ChainClassD classes[CHAIN_MAX_CLASSES];
};

// ASCII membership of one byte in an atom / class (shared by host reference executor and lowering checks)
inline bool atom_has(const AtomD& a, unsigned c)
{
    switch (a.kind) {
    case AK_EQ: return c == a.lo;
    case AK_RANGE: return c >= a.lo && c <= a.hi;
    case AK_WORD: return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
    case AK_ALNUM: return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
    case AK_DIGIT: return c >= '0' && c <= '9';
    case AK_SPACE: return (c >= 9 && c <= 13) || (c >= 28 && c <= 32);
    case AK_LOWER: return c >= 'a' && c <= 'z';
    case AK_UPPER: return c >= 'A' && c <= 'Z';
    case AK_ANY: return true;
    }
    return false;
}
inline bool class_has(const ClassD& k, unsigned c)
{
    bool in = false;
    for (int i = 0; i < k.natoms; ++i) in = in || atom_has(k.atoms[i], c);
    return in != (k.negate != 0);
}

}  // namespace bits
}  // namespace custr
