// Host-side simulation harness: runs the SAME __host__ __device__ per-row code the CUDA kernels run
// (custrings_b200/csrc/regex_vm.cuh, rowops.cuh) on the CPU, one row at a time, so that the engine logic can be
// fuzzed against the oracle in the GPU-less container (`-m "not gpu"` tests).  Test infrastructure only — it is
// not a CPU fallback: nothing in custrings_b200/ links or loads it.
#include "../../custrings_b200/csrc/regex_vm.cuh"
#include "../../custrings_b200/csrc/rowops.cuh"
#include "../../custrings_b200/csrc/regex_bits.h"
#include "../../custrings_b200/csrc/chain_spans.cuh"
#include <vector>
#include <string>

using namespace custr;

static const uint8_t k_flags[65536] = {
#include "../../custrings_b200/csrc/unicode_flags.inc"
};

namespace {
struct Prog {
    rx::Program prog;
    std::vector<uint8_t> img;
    rxdev::DevProg P;
    explicit Prog(const char* pattern)
    {
        prog = rx::compile(pattern);
        img = rx::serialize(prog, k_flags);
        P = rxdev::bind_program(img.data(), k_flags);
    }
};
using L = rxdev::Lists<1024>;
}  // namespace

extern "C" {

int sim_describe(const char* pattern, char* buf, int buflen)
{
    Prog p(pattern);
    std::string d = p.prog.describe();
    std::shared_ptr<bits::Plan> plan = bits::lower(p.prog, false, k_flags);
    d += plan ? "bitstream: " + bits::describe(*plan) + "\n" : "bitstream: not eligible\n";
    snprintf(buf, buflen, "%s", d.c_str());
    return (int)p.prog.insts.size();
}

int sim_bool(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, int anchored, uint8_t* out)
{
    Prog p(pattern);
    if (p.prog.insts.size() > 1024) return -2;
    ColView col{chars, off, validity, 0, n};
    L* lists = new L;
    lists->init();
    int total = 0;
    for (int i = 0; i < n; ++i) {
        int hit = 0;
        if (col.valid(i)) {
            int len = off[i + 1] - off[i], mb, me;
            hit = rxdev::vm_find<1024>(p.P, (const uint8_t*)chars + off[i], len, 0, anchored ? 1 : len, mb, me, *lists);
        }
        out[i] = (uint8_t)hit;
        total += hit;
    }
    delete lists;
    return total;
}

int sim_count(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, int32_t* out)
{
    Prog p(pattern);
    if (p.prog.insts.size() > 1024) return -2;
    ColView col{chars, off, validity, 0, n};
    L* lists = new L;
    lists->init();
    int total = 0;
    for (int i = 0; i < n; ++i) {
        int c = 0;
        if (col.valid(i)) c = rxdev::row_count<1024>(p.P, (const uint8_t*)chars + off[i], off[i + 1] - off[i], *lists);
        out[i] = c;
        total += c != 0;
    }
    delete lists;
    return total;
}

// bitstream tier on the host: plan reference executor for ASCII rows + exact VM for the rows it flags dirty.
// returns -1 when the pattern is not eligible for the bitstream tier.
int sim_bits_bool(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, int anchored, uint8_t* out)
{
    Prog p(pattern);
    std::shared_ptr<bits::Plan> plan = bits::lower(p.prog, anchored != 0, k_flags);
    if (!plan) return -1;
    std::vector<uint8_t> dirty(n ? n : 1);
    bits::reference_execute(*plan, chars, off, validity, n, out, dirty.data());
    ColView col{chars, off, validity, 0, n};
    L* lists = new L;
    lists->init();
    int total = 0;
    for (int i = 0; i < n; ++i) {
        if (dirty[i]) {
            int len = off[i + 1] - off[i], mb, me;
            out[i] = col.valid(i) ? (uint8_t)rxdev::vm_find<1024>(p.P, (const uint8_t*)chars + off[i], len, 0, anchored ? 1 : len, mb, me, *lists) : 0;
        }
        total += out[i];
    }
    delete lists;
    return total;
}

// the CHAIN model (ChainDev with loop / opt / exit flags) executed on the host; -1 when the pattern is not a chain
int sim_chain_bool(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, int anchored, uint8_t* out)
{
    Prog p(pattern);
    std::shared_ptr<bits::Plan> plan = bits::lower(p.prog, anchored != 0, k_flags);
    if (!plan) return -1;
    std::vector<uint8_t> dirty(n ? n : 1);
    if (!bits::reference_execute_chain(*plan, chars, off, n, out, dirty.data())) return -1;
    ColView col{chars, off, validity, 0, n};
    L* lists = new L;
    lists->init();
    int total = 0;
    for (int i = 0; i < n; ++i) {
        if (dirty[i]) {
            int len = off[i + 1] - off[i], mb, me;
            out[i] = col.valid(i) ? (uint8_t)rxdev::vm_find<1024>(p.P, (const uint8_t*)chars + off[i], len, 0, anchored ? 1 : len, mb, me, *lists) : 0;
        }
        total += out[i];
    }
    delete lists;
    return total;
}

// span fast path (chain_spans.cuh) on the host: count_re.  returns -1 when the pattern is not a last-loop chain
int sim_chain_count(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, int32_t* out)
{
    Prog p(pattern);
    std::shared_ptr<bits::Plan> plan = bits::lower(p.prog, false, k_flags);
    const bits::ChainDev* cd = plan ? bits::span_chain(*plan) : nullptr;
    if (!cd) return -1;
    ColView col{chars, off, validity, 0, n};
    for (int i = 0; i < n; ++i)
        if (col.valid(i) && spans::has_nul((const uint8_t*)chars + off[i], off[i + 1] - off[i])) return -1;  // product falls back to the VM
    int total = 0;
    for (int i = 0; i < n; ++i) {
        out[i] = col.valid(i) ? spans::row_count(*cd, (const uint8_t*)chars + off[i], off[i + 1] - off[i], k_flags) : 0;
        total += out[i] != 0;
    }
    return total;
}

long sim_chain_replace(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, const char* repl,
                       int maxrepl, int32_t* out_off, char* out_chars)
{
    Prog p(pattern);
    std::shared_ptr<bits::Plan> plan = bits::lower(p.prog, false, k_flags);
    const bits::ChainDev* cd = plan ? bits::span_chain(*plan) : nullptr;
    if (!cd) return -1;
    ColView col{chars, off, validity, 0, n};
    for (int i = 0; i < n; ++i)
        if (col.valid(i) && spans::has_nul((const uint8_t*)chars + off[i], off[i + 1] - off[i])) return -1;
    int rl = (int)strlen(repl);
    long run = 0;
    for (int i = 0; i < n; ++i) {
        out_off[i] = (int32_t)run;
        if (col.valid(i))
            run += spans::row_replace(*cd, (const uint8_t*)chars + off[i], off[i + 1] - off[i], k_flags, repl, rl, maxrepl,
                                      out_chars ? out_chars + run : nullptr);
    }
    out_off[n] = (int32_t)run;
    return run;
}

// out_off[n+1] always written; out_chars written when non-null (second call)
long sim_replace_re(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* pattern, const char* repl,
                    int maxrepl, int32_t* out_off, char* out_chars)
{
    Prog p(pattern);
    if (p.prog.insts.size() > 1024) return -2;
    ColView col{chars, off, validity, 0, n};
    L* lists = new L;
    lists->init();
    int rl = (int)strlen(repl);
    long run = 0;
    for (int i = 0; i < n; ++i) {
        out_off[i] = (int32_t)run;
        if (col.valid(i))
            run += rxdev::row_replace<1024>(p.P, (const uint8_t*)chars + off[i], off[i + 1] - off[i], repl, rl, maxrepl,
                                            out_chars ? out_chars + run : nullptr, *lists);
    }
    out_off[n] = (int32_t)run;
    delete lists;
    return run;
}

long sim_replace_re_multi(const char* chars, const int32_t* off, const uint8_t* validity, int n, const char* const* patterns,
                          int npat, const char* rchars, const int32_t* roff, const uint8_t* rvalid, int nrepl, int32_t* out_off,
                          char* out_chars)
{
    std::vector<Prog*> progs;
    std::vector<const uint8_t*> imgs;
    for (int t = 0; t < npat; ++t) { progs.push_back(new Prog(patterns[t])); imgs.push_back(progs.back()->img.data()); }
    ColView col{chars, off, validity, 0, n};
    ColView repls{rchars, roff, rvalid, 0, nrepl};
    L* lists = new L;
    lists->init();
    long run = 0;
    for (int i = 0; i < n; ++i) {
        out_off[i] = (int32_t)run;
        if (col.valid(i))
            run += rxdev::row_replace_multi<1024>(imgs.data(), npat, k_flags, repls, (const uint8_t*)chars + off[i],
                                                  off[i + 1] - off[i], out_chars ? out_chars + run : nullptr, *lists);
    }
    out_off[n] = (int32_t)run;
    delete lists;
    for (Prog* q : progs) delete q;
    return run;
}

// this repo's compiled program in the reference's vocabulary (instruction type codes of regcomp.h:25-40), same word
// layout as oracle's ref_regex_dump so the two compilers can be compared graph against graph
int sim_program_dump(const char* pattern, int* out, int cap)
{
    rx::Program prog = rx::compile(pattern);
    auto type_of = [](int op) {
        switch (op) {
            case rx::OP_CHAR: return 0177;  case rx::OP_ANY: return 0300;   case rx::OP_ANYNL: return 0301;
            case rx::OP_CLASS: return 0305; case rx::OP_NCLASS: return 0306; case rx::OP_END: return 0377;
            case rx::OP_LBRA: return 0202;  case rx::OP_RBRA: return 0201;  case rx::OP_BOL: return 0303;
            case rx::OP_EOL: return 0304;   case rx::OP_BOW: return 0307;   case rx::OP_NBOW: return 0310;
            case rx::OP_SPLIT: return 0204; default: return -op;
        }
    };
    std::vector<int> w;
    w.push_back((int)prog.insts.size());
    w.push_back(prog.start_inst);
    w.push_back(prog.ngroups);
    w.push_back((int)prog.starts.size());
    w.push_back((int)prog.classes.size());
    for (const rx::Inst& in : prog.insts) {
        w.push_back(type_of(in.op));
        w.push_back(in.op == rx::OP_SPLIT ? in.other : (int)in.arg);
        w.push_back(in.next);
    }
    for (int s : prog.starts) w.push_back(s);
    for (const rx::Class& c : prog.classes) {
        w.push_back(c.builtins);
        w.push_back((int)c.ranges.size());
        for (uint32_t r : c.ranges) w.push_back((int)r);
    }
    for (size_t i = 0; i < w.size() && (int)i < cap; ++i) out[i] = w[i];
    return (int)w.size();
}

}  // extern "C"
