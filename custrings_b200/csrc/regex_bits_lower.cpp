// Host lowering: rx::Program (NFA graph) -> bits::PlanDev (marker-stream plan), plus a plain host executor of the
// plan used by tests/sim to validate the lowering against the oracle without a GPU.  See regex_bits_plan.h for the
// model and the equivalence argument.  A pattern is accepted only when
//   - every consuming instruction tests an ASCII-decidable class (rows with non-ASCII bytes are re-run exactly),
//   - the consuming instructions form a DAG apart from self-loops (x*, x+, x{n,}),
//   - the pattern cannot match the empty string (the reference's rules for empty matches depend on seeding
//     position, regexec.inl:260-267 — those patterns stay on the exact VM),
//   - it fits the fixed plan limits.
#include "regex_bits.h"
#include "regex_bits_plan.h"
#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

namespace custr {
namespace bits {

struct Plan {
    PlanDev dev;
    bool is_chain = false;
    bool span_ok = false;  // chain whose only loop is a GREEDY loop on the last step (see span_chain)
    bool chain_has_opt = false;  // some step is optional: only the 64-bit chain kernel understands that
    ChainDev chain;
    std::string text;
};

namespace {

struct Bitmap128 {
    uint32_t w[4] = {0, 0, 0, 0};
    bool get(unsigned c) const { return (w[c >> 5] >> (c & 31)) & 1u; }
    void set(unsigned c) { w[c >> 5] |= 1u << (c & 31); }
    bool operator<(const Bitmap128& o) const { return memcmp(w, o.w, sizeof(w)) < 0; }
};

bool decompose_set(const Bitmap128& want, ClassD& out)  // positive set (bit 0 is "don't care")
{
    bool rest[128];
    for (unsigned c = 0; c < 128; ++c) rest[c] = c ? want.get(c) : false;
    int n = 0;
    auto covered = [&](uint8_t kind) {
        AtomD a{kind, 0, 0, 0};
        bool any = false;
        for (unsigned c = 1; c < 128; ++c)
            if (atom_has(a, c)) { if (!want.get(c)) return false; any = any || rest[c]; }
        return any;
    };
    auto take = [&](uint8_t kind) {
        AtomD a{kind, 0, 0, 0};
        for (unsigned c = 1; c < 128; ++c) if (atom_has(a, c)) rest[c] = false;
        if (n < MAX_ATOMS) out.atoms[n] = a;
        ++n;
    };
    for (uint8_t kind : {AK_ANY, AK_WORD, AK_ALNUM, AK_LOWER, AK_UPPER, AK_DIGIT, AK_SPACE})
        if (covered(kind)) take(kind);
    for (unsigned c = 1; c < 128;) {
        if (!rest[c]) { ++c; continue; }
        unsigned e = c;
        while (e + 1 < 128 && rest[e + 1]) ++e;
        // a run may be extended over bytes that are already covered / wanted: keeps atoms few
        if (n < MAX_ATOMS) out.atoms[n] = AtomD{(uint8_t)(c == e ? AK_EQ : AK_RANGE), (uint8_t)c, (uint8_t)e, 0};
        ++n;
        c = e + 1;
    }
    if (n > MAX_ATOMS) return false;
    out.natoms = (uint8_t)n;
    return true;
}

bool decompose(const Bitmap128& bm, ClassD& out)
{
    ClassD pos{}, neg{};
    Bitmap128 inv;
    for (int i = 0; i < 4; ++i) inv.w[i] = ~bm.w[i];
    bool okp = decompose_set(bm, pos), okn = decompose_set(inv, neg);
    neg.negate = 1;
    if (okp && (!okn || pos.natoms <= neg.natoms)) out = pos;
    else if (okn) out = neg;
    else return false;
    for (unsigned c = 1; c < 128; ++c)
        if (class_has(out, c) != bm.get(c)) return false;  // self-check
    return true;
}

struct Edge { int dst; uint8_t mask; };  // dst = inst id, or -1 for END

struct Lowering {
    const rx::Program& prog;
    const uint8_t* uflags;
    bool ok = true;
    std::string why;
    explicit Lowering(const rx::Program& p, const uint8_t* f) : prog(p), uflags(f) {}

    void fail(const char* w) { if (ok) { ok = false; why = w; } }

    static bool consuming(int op)
    {
        return op == rx::OP_CHAR || op == rx::OP_ANY || op == rx::OP_ANYNL || op == rx::OP_CLASS || op == rx::OP_NCLASS;
    }

    // all consuming instructions / END reachable from `from` through zero-width instructions
    void closure(int from, std::vector<Edge>& out)
    {
        std::set<std::pair<int, int>> seen;
        std::vector<int> path;
        walk(from, 0, seen, path, out);
    }
    void walk(int id, uint8_t mask, std::set<std::pair<int, int>>& seen, std::vector<int>& path, std::vector<Edge>& out)
    {
        if (!ok) return;
        if (std::find(path.begin(), path.end(), id) != path.end()) { fail("zero-width loop"); return; }
        if (!seen.insert({id, mask}).second) return;
        if (seen.size() > 4096) { fail("closure too large"); return; }
        const rx::Inst& in = prog.insts[id];
        if (consuming(in.op)) { out.push_back({id, mask}); return; }
        if (in.op == rx::OP_END) { out.push_back({-1, mask}); return; }
        path.push_back(id);
        switch (in.op) {
        case rx::OP_SPLIT:
            walk(in.other, mask, seen, path, out);
            walk(in.next, mask, seen, path, out);
            break;
        case rx::OP_LBRA: case rx::OP_RBRA: walk(in.next, mask, seen, path, out); break;
        case rx::OP_BOW: walk(in.next, mask | AS_BOW, seen, path, out); break;
        case rx::OP_NBOW: walk(in.next, mask | AS_NBOW, seen, path, out); break;
        case rx::OP_BOL: walk(in.next, mask | (in.arg == '^' ? AS_BOL_CARET : AS_BOL_A), seen, path, out); break;
        case rx::OP_EOL: walk(in.next, mask | (in.arg == '$' ? AS_EOL_DOLLAR : AS_EOL_Z), seen, path, out); break;
        default: break;  // OP_BAD: dead end
        }
        path.pop_back();
    }

    static void simplify(std::vector<Edge>& e)
    {
        std::vector<Edge> keep;
        for (const Edge& x : e) {
            if ((x.mask & AS_BOW) && (x.mask & AS_NBOW)) continue;  // contradiction
            bool dup = false;
            for (const Edge& k : keep) if (k.dst == x.dst && k.mask == x.mask) dup = true;
            if (!dup) keep.push_back(x);
        }
        // an unconditional edge to dst subsumes every conditional one
        std::vector<Edge> out;
        for (const Edge& x : keep) {
            bool subsumed = false;
            for (const Edge& k : keep) if (k.dst == x.dst && k.mask != x.mask && (k.mask & ~x.mask) == 0) subsumed = true;
            if (!subsumed) out.push_back(x);
        }
        e.swap(out);
    }

    bool class_bitmap(const rx::Inst& in, Bitmap128& bm)
    {
        switch (in.op) {
        case rx::OP_CHAR:
            if (in.arg == 0) return false;
            if (in.arg < 128) bm.set(in.arg);  // a multi-byte literal has an empty ASCII set (decided by decoding)
            return true;
        case rx::OP_ANY: for (unsigned c = 0; c < 128; ++c) if (c != '\n') bm.set(c); return true;
        case rx::OP_ANYNL: for (unsigned c = 0; c < 128; ++c) bm.set(c); return true;
        case rx::OP_CLASS: case rx::OP_NCLASS: {
            if (in.arg >= prog.classes.size()) return false;
            for (unsigned c = 0; c < 128; ++c)
                if (rx::class_matches(prog.classes[in.arg], c, uflags) != (in.op == rx::OP_NCLASS)) bm.set(c);
            return true;
        }
        default: return false;
        }
    }
};

}  // namespace

std::shared_ptr<Plan> lower(const rx::Program& prog, bool anchored, const uint8_t* uflags)
{
    if (prog.malformed || prog.insts.empty() || prog.insts.size() > 512) return nullptr;
    Lowering L(prog, uflags);
    std::vector<Edge> from_start;
    L.closure(prog.start_inst, from_start);
    if (!L.ok) return nullptr;
    Lowering::simplify(from_start);
    for (const Edge& e : from_start)
        if (e.dst < 0) return nullptr;  // can match the empty string: stays on the exact VM

    // discover reachable consuming instructions, their out-edges
    std::map<int, std::vector<Edge>> out_edges;
    std::vector<int> work;
    for (const Edge& e : from_start) if (!out_edges.count(e.dst)) { out_edges[e.dst]; work.push_back(e.dst); }
    while (!work.empty()) {
        int id = work.back();
        work.pop_back();
        std::vector<Edge> ed;
        L.closure(prog.insts[id].next, ed);
        if (!L.ok) return nullptr;
        Lowering::simplify(ed);
        for (const Edge& e : ed)
            if (e.dst >= 0 && !out_edges.count(e.dst)) { out_edges[e.dst]; work.push_back(e.dst); }
        out_edges[id] = ed;
        if (out_edges.size() > (size_t)MAX_STEPS) return nullptr;
    }
    // topological order ignoring self loops
    std::map<int, int> indeg;
    for (auto& kv : out_edges) indeg[kv.first];
    for (auto& kv : out_edges)
        for (const Edge& e : kv.second) if (e.dst >= 0 && e.dst != kv.first) ++indeg[e.dst];
    std::vector<int> order, ready;
    for (auto& kv : indeg) if (kv.second == 0) ready.push_back(kv.first);
    while (!ready.empty()) {
        std::sort(ready.begin(), ready.end(), std::greater<int>());
        int id = ready.back();
        ready.pop_back();
        order.push_back(id);
        for (const Edge& e : out_edges[id])
            if (e.dst >= 0 && e.dst != id && --indeg[e.dst] == 0) ready.push_back(e.dst);
    }
    if (order.size() != out_edges.size()) return nullptr;  // a cycle through more than one instruction
    std::map<int, int> step_of;
    for (size_t k = 0; k < order.size(); ++k) step_of[order[k]] = (int)k;

    auto plan = std::make_shared<Plan>();
    PlanDev& P = plan->dev;
    memset(&P, 0, sizeof(P));
    P.anchored = anchored ? 1 : 0;
    P.nsteps = (uint8_t)order.size();
    std::map<std::pair<Bitmap128, std::pair<uint32_t, uint32_t>>, int> class_ids;
    for (size_t k = 0; k < order.size(); ++k) {
        const rx::Inst& in = prog.insts[order[k]];
        Bitmap128 bm;
        if (!L.class_bitmap(in, bm)) return nullptr;
        bm.w[0] &= ~1u;  // NUL is decided by the exact path
        uint32_t na_kind = NA_NEVER, na_arg = 0;
        switch (in.op) {
        case rx::OP_CHAR: if (in.arg >= 128) { na_kind = NA_CHAR_EQ; na_arg = in.arg; } break;
        case rx::OP_ANY: case rx::OP_ANYNL: na_kind = NA_ALWAYS; break;
        case rx::OP_CLASS: na_kind = NA_CLASS; na_arg = in.arg; break;
        case rx::OP_NCLASS: na_kind = NA_NCLASS; na_arg = in.arg; break;
        default: break;
        }
        auto key = std::make_pair(bm, std::make_pair(na_kind, na_arg));
        auto it = class_ids.find(key);
        if (it == class_ids.end()) {
            if (class_ids.size() >= (size_t)MAX_CLASSES) return nullptr;
            ClassD cd{};
            if (!decompose(bm, cd)) return nullptr;
            cd.na_kind = na_kind;
            cd.na_arg = na_arg;
            int id = (int)class_ids.size();
            P.classes[id] = cd;
            it = class_ids.emplace(key, id).first;
        }
        P.steps[k].cls = (uint8_t)it->second;
    }
    P.nclasses = (uint8_t)class_ids.size();
    auto add_pred = [&](int dst_step, uint8_t src, uint8_t mask) {
        StepD& s = P.steps[dst_step];
        if (s.npreds >= MAX_PREDS) return false;
        s.preds[s.npreds++] = PredD{src, mask};
        P.before_needs |= mask;
        return true;
    };
    for (const Edge& e : from_start)
        if (!add_pred(step_of[e.dst], SRC_START, e.mask)) return nullptr;
    for (size_t k = 0; k < order.size(); ++k) {
        int loops = 0;
        for (const Edge& e : out_edges[order[k]]) {
            if (e.dst < 0) {
                if (P.nends >= MAX_ENDS) return nullptr;
                P.ends[P.nends++] = EndD{(uint8_t)k, e.mask};
                P.after_needs |= e.mask;
            } else if (e.dst == order[k]) {
                if (++loops > 1) return nullptr;  // two differently-guarded self loops: keep it simple
                P.steps[k].self_loop = 1;
                P.steps[k].self_mask = e.mask;
                P.before_needs |= e.mask;
            } else if (!add_pred(step_of[e.dst], (uint8_t)k, e.mask))
                return nullptr;
        }
    }
    if (P.nends == 0) return nullptr;
    // ---- linear chain?  Step s is fed by s-1 and, when s-1 is OPTIONAL (x?, x*), by whatever feeds s-1 as well (so the
    // predecessor set of every step is a contiguous run s-1, s-2, ... possibly ending in START); END hangs off the last
    // step and possibly off earlier ones (early exits: x{1,3}, trailing x?).  Unguarded self loops; assertions only on
    // the edges out of START (one common mask) and into END (one common mask).
    {
        bool chain = P.nsteps <= CHAIN_MAX_STEPS && P.nclasses <= CHAIN_MAX_CLASSES;
        bool opt[MAX_STEPS] = {false};
        const int N = P.nsteps;
        auto has_pred = [&](const StepD& st, int src) {
            for (int q = 0; q < st.npreds; ++q)
                if (st.preds[q].src == (uint8_t)src) return true;
            return false;
        };
        auto has_end = [&](int src) {
            for (int q = 0; q < P.nends; ++q)
                if (P.ends[q].src == src) return true;
            return false;
        };
        // opt[j]: step j+1 is also fed by what feeds step j (step j can be skipped)
        for (int j = 0; chain && j + 1 < N; ++j) opt[j] = has_pred(P.steps[j + 1], j == 0 ? (int)SRC_START : j - 1);
        uint8_t start_mask = 0;
        bool start_mask_set = false;
        for (int s2 = 0; chain && s2 < N; ++s2) {
            const StepD& st = P.steps[s2];
            chain = !st.self_loop || st.self_mask == 0;
            // expected predecessors: s2-1, then further back while the steps in between are optional
            int expect = 0;
            for (int j = s2 - 1;; --j) {
                const int src = j < 0 ? (int)SRC_START : j;
                ++expect;
                chain = chain && has_pred(st, src);
                if (j < 0 || !opt[j]) break;
            }
            chain = chain && st.npreds == expect;
            for (int q = 0; chain && q < st.npreds; ++q) {
                if (st.preds[q].src == SRC_START) {
                    if (!start_mask_set) { start_mask = st.preds[q].mask; start_mask_set = true; }
                    chain = st.preds[q].mask == start_mask;
                } else
                    chain = st.preds[q].mask == 0;
            }
        }
        // END: any steps may have an edge into it (early exits: x{1,3}, trailing x?), the last one must; one common mask
        chain = chain && has_end(N - 1);
        for (int q = 1; chain && q < P.nends; ++q) chain = P.ends[q].mask == P.ends[0].mask;
        if (chain) {
            ChainDev& C = plan->chain;
            memset(&C, 0, sizeof(C));
            C.nsteps = P.nsteps;
            C.nclasses = P.nclasses;
            C.anchored = P.anchored;
            C.end_mask = P.ends[0].mask;
            C.needs = P.before_needs | P.after_needs;
            bool any_opt = false;
            for (int s2 = 0; s2 < P.nsteps; ++s2) {
                C.steps[s2] = ChainStepD{P.steps[s2].cls, s2 == 0 ? (uint32_t)start_mask : 0u, P.steps[s2].self_loop, opt[s2] ? 1u : 0u,
                                         has_end(s2) ? 1u : 0u};
                any_opt = any_opt || opt[s2] || (has_end(s2) && s2 + 1 < P.nsteps);
            }
            plan->chain_has_opt = any_opt;
            for (int k = 0; k < P.nclasses; ++k) {
                ChainClassD& cc = C.classes[k];
                const ClassD& src = P.classes[k];
                cc.negate = src.negate;
                cc.na_kind = src.na_kind;
                cc.na_arg = src.na_arg;
                if ((src.na_kind == NA_CLASS || src.na_kind == NA_NCLASS) && src.na_arg < prog.classes.size() &&
                    prog.classes[src.na_arg].ranges.size() <= 8) {
                    const rx::Class& rc = prog.classes[src.na_arg];
                    cc.na_inline = 1;
                    cc.na_builtins = (uint32_t)rc.builtins;
                    cc.na_nranges = (uint32_t)rc.ranges.size();
                    for (size_t r = 0; r < rc.ranges.size(); ++r) cc.na_ranges[r] = rc.ranges[r];
                }
                for (int a2 = 0; a2 < src.natoms; ++a2) {
                    if (src.atoms[a2].kind >= AK_WORD) cc.builtins |= 1u << src.atoms[a2].kind;
                    else cc.atoms[cc.natoms++] = src.atoms[a2];
                }
                C.builtin_union |= cc.builtins;
                // exact 2-byte-character bitmap (covers Latin-1 .. Arabic): decided once here instead of per character
                const rx::Inst& src_inst = prog.insts[order[0]];
                (void)src_inst;
                for (uint32_t ch = 1; ch < 128; ++ch)
                    if (class_has(src, ch)) cc.ascii[ch >> 5] |= 1u << (ch & 31);
                for (uint32_t cp = 0x80; cp < 0x800; ++cp) {
                    const uint32_t packed = ((0xC0u | (cp >> 6)) << 8) | (0x80u | (cp & 0x3Fu));
                    bool in = false;
                    switch (src.na_kind) {
                    case NA_ALWAYS: in = true; break;
                    case NA_CHAR_EQ: in = packed == src.na_arg; break;
                    case NA_CLASS: in = rx::class_matches(prog.classes[src.na_arg], packed, uflags); break;
                    case NA_NCLASS: in = !rx::class_matches(prog.classes[src.na_arg], packed, uflags); break;
                    default: break;
                    }
                    if (in) cc.na2[cp >> 5] |= 1u << (cp & 31);
                }
            }
            for (uint32_t cp = 0x80; cp < 0x800; ++cp)
                if ((uflags[cp] & 15) != 0) C.na2_alnum[cp >> 5] |= 1u << (cp & 31);
            plan->is_chain = true;
            // span fast path: no loop before the last step, and the last step's loop (if any) must be greedy, i.e. the
            // SPLIT that follows the instruction tries the loop body first (regcomp PLUS, not PLUS_LAZY)
            bool span_ok = !any_opt;
            for (int s2 = 0; s2 + 1 < P.nsteps; ++s2) span_ok = span_ok && !P.steps[s2].self_loop;
            if (span_ok && P.steps[P.nsteps - 1].self_loop) {
                const int last = order[P.nsteps - 1];
                const rx::Inst& nx = prog.insts[prog.insts[last].next];
                span_ok = nx.op == rx::OP_SPLIT && nx.other == last;
            }
            plan->span_ok = span_ok;
        }
    }
    plan->text = describe(*plan);
    return plan;
}

std::string describe(const Plan& plan)
{
    const PlanDev& P = plan.dev;
    static const char* an[] = {"EQ", "RANGE", "WORD", "ALNUM", "DIGIT", "SPACE", "LOWER", "UPPER", "ANY"};
    std::ostringstream o;
    o << (plan.is_chain ? "chain " : "dag ") << (P.anchored ? "anchored " : "") << "classes=" << (int)P.nclasses << " steps=" << (int)P.nsteps << " ends=" << (int)P.nends << " {";
    for (int k = 0; k < P.nclasses; ++k) {
        o << " C" << k << "=" << (P.classes[k].negate ? "!" : "") << "(";
        for (int a = 0; a < P.classes[k].natoms; ++a) {
            const AtomD& t = P.classes[k].atoms[a];
            o << (a ? "|" : "") << an[t.kind];
            if (t.kind == AK_EQ) o << ":" << (int)t.lo;
            if (t.kind == AK_RANGE) o << ":" << (int)t.lo << "-" << (int)t.hi;
        }
        o << ")";
    }
    o << " ;";
    for (int s = 0; s < P.nsteps; ++s) {
        o << " S" << s << "=C" << (int)P.steps[s].cls << "[";
        for (int p = 0; p < P.steps[s].npreds; ++p) {
            if (p) o << ",";
            if (P.steps[s].preds[p].src == SRC_START) o << "^"; else o << "S" << (int)P.steps[s].preds[p].src;
            if (P.steps[s].preds[p].mask) o << "/" << (int)P.steps[s].preds[p].mask;
        }
        o << "]";
        if (P.steps[s].self_loop) { o << "*"; if (P.steps[s].self_mask) o << "/" << (int)P.steps[s].self_mask; }
    }
    o << " ; end:";
    for (int e = 0; e < P.nends; ++e) { o << " S" << (int)P.ends[e].src; if (P.ends[e].mask) o << "/" << (int)P.ends[e].mask; }
    o << " }";
    return o.str();
}

const PlanDev& device_plan(const Plan& plan) { return plan.dev; }
const ChainDev* device_chain(const Plan& plan) { return plan.is_chain ? &plan.chain : nullptr; }
// chain whose only loop (if any) is a greedy loop on its last step: leftmost start + longest admissible end == the Pike
// VM's span
const ChainDev* span_chain(const Plan& plan) { return plan.is_chain && plan.span_ok ? &plan.chain : nullptr; }

// ---- plain host executor (tests/sim only) ------------------------------------------------------------------------
// Plain host executor of the CHAIN model (ChainDev: steps with loop / opt / exit flags, one leading and one trailing
// assertion mask) for pure-ASCII rows — the model k_chain64 evaluates with bit streams (tests/sim only).  Returns false
// when the plan is not a chain.  Same conventions as reference_execute below.
bool reference_execute_chain(const Plan& plan, const char* chars, const int32_t* offsets, int32_t n, uint8_t* out, uint8_t* dirty)
{
    if (!plan.is_chain) return false;
    const ChainDev& C = plan.chain;
    const AtomD alnum{AK_ALNUM, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const uint8_t* s = (const uint8_t*)chars + offsets[i];
        const int len = offsets[i + 1] - offsets[i];
        out[i] = 0;
        dirty[i] = 0;
        for (int p = 0; p < len; ++p)
            if (s[p] >= 0x80 || s[p] == 0) dirty[i] = 1;
        if (dirty[i]) continue;
        auto is_al = [&](int p) { return p >= 0 && p < len && atom_has(alnum, s[p]); };
        auto is_nl = [&](int p) { return p >= 0 && p < len && s[p] == '\n'; };
        auto holds = [&](uint32_t m, int q) {  // assertions between byte q-1 and byte q (q = len: end of the row)
            const bool bow = is_al(q) != is_al(q - 1);
            if ((m & AS_BOW) && !bow) return false;
            if ((m & AS_NBOW) && bow) return false;
            if ((m & AS_BOL_CARET) && !(q == 0 || is_nl(q - 1))) return false;
            if ((m & AS_BOL_A) && q != 0) return false;
            if ((m & AS_EOL_DOLLAR) && !(q == len || is_nl(q))) return false;
            if ((m & AS_EOL_Z) && q != len) return false;
            return true;
        };
        // ready[q]: step s may start consuming at byte q; after the step: fin[q] = it has just consumed byte q
        std::vector<uint8_t> ready(len + 1, 0), fin(len, 0), done(len, 0), next(len + 1, 0);
        for (int q = 0; q <= len; ++q) ready[q] = (C.anchored ? q == 0 : true) && holds(C.steps[0].before, q);
        for (uint32_t st = 0; st < C.nsteps; ++st) {
            const ChainClassD& cc = C.classes[C.steps[st].cls];
            std::fill(fin.begin(), fin.end(), 0);
            for (int q = 0; q < len; ++q) {
                const bool in_class = (cc.ascii[s[q] >> 5] >> (s[q] & 31)) & 1u;
                bool v = ready[q] && in_class;
                if (!v && C.steps[st].loop && in_class && q > 0 && fin[q - 1]) v = true;
                fin[q] = v;
            }
            if (C.steps[st].exit)
                for (int q = 0; q < len; ++q) done[q] |= fin[q];
            std::fill(next.begin(), next.end(), 0);
            for (int q = 0; q < len; ++q)
                if (fin[q]) next[q + 1] = 1;
            if (C.steps[st].opt)
                for (int q = 0; q <= len; ++q) next[q] |= ready[q];
            ready = next;
        }
        for (int q = 0; q < len && !out[i]; ++q)
            if (done[q] && holds(C.end_mask, q + 1)) out[i] = 1;
    }
    return true;
}

void reference_execute(const Plan& plan, const char* chars, const int32_t* offsets, const uint8_t* validity, int32_t n,
                       uint8_t* out, uint8_t* dirty)
{
    const PlanDev& P = plan.dev;
    const int32_t base = n ? offsets[0] : 0;
    const int64_t N = n ? offsets[n] - base : 0;
    const uint8_t* s = (const uint8_t*)chars + base;
    std::vector<uint8_t> RS(N + 1, 0), A(N, 0), NL(N, 0);
    auto valid = [&](int i) { return !validity || ((validity[i >> 3] >> (i & 7)) & 1); };
    for (int i = 0; i < n; ++i) {
        out[i] = 0;
        dirty[i] = 0;
        int b = offsets[i] - base, e = offsets[i + 1] - base;
        if (e > b) RS[b] = 1;
        for (int p = b; p < e; ++p) if (s[p] >= 0x80 || s[p] == 0) dirty[i] = 1;
        (void)valid;
    }
    RS[N] = 1;
    const AtomD alnum{AK_ALNUM, 0, 0, 0};
    for (int64_t p = 0; p < N; ++p) { A[p] = s[p] < 0x80 && atom_has(alnum, s[p]); NL[p] = s[p] == '\n'; }
    auto before = [&](uint8_t m, int64_t q) {  // assertions between q-1 and q
        bool aprev = q > 0 && !RS[q] && A[q - 1];
        bool nlprev = q > 0 && !RS[q] && NL[q - 1];
        bool bow = (A[q] != 0) != aprev;
        if ((m & AS_BOW) && !bow) return false;
        if ((m & AS_NBOW) && bow) return false;
        if ((m & AS_BOL_CARET) && !(RS[q] || nlprev)) return false;
        if ((m & AS_BOL_A) && !RS[q]) return false;
        if ((m & AS_EOL_DOLLAR) && !NL[q]) return false;
        if (m & AS_EOL_Z) return false;
        return true;
    };
    auto after = [&](uint8_t m, int64_t p) {  // assertions between p and p+1
        bool last = RS[p + 1] != 0;
        bool anext = !last && A[p + 1];
        bool nlnext = !last && NL[p + 1];
        bool bow = (A[p] != 0) != anext;
        if ((m & AS_BOW) && !bow) return false;
        if ((m & AS_NBOW) && bow) return false;
        if ((m & AS_BOL_CARET) && !NL[p]) return false;
        if (m & AS_BOL_A) return false;
        if ((m & AS_EOL_DOLLAR) && !(last || nlnext)) return false;
        if ((m & AS_EOL_Z) && !last) return false;
        return true;
    };
    std::vector<std::vector<uint8_t>> M(P.nsteps, std::vector<uint8_t>(N, 0));
    for (int k = 0; k < P.nsteps; ++k) {
        const StepD& st = P.steps[k];
        const ClassD& cd = P.classes[st.cls];
        for (int64_t q = 0; q < N; ++q) {
            bool in_class = s[q] < 0x80 && class_has(cd, s[q]);
            bool entry = false;
            for (int j = 0; j < st.npreds && !entry; ++j) {
                const PredD& pr = st.preds[j];
                bool t = pr.src == SRC_START ? (P.anchored ? RS[q] != 0 : true) : (q > 0 && !RS[q] && M[pr.src][q - 1]);
                entry = t && before(pr.mask, q);
            }
            bool v = entry && in_class;
            if (!v && st.self_loop && in_class && q > 0 && !RS[q] && M[k][q - 1] && before(st.self_mask, q)) v = true;
            M[k][q] = v;
        }
    }
    std::vector<uint8_t> E(N, 0);
    for (int e = 0; e < P.nends; ++e)
        for (int64_t p = 0; p < N; ++p)
            if (M[P.ends[e].src][p] && after(P.ends[e].mask, p)) E[p] = 1;
    for (int i = 0; i < n; ++i)
        for (int p = offsets[i] - base; p < offsets[i + 1] - base; ++p)
            if (E[p]) { out[i] = 1; break; }
}

}  // namespace bits
}  // namespace custr
