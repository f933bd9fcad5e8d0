// Host regex compiler: UTF-8 pattern -> rx::Program (see regex_prog.h).
//
// Behavioural contract = the reference compiler cpp/src/regex/regcomp.cpp (lexer :314-539, class builder
// :170-312, counted-repeat expansion :772-885, operator-precedence compile :584-951, NOP removal :968-1037,
// start-set extraction :1040-1061), including its documented quirks (SURVEY.md §7 "Exact quirk parity"):
//   - every \NNN is octal and swallows the character that follows the last digit (:329-339)
//   - \xHH drops hex digits 'a'/'A' (`a > 'a'`, :362-366)
//   - negated classes never match '\n' (:184-185); \W and \D carry the same exclusion (:386-391,:436-441)
//   - {n,m} is expanded on the TOKEN stream by duplication, capture groups included (:772-885)
//   - malformed constructs do not raise: missing operands become no-ops (:616-621)
// The graph that comes out has the same topology as the reference's Reinst graph, which is what fixes thread
// priority in the Pike VM and therefore the match spans seen by count_re / replace_re.
// This is a restatement in this repo's own data structures, not a copy: tokens are a tagged struct, operators
// carry explicit precedence, fragments are (first,last) pairs over rx::Inst.
#include "regex_prog.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace custr {
namespace rx {

namespace {

// ---- UTF-8 helpers (reference custring_view.inl:48-57,1724-1744) -----------------------------------------
inline int lead_width(uint8_t b)
{
    int w = 1;
    w += (b & 0xF0) == 0xF0;
    w += (b & 0xE0) == 0xE0;
    w += (b & 0xC0) == 0xC0;
    w -= (b & 0xC0) == 0x80;
    return w;
}

std::vector<uint32_t> to_packed(const char* s)
{
    std::vector<uint32_t> out;
    size_t n = strlen(s);
    size_t i = 0;
    while (i < n) {
        uint8_t b = (uint8_t)s[i];
        int w = lead_width(b);
        uint32_t c = b;
        for (int k = 1; k < w && i + k < n; ++k) c = (c << 8) | (uint8_t)s[i + k];
        out.push_back(c);
        i += w > 0 ? w : 1;
    }
    out.push_back(0);
    return out;
}

// ---- tokens ---------------------------------------------------------------------------------------------
enum Kind : int {
    // operands
    K_CHAR, K_ANY, K_CLASS, K_NCLASS, K_BOL, K_EOL, K_BOW, K_NBOW, K_BAD,
    // operators; numeric order == binding strength used by reduce_until()
    K_OPS = 100,
    K_START = 100, K_RBRA, K_LBRA, K_LBRA_NC, K_OR, K_CAT, K_STAR, K_STAR_LAZY, K_PLUS, K_PLUS_LAZY, K_QUEST,
    K_QUEST_LAZY, K_COUNTED, K_COUNTED_LAZY,
    K_EOF = 1000
};
inline bool is_operator(int k) { return k >= K_OPS && k < K_EOF; }

struct Tok {
    int kind = K_EOF;
    uint32_t ch = 0;  // K_CHAR / K_BOL / K_EOL payload
    int cls = -1;     // K_CLASS / K_NCLASS payload
    int lo = 0, hi = 0;  // K_COUNTED payload (hi < 0: unbounded)
};

class Lexer {
public:
    Lexer(const std::vector<uint32_t>& pat, Program& prog) : p_(pat), prog_(prog) {}

    std::vector<Tok> run(bool& has_counted)
    {
        std::vector<Tok> toks;
        has_counted = false;
        for (;;) {
            Tok t = next();
            if (t.kind == K_EOF) break;
            if (t.kind == K_COUNTED || t.kind == K_COUNTED_LAZY) has_counted = true;
            toks.push_back(t);
        }
        return toks;
    }

private:
    const std::vector<uint32_t>& p_;
    Program& prog_;
    size_t i_ = 0;
    bool done_ = false;
    int id_w_ = -1, id_W_ = -1, id_s_ = -1, id_d_ = -1, id_D_ = -1;

    uint32_t peek(size_t k = 0) const { return i_ + k < p_.size() ? p_[i_ + k] : 0; }
    // raw read that never runs past the terminator (the reference would read out of bounds there)
    uint32_t take() { return i_ < p_.size() ? p_[i_++] : 0; }

    // returns true when the character was backslash-escaped (or the input is exhausted)
    bool read(uint32_t& c)
    {
        if (done_) { c = 0; return true; }
        c = take();
        if (c == '\\') { c = take(); return true; }
        if (c == 0) done_ = true;
        return false;
    }

    int shared_class(int& slot, int builtins, bool with_newline)
    {
        if (slot < 0) {
            Class c;
            c.builtins = builtins;
            if (with_newline) { c.ranges.push_back('\n'); c.ranges.push_back('\n'); }
            prog_.classes.push_back(c);
            slot = (int)prog_.classes.size() - 1;
        }
        return slot;
    }

    Tok make(int kind, uint32_t ch = 0) { Tok t; t.kind = kind; t.ch = ch; return t; }
    Tok make_cls(int kind, int id) { Tok t; t.kind = kind; t.cls = id; return t; }

    Tok next()
    {
        uint32_t c;
        bool quoted = read(c);
        if (quoted) {
            if (c == 0) return make(K_EOF);
            if (c >= '0' && c <= '7') {  // always octal; eats one char past the digits
                uint32_t v = c - '0';
                uint32_t d = take();
                while (d >= '0' && d <= '7') { v = (v << 3) | (d - '0'); d = take(); }
                return make(K_CHAR, v);
            }
            switch (c) {
            case 't': return make(K_CHAR, '\t');
            case 'n': return make(K_CHAR, '\n');
            case 'r': return make(K_CHAR, '\r');
            case 'a': return make(K_CHAR, 0x07);
            case 'f': return make(K_CHAR, 0x0C);
            case 'x': {
                uint32_t a = take(), b = take(), v = 0;
                if (a >= '0' && a <= '9') v += (a - '0') << 4;
                else if (a > 'a' && a <= 'f') v += (a - 'a' + 10) << 4;
                else if (a > 'A' && a <= 'F') v += (a - 'A' + 10) << 4;
                if (b >= '0' && b <= '9') v += b - '0';
                else if (b > 'a' && b <= 'f') v += b - 'a' + 10;
                else if (b > 'A' && b <= 'F') v += b - 'A' + 10;
                return make(K_CHAR, v);
            }
            case 'w': return make_cls(K_CLASS, shared_class(id_w_, CB_W, false));
            case 'W': return make_cls(K_NCLASS, shared_class(id_W_, CB_W, true));
            case 's': return make_cls(K_CLASS, shared_class(id_s_, CB_S, false));
            case 'S': return make_cls(K_NCLASS, shared_class(id_s_, CB_S, false));
            case 'd': return make_cls(K_CLASS, shared_class(id_d_, CB_D, false));
            case 'D': return make_cls(K_NCLASS, shared_class(id_D_, CB_D, true));
            case 'b': return make(K_BOW);
            case 'B': return make(K_NBOW);
            case 'A': return make(K_BOL, 'A');
            case 'Z': return make(K_EOL, 'Z');
            }
            return make(K_CHAR, c);
        }
        switch (c) {
        case 0: return make(K_EOF);
        case '*': if (peek() == '?') { take(); return make(K_STAR_LAZY); } return make(K_STAR);
        case '?': if (peek() == '?') { take(); return make(K_QUEST_LAZY); } return make(K_QUEST);
        case '+': if (peek() == '?') { take(); return make(K_PLUS_LAZY); } return make(K_PLUS);
        case '{': {
            Tok t;
            if (counted(t)) return t;
            break;
        }
        case '|': return make(K_OR);
        case '.': return make(K_ANY);
        case '(':
            if (peek() == '?' && peek(1) == ':') { take(); take(); return make(K_LBRA_NC); }
            return make(K_LBRA);
        case ')': return make(K_RBRA);
        case '^': return make(K_BOL, '^');
        case '$': return make(K_EOL, '$');
        case '[': return bracket();
        }
        return make(K_CHAR, c);
    }

    // "{n}", "{n,}", "{n,m}" (+ optional '?'): at most 7 characters per number, "%hd" conversion
    static int short_of(const std::string& s, int fallback)
    {
        short v = 0;
        return sscanf(s.c_str(), "%hd", &v) == 1 ? (int)v : fallback;
    }
    bool counted(Tok& out)
    {
        if (peek() < '0' || peek() > '9') return false;
        size_t save = i_;
        std::string a;
        for (int k = 0; k < 7 && peek() != '}' && peek() != ',' && peek() != 0; ++k) a.push_back((char)take());
        if (peek() != '}' && peek() != ',') { i_ = save; return false; }
        int lo = short_of(a, 0), hi;
        if (peek() != ',') hi = lo;
        else {
            hi = -1;
            take();
            std::string b;
            for (int k = 0; k < 7 && peek() != '}' && peek() != 0; ++k) b.push_back((char)take());
            if (peek() != '}') { i_ = save; return false; }
            if (!b.empty()) hi = short_of(b, -1);
        }
        take();  // '}'
        out.kind = K_COUNTED;
        if (peek() == '?') { take(); out.kind = K_COUNTED_LAZY; }
        out.lo = lo;
        out.hi = hi;
        return true;
    }

    Tok bracket()
    {
        int kind = K_CLASS;
        std::vector<std::pair<uint32_t, uint32_t>> spans;
        int builtins = 0;
        uint32_t c;
        bool quoted = read(c);
        if (!quoted && c == '^') {
            kind = K_NCLASS;
            quoted = read(c);
            spans.push_back({'\n', '\n'});
        }
        for (int seen = 1;; ++seen) {
            if (c == 0) return make(K_BAD);
            if (quoted) {
                bool builtin = true;
                switch (c) {
                case 'n': c = '\n'; builtin = false; break;
                case 'r': c = '\r'; builtin = false; break;
                case 't': c = '\t'; builtin = false; break;
                case 'a': c = 0x07; builtin = false; break;
                case 'b': c = 0x08; builtin = false; break;
                case 'f': c = 0x0C; builtin = false; break;
                case 'w': builtins |= CB_W; break;
                case 's': builtins |= CB_S; break;
                case 'd': builtins |= CB_D; break;
                case 'W': builtins |= CB_NW; break;
                case 'S': builtins |= CB_NS; break;
                case 'D': builtins |= CB_ND; break;
                default: builtin = false; break;
                }
                if (builtin) { quoted = read(c); continue; }
            }
            if (!quoted && c == ']' && seen > 1) break;
            if (!quoted && c == '-') {
                if (spans.empty()) return make(K_BAD);
                quoted = read(c);
                if ((!quoted && c == ']') || c == 0) return make(K_BAD);
                spans.back().second = c;
            } else
                spans.push_back({c, c});
            quoted = read(c);
        }
        std::stable_sort(spans.begin(), spans.end(),
                         [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) { return a.first < b.first; });
        Class cls;
        cls.builtins = builtins;
        for (size_t k = 0; k < spans.size(); ++k) {
            if (!cls.ranges.empty() && spans[k].first <= cls.ranges.back() + 1) {
                if (spans[k].second >= cls.ranges.back()) cls.ranges.back() = spans[k].second;
            } else {
                cls.ranges.push_back(spans[k].first);
                cls.ranges.push_back(spans[k].second);
            }
        }
        prog_.classes.push_back(cls);
        return make_cls(kind, (int)prog_.classes.size() - 1);
    }
};

// ---- {n,m} expansion on the token stream ----------------------------------------------------------------
std::vector<Tok> expand_counted(const std::vector<Tok>& in)
{
    std::vector<Tok> out;
    std::vector<int> open;
    int atom = -1;  // index in `in` where the repeatable atom starts
    for (int i = 0; i < (int)in.size(); ++i) {
        const Tok& t = in[i];
        if (t.kind != K_COUNTED && t.kind != K_COUNTED_LAZY) {
            out.push_back(t);
            if (t.kind == K_LBRA || t.kind == K_LBRA_NC) { open.push_back(i); atom = -1; }
            else if (t.kind == K_RBRA) {
                if (open.empty()) return out;  // the reference indexes an empty stack here; treat as broken
                atom = open.back();
                open.pop_back();
            } else if (!is_operator(t.kind))
                atom = i;
            continue;
        }
        if (atom < 0) return out;  // broken regex: stop expanding
        bool lazy = t.kind == K_COUNTED_LAZY;
        if (t.lo <= 0) {
            for (int j = 0; j < i - atom && !out.empty(); ++j) out.pop_back();
        } else {
            for (int j = 1; j < t.lo; ++j)
                for (int k = atom; k < i; ++k) out.push_back(in[k]);
        }
        Tok op;
        if (t.hi >= 0) {
            for (int j = t.lo; j < t.hi; ++j) {
                op.kind = K_LBRA_NC;
                out.push_back(op);
                for (int k = atom; k < i; ++k) out.push_back(in[k]);
            }
            for (int j = t.lo; j < t.hi; ++j) {
                op.kind = K_RBRA;
                out.push_back(op);
                op.kind = lazy ? K_QUEST_LAZY : K_QUEST;
                out.push_back(op);
            }
        } else if (t.lo > 0) {
            op.kind = lazy ? K_PLUS_LAZY : K_PLUS;
            out.push_back(op);
        } else {
            for (int k = atom; k < i; ++k) out.push_back(in[k]);
            op.kind = lazy ? K_STAR_LAZY : K_STAR;
            out.push_back(op);
        }
    }
    return out;
}

// ---- operator-precedence compile to an instruction graph ------------------------------------------------
class Builder {
public:
    explicit Builder(Program& p) : prog_(p) {}

    void run(const std::vector<Tok>& toks)
    {
        ops_.push_back({K_START - 1, 0});  // sentinel below every operator
        for (const Tok& t : toks) {
            int kind = t.kind;
            if (kind == K_LBRA) { ++groups_; pending_group_ = groups_; }
            else if (kind == K_LBRA_NC) { pending_group_ = 0; kind = K_LBRA; }
            if (is_operator(kind)) on_operator(kind);
            else on_operand(t, kind);
        }
        reduce_until(K_START);
        Tok end;
        end.kind = K_EOF;
        on_operand(end, K_EOF);
        reduce_until(K_START);
        prog_.start_inst = frags_.empty() ? 0 : frags_.back().first;
        prog_.ngroups = groups_;
    }

private:
    struct Frag { int first, last; };
    struct Pending { int kind; int group; };
    Program& prog_;
    std::vector<Frag> frags_;
    std::vector<Pending> ops_;
    bool last_was_operand_ = false;
    int depth_ = 0, groups_ = 0, pending_group_ = 0;

    int emit(int op, uint32_t arg = 0)
    {
        prog_.insts.push_back(Inst{op, arg, 0, 0});
        return (int)prog_.insts.size() - 1;
    }
    Inst& at(int id) { return prog_.insts[id]; }

    Frag pop_frag()
    {
        if (frags_.empty()) {  // missing operand
            int n = emit(OP_NOP);
            frags_.push_back({n, n});
        }
        Frag f = frags_.back();
        frags_.pop_back();
        return f;
    }

    void on_operand(const Tok& t, int kind)
    {
        if (last_was_operand_) on_operator(K_CAT);
        int id;
        switch (kind) {
        case K_CHAR: id = emit(OP_CHAR, t.ch); break;
        case K_ANY: id = emit(OP_ANY); break;
        case K_CLASS: id = emit(OP_CLASS, (uint32_t)t.cls); break;
        case K_NCLASS: id = emit(OP_NCLASS, (uint32_t)t.cls); break;
        case K_BOL: id = emit(OP_BOL, t.ch); break;
        case K_EOL: id = emit(OP_EOL, t.ch); break;
        case K_BOW: id = emit(OP_BOW); break;
        case K_NBOW: id = emit(OP_NBOW); break;
        case K_EOF: id = emit(OP_END); break;
        default: id = emit(OP_BAD); prog_.malformed = true; break;
        }
        frags_.push_back({id, id});
        last_was_operand_ = true;
    }

    void on_operator(int kind)
    {
        if (kind == K_RBRA && --depth_ < 0) return;  // unmatched ')'
        if (kind == K_LBRA) {
            ++depth_;
            if (last_was_operand_) on_operator(K_CAT);
        } else
            reduce_until(kind);
        if (kind != K_RBRA) ops_.push_back({kind, pending_group_});
        last_was_operand_ = kind == K_STAR || kind == K_QUEST || kind == K_PLUS || kind == K_STAR_LAZY ||
                            kind == K_QUEST_LAZY || kind == K_PLUS_LAZY || kind == K_RBRA;
    }

    void reduce_until(int prio)
    {
        while (!ops_.empty() && (prio == K_RBRA || ops_.back().kind >= prio)) {
            Pending op = ops_.back();
            ops_.pop_back();
            switch (op.kind) {
            case K_LBRA: {  // closed by the ')' being processed
                Frag body = pop_frag();
                int r = emit(OP_RBRA, (uint32_t)op.group);
                at(body.last).next = r;
                int l = emit(OP_LBRA, (uint32_t)op.group);
                at(l).next = body.first;
                frags_.push_back({l, r});
                return;
            }
            case K_OR: {
                Frag rhs = pop_frag(), lhs = pop_frag();
                int join = emit(OP_NOP);
                at(rhs.last).next = join;
                at(lhs.last).next = join;
                int s = emit(OP_SPLIT);
                at(s).other = lhs.first;  // left alternative has priority
                at(s).next = rhs.first;
                frags_.push_back({s, join});
                break;
            }
            case K_CAT: {
                Frag rhs = pop_frag(), lhs = pop_frag();
                at(lhs.last).next = rhs.first;
                frags_.push_back({lhs.first, rhs.last});
                break;
            }
            case K_STAR: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT);
                at(body.last).next = s;
                at(s).other = body.first;  // greedy: loop body first; `next` is patched by concatenation
                frags_.push_back({s, s});
                break;
            }
            case K_STAR_LAZY: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT), out = emit(OP_NOP);
                at(body.last).next = s;
                at(s).next = body.first;
                at(s).other = out;  // lazy: leave first
                frags_.push_back({s, out});
                break;
            }
            case K_PLUS: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT);
                at(body.last).next = s;
                at(s).other = body.first;
                frags_.push_back({body.first, s});
                break;
            }
            case K_PLUS_LAZY: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT), out = emit(OP_NOP);
                at(body.last).next = s;
                at(s).next = body.first;
                at(s).other = out;
                frags_.push_back({body.first, out});
                break;
            }
            case K_QUEST: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT), out = emit(OP_NOP);
                at(s).next = out;
                at(s).other = body.first;
                at(body.last).next = out;
                frags_.push_back({s, out});
                break;
            }
            case K_QUEST_LAZY: {
                Frag body = pop_frag();
                int s = emit(OP_SPLIT), out = emit(OP_NOP);
                at(s).next = body.first;
                at(s).other = out;
                at(body.last).next = out;
                frags_.push_back({s, out});
                break;
            }
            default: break;  // sentinel / stray counted operator: ignored
            }
        }
    }
};

// ---- NOP removal + start set ----------------------------------------------------------------------------
void strip_nops(Program& p)
{
    std::vector<Inst>& v = p.insts;
    const int n = (int)v.size();
    for (Inst& in : v)
        if ((in.op == OP_LBRA || in.op == OP_RBRA) && (int)in.arg < 1) in.op = OP_NOP;
    auto resolve = [&](int id) {
        for (int guard = 0; guard <= n && v[id].op == OP_NOP; ++guard) id = v[id].next;
        return id;
    };
    for (Inst& in : v) {
        if (in.op == OP_NOP) continue;
        in.next = resolve(in.next);
        if (in.op == OP_SPLIT) in.other = resolve(in.other);
    }
    p.start_inst = resolve(p.start_inst);
    std::vector<int> remap(n);
    int live = 0;
    for (int i = 0; i < n; ++i) {
        remap[i] = live;
        if (v[i].op != OP_NOP) v[live++] = v[i];
    }
    v.resize(live);
    for (Inst& in : v) {
        in.next = remap[in.next];
        if (in.op == OP_SPLIT) in.other = remap[in.other];
    }
    p.start_inst = remap[p.start_inst];
}

void collect_starts(Program& p)
{
    p.starts.clear();
    if (p.insts.empty()) return;
    std::vector<int> stack{p.start_inst};
    size_t guard = 0;
    while (!stack.empty() && guard++ < 4 * p.insts.size() + 16) {
        int id = stack.back();
        stack.pop_back();
        const Inst& in = p.insts[id];
        if (in.op == OP_SPLIT) {
            stack.push_back(in.next);
            stack.push_back(in.other);  // popped first: higher priority
        } else
            p.starts.push_back(id);
    }
}

}  // namespace

Program compile(const char* pattern_utf8)
{
    Program prog;
    std::vector<uint32_t> pat = to_packed(pattern_utf8 ? pattern_utf8 : "");
    bool has_counted = false;
    Lexer lex(pat, prog);
    std::vector<Tok> toks = lex.run(has_counted);
    if (has_counted) toks = expand_counted(toks);
    Builder b(prog);
    b.run(toks);
    strip_nops(prog);
    collect_starts(prog);
    return prog;
}

uint32_t packed_to_codepoint(uint32_t c)  // reference util.inl:51-75
{
    if (c < 0x80u) return c;
    if (c < 0xE000u) return ((c & 0x1F00u) >> 2) | (c & 0x3Fu);
    if (c < 0xF00000u) return ((c & 0x0F0000u) >> 4) | ((c & 0x3F00u) >> 2) | (c & 0x3Fu);
    if (c <= 0xF8000000u) return ((c & 0x03000000u) >> 6) | ((c & 0x3F0000u) >> 4) | ((c & 0x3F00u) >> 2) | (c & 0x3Fu);
    return 0;
}

bool class_matches(const Class& c, uint32_t ch, const uint8_t* fl)  // reference regexec.inl:127-155
{
    for (size_t i = 0; i + 1 < c.ranges.size(); i += 2)
        if (ch >= c.ranges[i] && ch <= c.ranges[i + 1]) return true;
    if (!c.builtins) return false;
    uint32_t cp = packed_to_codepoint(ch);
    if (cp > 0xFFFFu) return false;
    uint8_t f = fl[cp];
    bool alnum = (f & 15) != 0, space = (f & 16) != 0, digit = (f & 4) != 0;
    if ((c.builtins & CB_W) && (ch == '_' || alnum)) return true;
    if ((c.builtins & CB_S) && space) return true;
    if ((c.builtins & CB_D) && digit) return true;
    if ((c.builtins & CB_NW) && ch != '\n' && ch != '_' && !alnum) return true;
    if ((c.builtins & CB_NS) && !space) return true;
    if ((c.builtins & CB_ND) && ch != '\n' && !digit) return true;
    return false;
}

std::vector<uint8_t> serialize(const Program& p, const uint8_t* fl)
{
    DevHeader h{};
    h.ninsts = (int32_t)p.insts.size();
    h.nstarts = (int32_t)p.starts.size();
    h.nclasses = (int32_t)p.classes.size();
    h.start_inst = p.start_inst;
    h.ngroups = p.ngroups;
    std::vector<DevClass> dc(p.classes.size());
    std::vector<uint32_t> ranges;
    for (size_t k = 0; k < p.classes.size(); ++k) {
        DevClass& d = dc[k];
        memset(&d, 0, sizeof(d));
        d.builtins = p.classes[k].builtins;
        d.range_begin = (int32_t)ranges.size();
        d.range_count = (int32_t)p.classes[k].ranges.size();
        ranges.insert(ranges.end(), p.classes[k].ranges.begin(), p.classes[k].ranges.end());
        for (uint32_t c = 0; c < 128; ++c)
            if (class_matches(p.classes[k], c, fl)) d.ascii[c >> 5] |= 1u << (c & 31);
    }
    h.nranges = (int32_t)ranges.size();
    if (!p.insts.empty()) {
        const Inst& s = p.insts[p.start_inst];
        if (s.op == OP_CHAR || s.op == OP_BOL) { h.start_op = s.op; h.start_arg = s.arg; }
    }
    std::vector<uint8_t> img(sizeof(h) + sizeof(Inst) * p.insts.size() + 4 * p.starts.size() + sizeof(DevClass) * dc.size() +
                             4 * ranges.size());
    uint8_t* w = img.data();
    auto put = [&](const void* src, size_t n) { if (n) memcpy(w, src, n); w += n; };
    put(&h, sizeof(h));
    put(p.insts.data(), sizeof(Inst) * p.insts.size());
    put(p.starts.data(), 4 * p.starts.size());
    put(dc.data(), sizeof(DevClass) * dc.size());
    put(ranges.data(), 4 * ranges.size());
    return img;
}

std::string Program::describe() const
{
    static const char* names[] = {"?", "CHAR", "ANY", "ANYNL", "CLASS", "NCLASS", "END", "LBRA", "RBRA", "BOL", "EOL",
                                  "BOW", "NBOW", "SPLIT", "NOP", "BAD"};
    std::ostringstream o;
    for (size_t i = 0; i < insts.size(); ++i) {
        const Inst& in = insts[i];
        o << i << ": " << names[in.op >= 0 && in.op <= OP_BAD ? in.op : 0];
        if (in.op == OP_CHAR || in.op == OP_BOL || in.op == OP_EOL) o << " 0x" << std::hex << in.arg << std::dec;
        if (in.op == OP_CLASS || in.op == OP_NCLASS || in.op == OP_LBRA || in.op == OP_RBRA) o << " " << in.arg;
        if (in.op == OP_SPLIT) o << " first=" << in.other << " then=" << in.next;
        else if (in.op != OP_END) o << " -> " << in.next;
        o << "\n";
    }
    o << "start=" << start_inst << " starts=";
    for (int s : starts) o << s << " ";
    o << "groups=" << ngroups << "\n";
    for (size_t k = 0; k < classes.size(); ++k) {
        o << "class " << k << ": builtins=" << classes[k].builtins << " ranges=";
        for (size_t j = 0; j + 1 < classes[k].ranges.size(); j += 2)
            o << std::hex << "[" << classes[k].ranges[j] << "-" << classes[k].ranges[j + 1] << "]" << std::dec;
        o << "\n";
    }
    return o.str();
}

}  // namespace rx
}  // namespace custr
