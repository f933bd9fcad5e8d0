// C-ABI harness around the UNMODIFIED reference classes (NVStrings / NVCategory / NVText from
// /root/reference/cpp/include), compiled for the host by oracle/Makefile.  TEST INFRASTRUCTURE ONLY:
// it is the parity oracle for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
// Nothing under custrings_b200/ links, loads or calls it.
//
// All buffers are host buffers (the reference's "device" memory is malloc in this build), columns
// travel as Arrow-style (chars, int32 offsets[n+1], LSB-first validity bits).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <NVStrings.h>
#include <NVCategory.h>
#include <NVText.h>
#include "regex/regcomp.h"   // reference's host regex compiler (cpp/src/regex/regcomp.h), for ref_regex_dump

char32_t* to_char32(const char* ca);  // cpp/src/strings/NVStringsImpl.cu:49

static thread_local std::string g_err;
#define GUARD(expr, fail)                      \
    try { expr; }                              \
    catch (const std::exception& e) { g_err = e.what(); return fail; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_create(const char* chars, int n, const int* offsets, const unsigned char* mask, int nulls)
{
    GUARD(return NVStrings::create_from_offsets(chars, n, offsets, mask, nulls, false), nullptr);
}

void* ref_create_from_array(const char** strs, unsigned n)
{
    GUARD(return NVStrings::create_from_array(strs, n), nullptr);
}

void ref_destroy(void* h) { if (h) NVStrings::destroy((NVStrings*)h); }

unsigned ref_size(void* h) { return ((NVStrings*)h)->size(); }

// total bytes of all strings (nulls count 0)
long ref_total_bytes(void* h)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<int> lens(s->size());
    s->byte_count(lens.data(), false);
    long t = 0;
    for (int v : lens) t += v > 0 ? v : 0;
    return t;
}

// export to (chars, offsets[n+1], mask[(n+7)/8]); returns null count
int ref_export(void* h, char* chars, int* offsets, unsigned char* mask)
{
    NVStrings* s = (NVStrings*)h;
    GUARD(return s->create_offsets(chars, offsets, mask, false), -100);
}

int ref_len(void* h, int* out) { GUARD(return (int)((NVStrings*)h)->len(out, false), -100); }
int ref_hash(void* h, unsigned* out) { GUARD(return ((NVStrings*)h)->hash(out, false), -100); }

int ref_contains_re(void* h, const char* pat, bool* out) { GUARD(return ((NVStrings*)h)->contains_re(pat, out, false), -100); }
int ref_match(void* h, const char* pat, bool* out) { GUARD(return ((NVStrings*)h)->match(pat, out, false), -100); }
int ref_count_re(void* h, const char* pat, int* out) { GUARD(return ((NVStrings*)h)->count_re(pat, out, false), -100); }

void* ref_replace_re(void* h, const char* pat, const char* repl, int maxrepl)
{
    GUARD(return ((NVStrings*)h)->replace_re(pat, repl, maxrepl), nullptr);
}
void* ref_replace_re_multi(void* h, const char** pats, int npats, void* repls)
{
    std::vector<const char*> v(pats, pats + npats);
    GUARD(return ((NVStrings*)h)->replace_re(v, *(NVStrings*)repls), nullptr);
}
void* ref_replace_with_backrefs(void* h, const char* pat, const char* repl)
{
    GUARD(return ((NVStrings*)h)->replace_with_backrefs(pat, repl), nullptr);
}
void* ref_replace(void* h, const char* str, const char* repl, int maxrepl)
{
    GUARD(return ((NVStrings*)h)->replace(str, repl, maxrepl), nullptr);
}
void* ref_replace_multi(void* h, void* tgts, void* repls)
{
    GUARD(return ((NVStrings*)h)->replace(*(NVStrings*)tgts, *(NVStrings*)repls), nullptr);
}

int ref_find(void* h, const char* str, int start, int end, int* out) { GUARD(return (int)((NVStrings*)h)->find(str, start, end, out, false), -100); }
int ref_find_from(void* h, const char* str, int* starts, int* ends, int* out) { GUARD(return (int)((NVStrings*)h)->find_from(str, starts, ends, out, false), -100); }
int ref_match_strings(void* h, void* other, bool* out) { GUARD(return ((NVStrings*)h)->match_strings(*(NVStrings*)other, out, false), -100); }
int ref_rfind(void* h, const char* str, int start, int end, int* out) { GUARD(return (int)((NVStrings*)h)->rfind(str, start, end, out, false), -100); }
int ref_contains(void* h, const char* str, bool* out) { GUARD(return ((NVStrings*)h)->contains(str, out, false), -100); }
int ref_startswith(void* h, const char* str, bool* out) { GUARD(return (int)((NVStrings*)h)->startswith(str, out, false), -100); }
int ref_endswith(void* h, const char* str, bool* out) { GUARD(return (int)((NVStrings*)h)->endswith(str, out, false), -100); }
int ref_find_multiple(void* h, void* strs, int* out) { GUARD(return (int)((NVStrings*)h)->find_multiple(*(NVStrings*)strs, out, false), -100); }

// column-major split; delimiter NULL => whitespace variant. Writes up to cap handles; returns #columns.
int ref_split(void* h, const char* delim, int maxsplit, int right, void** out, int cap)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<NVStrings*> res;
    try {
        if (right) { if (delim) s->rsplit(delim, maxsplit, res); else s->rsplit(maxsplit, res); }
        else       { if (delim) s->split(delim, maxsplit, res);  else s->split(maxsplit, res); }
    } catch (const std::exception& e) { g_err = e.what(); return -100; }
    int n = (int)res.size();
    for (int i = 0; i < n; ++i) { if (i < cap) out[i] = res[i]; else NVStrings::destroy(res[i]); }
    return n;
}

// row-major split; out must hold size() handles (NULL for null rows); returns total tokens
int ref_split_record(void* h, const char* delim, int maxsplit, int right, void** out)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<NVStrings*> res;
    int rc;
    try {
        if (right) rc = delim ? s->rsplit_record(delim, maxsplit, res) : s->rsplit_record(maxsplit, res);
        else       rc = delim ? s->split_record(delim, maxsplit, res)  : s->split_record(maxsplit, res);
    } catch (const std::exception& e) { g_err = e.what(); return -100; }
    for (size_t i = 0; i < res.size(); ++i) out[i] = res[i];
    return rc;
}

int ref_partition(void* h, const char* delim, int right, void** out)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<NVStrings*> res;
    int rc;
    GUARD(rc = right ? s->rpartition(delim, res) : s->partition(delim, res), -100);
    for (size_t i = 0; i < res.size(); ++i) out[i] = res[i];
    return rc;
}

// column-major regex results: findall (kind 0) / extract (kind 1); writes up to cap handles, returns #columns
int ref_regex_columns(void* h, const char* pat, int kind, void** out, int cap)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<NVStrings*> res;
    int rc;
    GUARD(rc = kind == 0 ? s->findall(pat, res) : s->extract(pat, res), -100);
    int n = (int)res.size();
    for (int i = 0; i < n; ++i) { if (i < cap) out[i] = res[i]; else NVStrings::destroy(res[i]); }
    (void)rc;
    return n;
}
// row-major: findall_record (kind 0) / extract_record (kind 1); out holds size() handles
int ref_regex_records(void* h, const char* pat, int kind, void** out)
{
    NVStrings* s = (NVStrings*)h;
    std::vector<NVStrings*> res;
    int rc;
    GUARD(rc = kind == 0 ? s->findall_record(pat, res) : s->extract_record(pat, res), -100);
    for (size_t i = 0; i < res.size(); ++i) out[i] = res[i];
    return rc;
}

// Dump of the reference's compiled program for `pattern` as int32 words:
//   [ninsts, start_inst, ngroups, nstarts, nclasses, then ninsts x (type, u1, u2), then nstarts start ids,
//    then per class: builtins, count, count x char]          (regcomp.h:51-103)
// Returns the number of words written (or needed if cap is too small).
int ref_regex_dump(const char* pattern, int* out, int cap)
{
    const char32_t* p32 = to_char32(pattern);
    Reprog* prog = Reprog::create_from(p32);
    delete p32;
    std::vector<int> w;
    w.push_back(prog->inst_count());
    w.push_back(prog->get_start_inst());
    w.push_back(prog->groups_count());
    int nstarts = prog->starts_count() - 1;  // last entry is the -1 terminator
    w.push_back(nstarts);
    w.push_back(prog->classes_count());
    for (int i = 0; i < prog->inst_count(); ++i) {
        Reinst& in = prog->inst_at(i);
        w.push_back(in.type);
        w.push_back(in.u1.right_id);
        w.push_back(in.u2.left_id);
    }
    for (int i = 0; i < nstarts; ++i) w.push_back(prog->starts_data()[i]);
    for (int k = 0; k < prog->classes_count(); ++k) {
        Reclass& c = prog->class_at(k);
        w.push_back(c.builtins);
        w.push_back((int)c.chrs.size());
        for (char32_t ch : c.chrs) w.push_back((int)ch);
    }
    delete prog;
    for (size_t i = 0; i < w.size() && (int)i < cap; ++i) out[i] = w[i];
    return (int)w.size();
}

void* ref_tokenize(void* h, const char* delim) { GUARD(return NVText::tokenize(*(NVStrings*)h, delim), nullptr); }
void* ref_tokenize_multi(void* h, void* delims) { GUARD(return NVText::tokenize(*(NVStrings*)h, *(NVStrings*)delims), nullptr); }
int ref_token_count(void* h, const char* delim, unsigned* out) { GUARD(return (int)NVText::token_count(*(NVStrings*)h, delim, out, false), -100); }

void* ref_cat_create(void* h) { GUARD(return NVCategory::create_from_strings(*(NVStrings*)h), nullptr); }
void* ref_cat_create_multi(void** hs, int n)
{
    std::vector<NVStrings*> v;
    for (int i = 0; i < n; ++i) v.push_back((NVStrings*)hs[i]);
    GUARD(return NVCategory::create_from_strings(v), nullptr);
}
void ref_cat_destroy(void* c) { if (c) NVCategory::destroy((NVCategory*)c); }
unsigned ref_cat_size(void* c) { return ((NVCategory*)c)->size(); }
unsigned ref_cat_keys_size(void* c) { return ((NVCategory*)c)->keys_size(); }
void* ref_cat_keys(void* c) { GUARD(return ((NVCategory*)c)->get_keys(), nullptr); }
int ref_cat_values(void* c, int* out) { GUARD(return ((NVCategory*)c)->get_values(out, false), -100); }
// kind 0: merge_category, 1: merge_and_remap
void* ref_cat_merge(void* c1, void* c2, int kind)
{
    GUARD(return kind ? ((NVCategory*)c1)->merge_and_remap(*(NVCategory*)c2) : ((NVCategory*)c1)->merge_category(*(NVCategory*)c2), nullptr);
}
void* ref_cat_from_categories(void** cs, int n)
{
    std::vector<NVCategory*> v;
    for (int i = 0; i < n; ++i) v.push_back((NVCategory*)cs[i]);
    GUARD(return NVCategory::create_from_categories(v), nullptr);
}
void* ref_cat_to_strings(void* c) { GUARD(return ((NVCategory*)c)->to_strings(), nullptr); }
// cheap attributes / transforms (strings/attrs.cu, case.cu, strip.cu, substr.cu)
int ref_is_class(void* h, int kind, bool* out)
{
    NVStrings* s = (NVStrings*)h;
    GUARD(return (int)(kind == 0 ? s->isalnum(out, false) : kind == 1 ? s->isalpha(out, false) : kind == 2 ? s->isdigit(out, false)
                 : kind == 3 ? s->isspace(out, false) : kind == 4 ? s->isdecimal(out, false) : kind == 5 ? s->isnumeric(out, false)
                 : kind == 6 ? s->islower(out, false) : kind == 7 ? s->isupper(out, false) : s->is_empty(out, false)), -100);
}
void* ref_case(void* h, int upper) { GUARD(return upper ? ((NVStrings*)h)->upper() : ((NVStrings*)h)->lower(), nullptr); }
void* ref_strip(void* h, const char* chars, int side)
{
    NVStrings* s = (NVStrings*)h;
    GUARD(return side == 0 ? s->strip(chars) : side == 1 ? s->lstrip(chars) : s->rstrip(chars), nullptr);
}
void* ref_slice(void* h, int start, int stop, int step) { GUARD(return ((NVStrings*)h)->slice(start, stop, step), nullptr); }
// key-set algebra / gathers (NVCategory.cu:1084-1220,1375-1820): op 0 add_keys, 1 remove_keys, 2 set_keys, 3 remove_unused
void* ref_cat_keys_op(void* c, void* strs, int op)
{
    NVCategory* cat = (NVCategory*)c;
    GUARD(return op == 0 ? cat->add_keys_and_remap(*(NVStrings*)strs) : op == 1 ? cat->remove_keys_and_remap(*(NVStrings*)strs)
                 : op == 2 ? cat->set_keys_and_remap(*(NVStrings*)strs) : cat->remove_unused_keys_and_remap(), nullptr);
}
void* ref_cat_gather(void* c, const int* pos, unsigned n, int remap)
{
    NVCategory* cat = (NVCategory*)c;
    GUARD(return remap ? cat->gather_and_remap(pos, n, false) : cat->gather(pos, n, false), nullptr);
}
void* ref_cat_gather_strings(void* c, const int* pos, unsigned n) { GUARD(return ((NVCategory*)c)->gather_strings(pos, n, false), nullptr); }

}  // extern "C"
