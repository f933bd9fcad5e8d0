// NVStrings / NVCategory / NVText: the reference's C++ class surface for the hot path, as a thin host layer over the
// C-ABI (include/custr.h).  Status codes are translated back into the reference's exception types
// (std::invalid_argument: replace.cu:112, modify.cu:111; std::runtime_error: util.h:47-55).
#include "../../include/NVStrings.h"
#include "../../include/NVCategory.h"
#include "../../include/NVText.h"
#include "../../include/custr.h"
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void raise(int code)
{
    std::string msg = custr_last_error();
    if (code == CUSTR_ERR_CUDA || code == CUSTR_ERR_ALLOC) throw std::runtime_error(msg);
    throw std::invalid_argument(msg);
}
int checked(int rc)
{
    if (rc <= CUSTR_ERR_INVALID) raise(rc);
    return rc;
}
custr_column* checked(custr_column* c)
{
    if (!c) {
        std::string msg = custr_last_error();
        if (msg.find("cuda") != std::string::npos || msg.find("allocation") != std::string::npos) throw std::runtime_error(msg);
        throw std::invalid_argument(msg);
    }
    return c;
}
}  // namespace

NVStrings::~NVStrings() { custr_column_free(col_); }
void NVStrings::destroy(NVStrings* inst) { delete inst; }

NVStrings* NVStrings::create_from_array(const char** strs, unsigned int count)
{
    return new NVStrings(checked(custr_create_from_array(strs, count)));
}
NVStrings* NVStrings::create_from_index(std::pair<const char*, size_t>* strs, unsigned int count, bool devmem, sorttype stype)
{
    static_assert(sizeof(std::pair<const char*, size_t>) == 16, "pair layout");
    // bad device pointers surface as std::invalid_argument("nvstrings::create_from_index bad_device_ptr"), NVStrings.cu:98-99
    return new NVStrings(checked(custr_create_from_index(strs, count, devmem, (int)stype)));
}
NVStrings* NVStrings::create_from_offsets(const char* strs, int count, const int* offsets, const unsigned char* nullbitmask, int nulls,
                                          bool devmem)
{
    return new NVStrings(checked(custr_create_from_offsets(strs, count, offsets, nullbitmask, nulls, devmem)));
}
unsigned int NVStrings::size() const { return custr_size(col_); }
int NVStrings::create_offsets(char* strs, int* offsets, unsigned char* nullbitmask, bool devmem)
{
    if (size() == 0 || !strs || !offsets) return 0;
    return checked(custr_create_offsets(col_, strs, offsets, nullbitmask, devmem));
}
unsigned int NVStrings::set_null_bitarray(unsigned char* bitarray, bool emptyIsNull, bool devmem)
{
    return (unsigned)checked(custr_set_null_bitarray(col_, bitarray, emptyIsNull, devmem));
}
int NVStrings::to_host(char** list, int start, int end)
{
    unsigned n = size();
    if (!list || n == 0) return 0;
    if (end < 0 || end > (int)n) end = (int)n;
    if (start < 0) start = 0;
    if (start >= end) return 0;
    std::vector<char> chars((size_t)custr_chars_bytes(col_) + 1);
    std::vector<int> off(n + 1);
    std::vector<unsigned char> val((n + 7) / 8);
    checked(custr_create_offsets(col_, chars.data(), off.data(), val.data(), 0));
    for (int i = start; i < end; ++i) {
        char* dst = list[i - start];
        if (!dst || !((val[i >> 3] >> (i & 7)) & 1)) continue;  // null rows are skipped like NVStrings.cu:305-312
        memcpy(dst, chars.data() + off[i], (size_t)(off[i + 1] - off[i]));
    }
    return 0;
}
unsigned int NVStrings::len(int* lengths, bool devmem) { return (unsigned)checked(custr_len(col_, lengths, devmem)); }
size_t NVStrings::byte_count(int* lengths, bool devmem)
{
    long long rc = custr_byte_count(col_, lengths, devmem);
    if (rc < 0) raise((int)rc);
    return (size_t)rc;
}
int NVStrings::hash(unsigned int* results, bool devmem)
{
    if (size() == 0 || !results) return -1;
    return checked(custr_hash(col_, results, devmem));
}

int NVStrings::contains_re(const char* pattern, bool* results, bool devmem)
{
    return checked(custr_contains_re(col_, pattern, (unsigned char*)results, devmem));
}
int NVStrings::match(const char* pattern, bool* results, bool devmem)
{
    return checked(custr_match(col_, pattern, (unsigned char*)results, devmem));
}
int NVStrings::count_re(const char* pattern, int* results, bool devmem) { return checked(custr_count_re(col_, pattern, results, devmem)); }
NVStrings* NVStrings::replace_re(const char* pattern, const char* repl, int maxrepl)
{
    return new NVStrings(checked(custr_replace_re(col_, pattern, repl, maxrepl)));
}
NVStrings* NVStrings::replace_with_backrefs(const char* pattern, const char* repl)
{
    return new NVStrings(checked(custr_replace_with_backrefs(col_, pattern, repl)));
}
NVStrings* NVStrings::replace_re(std::vector<const char*>& patterns, NVStrings& repls)
{
    return new NVStrings(checked(custr_replace_re_multi(col_, patterns.data(), (int)patterns.size(), repls.col_)));
}

static int collect_columns(int (*fn)(const custr_column*, const char*, custr_column**, int32_t), const custr_column* col,
                           const char* pattern, std::vector<NVStrings*>& results, NVStrings* (*wrap)(custr_column*))
{
    int cap = 64;
    for (;;) {
        std::vector<custr_column*> out((size_t)cap, nullptr);
        int cols = checked(fn(col, pattern, out.data(), cap));
        if (cols <= cap) {
            for (int c = 0; c < cols; ++c) results.push_back(wrap(out[c]));
            return cols;
        }
        for (int c = 0; c < cap; ++c) custr_column_free(out[c]);
        cap = cols;
    }
}
int NVStrings::findall(const char* pattern, std::vector<NVStrings*>& results)
{
    if (!pattern) return -1;
    return collect_columns(custr_findall, col_, pattern, results, [](custr_column* c) { return new NVStrings(c); });
}
int NVStrings::extract(const char* pattern, std::vector<NVStrings*>& results)
{
    if (!pattern) return -1;
    return collect_columns(custr_extract, col_, pattern, results, [](custr_column* c) { return new NVStrings(c); });
}
int NVStrings::findall_record(const char* pattern, std::vector<NVStrings*>& results)
{
    if (!pattern) return -1;
    unsigned n = size();
    custr_column* tokens = nullptr;
    std::vector<int> row_off(n + 1, 0);
    int total = checked(custr_findall_record(col_, pattern, &tokens, row_off.data(), 0));
    checked(tokens);
    for (unsigned i = 0; i < n; ++i) results.push_back(new NVStrings(checked(custr_slice_rows(tokens, row_off[i], row_off[i + 1]))));
    custr_column_free(tokens);
    return total;
}

unsigned int NVStrings::find(const char* str, int start, int end, int* results, bool devmem)
{
    return (unsigned)checked(custr_find(col_, str, start, end, results, devmem));
}
unsigned int NVStrings::rfind(const char* str, int start, int end, int* results, bool devmem)
{
    return (unsigned)checked(custr_rfind(col_, str, start, end, results, devmem));
}
unsigned int NVStrings::find_from(const char* str, int* starts, int* ends, int* results, bool devmem)
{
    return (unsigned)checked(custr_find_from(col_, str, starts, ends, results, devmem));
}
int NVStrings::match_strings(NVStrings& strs, bool* results, bool devmem)
{
    return checked(custr_match_strings(col_, strs.col_, (uint8_t*)results, devmem));
}
unsigned int NVStrings::find_multiple(NVStrings& strs, int* results, bool devmem)
{
    return (unsigned)checked(custr_find_multiple(col_, strs.col_, results, devmem));
}
int NVStrings::contains(const char* str, bool* results, bool devmem)
{
    return checked(custr_contains(col_, str, (unsigned char*)results, devmem));
}
unsigned int NVStrings::startswith(const char* str, bool* results, bool devmem)
{
    return (unsigned)checked(custr_startswith(col_, str, (unsigned char*)results, devmem));
}
unsigned int NVStrings::endswith(const char* str, bool* results, bool devmem)
{
    return (unsigned)checked(custr_endswith(col_, str, (unsigned char*)results, devmem));
}
NVStrings* NVStrings::replace(const char* str, const char* repl, int maxrepl)
{
    return new NVStrings(checked(custr_replace(col_, str, repl, maxrepl)));
}
NVStrings* NVStrings::replace(NVStrings& strs, NVStrings& repls)
{
    return new NVStrings(checked(custr_replace_multi(col_, strs.col_, repls.col_)));
}

typedef int (*RecordFn)(const custr_column*, const char*, int32_t, custr_column**, int32_t*, int);
typedef int (*ColumnsFn)(const custr_column*, const char*, int32_t, custr_column**, int32_t);

int NVStrings::split_record(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results)
{
    return record_split_(delimiter, maxsplit, results, false);
}
int NVStrings::rsplit_record(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results)
{
    return record_split_(delimiter, maxsplit, results, true);
}
int NVStrings::rsplit_record(int maxsplit, std::vector<NVStrings*>& results) { return rsplit_record(nullptr, maxsplit, results); }
int NVStrings::record_split_(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results, bool right)
{
    unsigned n = size();
    custr_column* tokens = nullptr;
    std::vector<int> row_off(n + 1, 0);
    RecordFn fn = right ? custr_rsplit_record : custr_split_record;
    int total = checked(fn(col_, delimiter, maxsplit, &tokens, row_off.data(), 0));
    checked(tokens);
    std::vector<unsigned char> val((n + 7) / 8 + 1);
    if (n) custr_set_null_bitarray(col_, val.data(), 0, 0);
    for (unsigned i = 0; i < n; ++i) {
        if (!((val[i >> 3] >> (i & 7)) & 1)) { results.push_back(nullptr); continue; }
        results.push_back(new NVStrings(checked(custr_slice_rows(tokens, row_off[i], row_off[i + 1]))));
    }
    custr_column_free(tokens);  // the per-row views keep the shared buffers alive
    return total;
}
int NVStrings::split_record(int maxsplit, std::vector<NVStrings*>& results) { return split_record(nullptr, maxsplit, results); }

unsigned int NVStrings::split(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results)
{
    return column_split_(delimiter, maxsplit, results, false);
}
unsigned int NVStrings::rsplit(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results)
{
    return column_split_(delimiter, maxsplit, results, true);
}
unsigned int NVStrings::rsplit(int maxsplit, std::vector<NVStrings*>& results) { return rsplit(nullptr, maxsplit, results); }
int NVStrings::partition_(const char* delimiter, std::vector<NVStrings*>& results, bool right)
{
    if (!delimiter || !*delimiter) return 0;  // split.cu:1167-1171
    custr_column* flat = checked(custr_partition(col_, delimiter, right ? 1 : 0));
    const unsigned n = size();
    for (unsigned i = 0; i < n; ++i) results.push_back(new NVStrings(checked(custr_slice_rows(flat, 3 * (int)i, 3 * (int)i + 3))));
    custr_column_free(flat);
    return (int)n;
}
int NVStrings::partition(const char* delimiter, std::vector<NVStrings*>& results) { return partition_(delimiter, results, false); }
int NVStrings::rpartition(const char* delimiter, std::vector<NVStrings*>& results) { return partition_(delimiter, results, true); }
unsigned int NVStrings::column_split_(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results, bool right)
{
    int cap = 64;
    ColumnsFn fn = right ? custr_rsplit : custr_split;
    for (;;) {
        std::vector<custr_column*> out((size_t)cap, nullptr);
        int cols = checked(fn(col_, delimiter, maxsplit, out.data(), cap));
        if (cols <= cap) {
            for (int c = 0; c < cols; ++c) results.push_back(new NVStrings(out[c]));
            return (unsigned)results.size();
        }
        for (int c = 0; c < cap; ++c) custr_column_free(out[c]);
        cap = cols;
    }
}
unsigned int NVStrings::split(int maxsplit, std::vector<NVStrings*>& results) { return split(nullptr, maxsplit, results); }

NVStrings* NVStrings::gather(const int* pos, unsigned int count, bool devmem)
{
    return new NVStrings(checked(custr_gather(col_, pos, (int)count, devmem)));
}
NVStrings* NVStrings::sublist(unsigned int start, unsigned int end, int step)
{
    if (step > 1) {
        std::vector<int> idx;
        for (unsigned i = start; i < end; i += (unsigned)step) idx.push_back((int)i);
        return gather(idx.data(), (unsigned)idx.size(), false);
    }
    if (end > size()) end = size();
    if (start > end) start = end;
    return new NVStrings(checked(custr_slice_rows(col_, (int)start, (int)end)));
}

#define CUSTR_IS(NAME, KIND) \
    unsigned int NVStrings::NAME(bool* results, bool todevice) { return (unsigned int)checked(custr_is_class(col_, KIND, (uint8_t*)results, todevice)); }
CUSTR_IS(isalnum, 0) CUSTR_IS(isalpha, 1) CUSTR_IS(isdigit, 2) CUSTR_IS(isspace, 3) CUSTR_IS(isdecimal, 4) CUSTR_IS(isnumeric, 5)
CUSTR_IS(islower, 6) CUSTR_IS(isupper, 7) CUSTR_IS(is_empty, 8)
#undef CUSTR_IS
NVStrings* NVStrings::lower() { return new NVStrings(checked(custr_case(col_, 0))); }
NVStrings* NVStrings::upper() { return new NVStrings(checked(custr_case(col_, 1))); }
NVStrings* NVStrings::strip(const char* to_strip) { return new NVStrings(checked(custr_strip(col_, to_strip, 0))); }
NVStrings* NVStrings::lstrip(const char* to_strip) { return new NVStrings(checked(custr_strip(col_, to_strip, 1))); }
NVStrings* NVStrings::rstrip(const char* to_strip) { return new NVStrings(checked(custr_strip(col_, to_strip, 2))); }
NVStrings* NVStrings::slice(int start, int stop, int step) { return new NVStrings(checked(custr_slice(col_, start, stop, step))); }
NVStrings* NVStrings::get(unsigned int pos) { return slice((int)pos, (int)pos + 1, 1); }

// ---------------------------------------------------------------------------------------------------------- NVCategory
NVCategory::~NVCategory() { custr_category_free(cat_); }
void NVCategory::destroy(NVCategory* inst) { delete inst; }

static custr_category* checked_cat(custr_category* c)
{
    if (!c) throw std::invalid_argument(custr_last_error());
    return c;
}
NVCategory* NVCategory::create_from_strings(NVStrings& strs)
{
    const custr_column* cols[1] = {strs.column()};
    return new NVCategory(checked_cat(custr_category_create(cols, 1)));
}
NVCategory* NVCategory::create_from_strings(std::vector<NVStrings*>& strs)
{
    std::vector<const custr_column*> cols;
    for (NVStrings* s : strs) cols.push_back(s ? s->column() : nullptr);
    return new NVCategory(checked_cat(custr_category_create(cols.data(), (int)cols.size())));
}
NVCategory* NVCategory::create_from_array(const char** strs, unsigned int count)
{
    NVStrings* s = NVStrings::create_from_array(strs, count);
    NVCategory* c = nullptr;
    try { c = create_from_strings(*s); } catch (...) { NVStrings::destroy(s); throw; }
    NVStrings::destroy(s);
    return c;
}
NVCategory* NVCategory::create_from_offsets(const char* strs, unsigned int count, const int* offsets, const unsigned char* nullbitmask,
                                            int nulls, bool devmem)
{
    NVStrings* s = NVStrings::create_from_offsets(strs, (int)count, offsets, nullbitmask, nulls, devmem);
    NVCategory* c = nullptr;
    try { c = create_from_strings(*s); } catch (...) { NVStrings::destroy(s); throw; }
    NVStrings::destroy(s);
    return c;
}
unsigned int NVCategory::size() { return custr_category_size(cat_); }
unsigned int NVCategory::keys_size() { return custr_category_keys_size(cat_); }
bool NVCategory::has_nulls()
{
    custr_column* k = custr_category_keys(cat_);
    bool r = k && custr_null_count(k) > 0;
    custr_column_free(k);
    return r;
}
NVStrings* NVCategory::get_keys() { return new NVStrings(checked(custr_category_keys(cat_))); }
int NVCategory::get_values(int* results, bool devmem) { return checked(custr_category_values(cat_, results, devmem)); }
const int* NVCategory::values_cptr() { return custr_category_values_cptr(cat_); }
NVStrings* NVCategory::to_strings()
{
    custr_column* k = checked(custr_category_keys(cat_));
    custr_column* r = custr_gather(k, custr_category_values_cptr(cat_), (int)custr_category_size(cat_), 1);
    custr_column_free(k);
    return new NVStrings(checked(r));
}

NVCategory* NVCategory::create_from_categories(std::vector<NVCategory*>& cats)
{
    std::vector<const custr_category*> v;
    for (NVCategory* c : cats) v.push_back(c->cat_);
    return new NVCategory(checked_cat(custr_category_merge(v.data(), (int)v.size(), 1)));
}
NVCategory* NVCategory::merge_category(NVCategory& cat)
{
    const custr_category* v[2] = {cat_, cat.cat_};
    return new NVCategory(checked_cat(custr_category_merge(v, 2, 0)));
}
NVCategory* NVCategory::merge_and_remap(NVCategory& cat)
{
    const custr_category* v[2] = {cat_, cat.cat_};
    return new NVCategory(checked_cat(custr_category_merge(v, 2, 1)));
}
NVCategory* NVCategory::add_keys_and_remap(NVStrings& strs) { return new NVCategory(checked_cat(custr_category_keys_op(cat_, strs.column(), 0))); }
NVCategory* NVCategory::remove_keys_and_remap(NVStrings& strs) { return new NVCategory(checked_cat(custr_category_keys_op(cat_, strs.column(), 1))); }
NVCategory* NVCategory::set_keys_and_remap(NVStrings& strs) { return new NVCategory(checked_cat(custr_category_keys_op(cat_, strs.column(), 2))); }
NVCategory* NVCategory::remove_unused_keys_and_remap() { return new NVCategory(checked_cat(custr_category_keys_op(cat_, nullptr, 3))); }
NVStrings* NVCategory::gather_strings(const int* pos, unsigned int elems, bool devmem)
{
    return new NVStrings(checked(custr_category_gather_strings(cat_, pos, (int)elems, devmem)));
}
NVCategory* NVCategory::gather_and_remap(const int* pos, unsigned int elems, bool devmem)
{
    return new NVCategory(checked_cat(custr_category_gather(cat_, pos, (int)elems, devmem, 1)));
}
NVCategory* NVCategory::gather(const int* pos, unsigned int elems, bool devmem)
{
    return new NVCategory(checked_cat(custr_category_gather(cat_, pos, (int)elems, devmem, 0)));
}

// -------------------------------------------------------------------------------------------------------------- NVText
NVStrings* NVText::tokenize(NVStrings& strs, const char* delimiter) { return new NVStrings(checked(custr_tokenize(strs.column(), delimiter))); }
unsigned int NVText::token_count(NVStrings& strs, const char* delimiter, unsigned int* results, bool devmem)
{
    checked(custr_token_count(strs.column(), delimiter, results, devmem));
    return 0;
}
