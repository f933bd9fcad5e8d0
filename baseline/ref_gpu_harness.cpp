// C-ABI harness around the UNMODIFIED reference classes built for the GPU (baseline/Makefile): what bench.py's
// `reference_gpu` leg times on the B200 next to this repo's kernels.  All pointers are DEVICE pointers unless said
// otherwise.  Baseline infrastructure only — nothing under custrings_b200/ links, loads or calls it.
#include <cstdint>
#include <string>
#include <vector>
#include <stdexcept>
#include <cuda_runtime.h>
#include <NVStrings.h>
#include <NVCategory.h>
#include <NVText.h>

static thread_local std::string g_err;
#define GUARD(expr, fail)                      \
    try { expr; }                              \
    catch (const std::exception& e) { g_err = e.what(); return fail; }

extern "C" {
const char* refgpu_last_error() { return g_err.c_str(); }

// NVStrings::create_from_offsets(devmem=true): chars / offsets / mask in device memory (cpp/include/NVStrings.h:116)
void* refgpu_create(const char* d_chars, int n, const int* d_offsets, const unsigned char* d_mask, int nulls)
{
    GUARD(return NVStrings::create_from_offsets(d_chars, n, d_offsets, d_mask, nulls, true), nullptr);
}
void refgpu_destroy(void* h) { if (h) NVStrings::destroy((NVStrings*)h); }
unsigned refgpu_size(void* h) { return ((NVStrings*)h)->size(); }

int refgpu_contains_re(void* h, const char* pat, bool* d_out) { GUARD(return ((NVStrings*)h)->contains_re(pat, d_out, true), -100); }
int refgpu_count_re(void* h, const char* pat, int* d_out) { GUARD(return ((NVStrings*)h)->count_re(pat, d_out, true), -100); }
void* refgpu_replace_re(void* h, const char* pat, const char* repl, int maxrepl) { GUARD(return ((NVStrings*)h)->replace_re(pat, repl, maxrepl), nullptr); }
void* refgpu_replace(void* h, const char* tgt, const char* repl, int maxrepl) { GUARD(return ((NVStrings*)h)->replace(tgt, repl, maxrepl), nullptr); }
int refgpu_contains(void* h, const char* s, bool* d_out) { GUARD(return ((NVStrings*)h)->contains(s, d_out, true), -100); }

// split(delimiter): returns the number of columns, handles written to out[0..cap)
int refgpu_split(void* h, const char* delim, int maxsplit, void** out, int cap)
{
    std::vector<NVStrings*> res;
    GUARD(((NVStrings*)h)->split(delim, maxsplit, res), -100);
    for (size_t i = 0; i < res.size(); ++i) {
        if ((int)i < cap) out[i] = res[i];
        else NVStrings::destroy(res[i]);
    }
    return (int)res.size();
}
// split_record: rows' token columns are destroyed again (the timing includes the reference's N allocations)
long refgpu_split_record_total(void* h, const char* delim, int maxsplit)
{
    std::vector<NVStrings*> res;
    long total = 0;
    GUARD(total = ((NVStrings*)h)->split_record(delim, maxsplit, res), -100);
    for (NVStrings* r : res) if (r) NVStrings::destroy(r);
    return total;
}
void* refgpu_tokenize(void* h, const char* delim) { GUARD(return NVText::tokenize(*(NVStrings*)h, delim), nullptr); }

void* refgpu_category(void* h) { GUARD(return NVCategory::create_from_strings(*(NVStrings*)h), nullptr); }
void refgpu_category_destroy(void* c) { if (c) NVCategory::destroy((NVCategory*)c); }
unsigned refgpu_category_keys_size(void* c) { return ((NVCategory*)c)->keys_size(); }
int refgpu_category_values(void* c, int* d_out) { GUARD(return ((NVCategory*)c)->get_values(d_out, true), -100); }

// export to device (chars, offsets[n+1], mask): create_offsets(devmem=true), cpp/include/NVStrings.h
int refgpu_export(void* h, char* d_chars, int* d_offsets, unsigned char* d_mask) { GUARD(return ((NVStrings*)h)->create_offsets(d_chars, d_offsets, d_mask, true), -100); }
long refgpu_memsize(void* h) { GUARD(return (long)((NVStrings*)h)->memsize(), -100); }
}
