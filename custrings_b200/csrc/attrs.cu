// Cheap per-character attribute / transform ops of SURVEY.md §8f row 4, restated over the Arrow column:
//   NVStrings::isalnum/isalpha/isdigit/isspace/isdecimal/isnumeric/islower/isupper/is_empty  (strings/attrs.cu:115-445)
//   NVStrings::lower / upper                                                                   (strings/case.cu:30-190)
//   NVStrings::strip / lstrip / rstrip                                (strings/strip.cu:30-200, custring_view.inl:1398-1600)
//   NVStrings::slice(start, stop, step)                               (strings/substr.cu:39-83, custring_view.inl:801-866)
// Character classes come from the same 65536-entry flag table the regex classes use (unicode/is_flags.h:33-40); the case map is
// the reference's charcases table rebuilt by tools/gen_unicode_cases.py.  One thread per row (these ops are a small fraction of
// the hot path; the transforms are length pass -> scan -> write pass like every other column-producing op here).
#include "common.cuh"
#include "device_utils.cuh"
#include <mutex>
#include <map>

namespace custr {

static const uint16_t k_cases_host[65536] = {
#include "unicode_cases.inc"
};
static const uint16_t* device_cases()
{
    static std::mutex mu;
    static std::map<int, uint16_t*> tables;
    int dev = 0;
    CUSTR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = tables.find(dev);
    if (it != tables.end()) return it->second;
    uint16_t* d = nullptr;
    CUSTR_CUDA(cudaMalloc(&d, sizeof(k_cases_host)));
    CUSTR_CUDA(cudaMemcpy(d, k_cases_host, sizeof(k_cases_host), cudaMemcpyHostToDevice));
    tables[dev] = d;
    return d;
}

// code point -> packed UTF-8 char (reference util.inl:22-49) and its byte width
__device__ __forceinline__ int put_cp(uint32_t cp, char* o)
{
    if (cp < 0x80u) { if (o) o[0] = (char)cp; return 1; }
    if (cp < 0x800u) { if (o) { o[0] = (char)(0xC0 | (cp >> 6)); o[1] = (char)(0x80 | (cp & 0x3F)); } return 2; }
    if (o) { o[0] = (char)(0xE0 | (cp >> 12)); o[1] = (char)(0x80 | ((cp >> 6) & 0x3F)); o[2] = (char)(0x80 | (cp & 0x3F)); }
    return 3;
}

enum { IS_ALNUM = 0, IS_ALPHA, IS_DIGIT, IS_SPACE, IS_DECIMAL, IS_NUMERIC, IS_LOWER, IS_UPPER, IS_EMPTY };

__global__ void k_is_class(ColView col, int kind, const uint8_t* __restrict__ uflags, uint8_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool r = false;
    if (i < col.n) {
        const bool valid = col.valid(i);
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const int n = col.offsets[i + 1] - col.offsets[i];
        if (kind == IS_EMPTY) r = !valid || n == 0;  // a null row is empty (attrs.cu:425)
        else if (valid && n > 0) {
            r = true;
            for (int p = 0; r && p < n;) {
                int w;
                const uint32_t cp = packed_to_cp(utf8_packed(s + p, s + n, w));
                const uint32_t f = cp <= 0xFFFFu ? uflags[cp] : 0u;
                switch (kind) {
                case IS_ALNUM: r = (f & 15u) != 0; break;
                case IS_ALPHA: r = (f & 8u) != 0; break;
                case IS_DIGIT: r = (f & 4u) != 0; break;
                case IS_SPACE: r = (f & 16u) != 0; break;
                case IS_DECIMAL: r = (f & 1u) != 0; break;
                case IS_NUMERIC: r = (f & 2u) != 0; break;
                case IS_LOWER: r = !(f & 8u) || (f & 64u); break;   // non-alphabetic characters do not count (attrs.cu:360)
                default: r = !(f & 8u) || (f & 32u); break;
                }
                p += w;
            }
        }
        out[i] = r;
    }
    const unsigned m = __ballot_sync(0xffffffffu, r);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
}

// lower (to_upper == 0) / upper: lens != null -> length pass, else write pass
__global__ void k_case(ColView col, int to_upper, const uint8_t* __restrict__ uflags, const uint16_t* __restrict__ cases, int32_t* __restrict__ lens,
                       const int32_t* __restrict__ out_off, char* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= col.n) return;
    const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
    const int n = col.offsets[i + 1] - col.offsets[i];
    char* o = out ? out + out_off[i] : nullptr;
    int bytes = 0;
    const uint32_t want = to_upper ? 64u : 32u;  // characters flagged LOWER are mapped by upper(), UPPER ones by lower()
    for (int p = 0; p < n;) {
        int w;
        const uint32_t ch = utf8_packed(s + p, s + n, w);
        const uint32_t cp = packed_to_cp(ch);
        if (cp <= 0xFFFFu && (uflags[cp] & want)) bytes += put_cp(cases[cp], o ? o + bytes : nullptr);
        else {
            if (o)
                for (int k = 0; k < w; ++k) o[bytes + k] = (char)s[p + k];
            bytes += w;
        }
        p += w;
    }
    if (lens) lens[i] = bytes;
}

__device__ __forceinline__ bool one_of(const uint8_t* set, int nset, uint32_t ch)
{
    for (int q = 0; q < nset;) {
        int w;
        if (utf8_packed(set + q, set + nset, w) == ch) return true;
        q += w;
    }
    return false;
}
// side: 0 both, 1 left, 2 right.  Writes the kept byte range [b, e) of every row
__global__ void k_strip_ranges(ColView col, const uint8_t* __restrict__ set, int nset, int side, int32_t* __restrict__ lens, int32_t* __restrict__ begins)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= col.n) return;
    const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
    const int n = col.offsets[i + 1] - col.offsets[i];
    int b = 0, e = n;
    if (side != 2)
        while (b < n) {
            int w;
            if (!one_of(set, nset, utf8_packed(s + b, s + n, w))) break;
            b += w;
        }
    if (side != 1 && b < n)
        while (e > b) {
            int q = e - 1;
            while (q > b && (s[q] & 0xC0) == 0x80) --q;
            int w;
            if (!one_of(set, nset, utf8_packed(s + q, s + n, w))) break;
            e = q;
        }
    if (e < b) e = b;
    lens[i] = e - b;
    begins[i] = b;
}
// slice(start, stop, step) in characters (substr.cu:39-83): [start, stop or end), every step-th character
__global__ void k_slice(ColView col, int start, int stop, int step, int32_t* __restrict__ lens, const int32_t* __restrict__ out_off, char* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= col.n) return;
    const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
    const int n = col.offsets[i + 1] - col.offsets[i];
    char* o = out ? out + out_off[i] : nullptr;
    int bytes = 0;
    if (start >= 0) {
        int c = 0;  // character index
        for (int p = 0; p < n; ++c) {
            const int w = utf8_width(s[p]);
            if (stop > 0 && c >= stop) break;
            if (c >= start && (step <= 1 || (c - start) % step == 0)) {
                if (o)
                    for (int k = 0; k < w && p + k < n; ++k) o[bytes + k] = (char)s[p + k];
                bytes += (p + w <= n) ? w : n - p;
            }
            p += w;
        }
    }
    if (lens) lens[i] = bytes;
}
__global__ void k_copy_ranges(ColView col, const int32_t* __restrict__ begins, const int32_t* __restrict__ out_off, char* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= col.n) return;
    const char* s = col.chars + col.offsets[i] + begins[i];
    const int n = out_off[i + 1] - out_off[i];
    char* o = out + out_off[i];
    for (int k = 0; k < n; ++k) o[k] = s[k];
}

static BufPtr validity_copy(const custr_column* col)
{
    if (!col->validity) return nullptr;
    BufPtr v = dev_alloc((col->n + 7) / 8 + 1);
    custr_create_offsets(col, nullptr, nullptr, (uint8_t*)v->ptr, 1);
    return v;
}

}  // namespace custr
using namespace custr;

extern "C" {

// kind: 0 isalnum, 1 isalpha, 2 isdigit, 3 isspace, 4 isdecimal, 5 isnumeric, 6 islower, 7 isupper, 8 is_empty.
// results: bool per row (null rows false; is_empty: true); returns the number of true rows.
int custr_is_class(const custr_column* col, int kind, uint8_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !results || kind < 0 || kind > 8) return fail(CUSTR_ERR_ARG, "is_class: bad argument");
            if (col->n == 0) return 0;
            ResultBuf<uint8_t> out(results, (size_t)col->n, devmem);
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            LAUNCH(k_is_class, (col->n + 255) / 256, 256, 0, view_of(col), kind, device_unicode_flags(), out.dev, total.get());
            unsigned long long h = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h, total.get(), 8, cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            out.finish();
            return (int)h;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_case(const custr_column* col, int to_upper)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "lower/upper: null column")};
            const int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            LAUNCH(k_case, (n + 255) / 256, 256, 0, view_of(col), to_upper, device_unicode_flags(), device_cases(), lens.get(), (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            const int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
            BufPtr chars = dev_alloc((size_t)total);
            LAUNCH(k_case, (n + 255) / 256, 256, 0, view_of(col), to_upper, device_unicode_flags(), device_cases(), (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, validity_copy(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

// side: 0 strip, 1 lstrip, 2 rstrip; to_strip == NULL: " \n\t" (custring_view.inl:1403)
custr_column* custr_strip(const custr_column* col, const char* to_strip, int side)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col || side < 0 || side > 2) throw ArgError{fail(CUSTR_ERR_ARG, "strip: bad argument")};
            const int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            const char* set = to_strip ? to_strip : " \n\t";
            const int nset = (int)strlen(set);
            BufPtr d_set = upload(set, (size_t)nset + 1);
            Scratch<int32_t> lens((size_t)n + 1), begins((size_t)n);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            LAUNCH(k_strip_ranges, (n + 255) / 256, 256, 0, view_of(col), (const uint8_t*)d_set->ptr, nset, side, lens.get(), begins.get());
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            const int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
            BufPtr chars = dev_alloc((size_t)total);
            LAUNCH(k_copy_ranges, (n + 255) / 256, 256, 0, view_of(col), (const int32_t*)begins.get(), (const int32_t*)off->ptr, (char*)chars->ptr);
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, validity_copy(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_slice(const custr_column* col, int32_t start, int32_t stop, int32_t step)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "slice: null column")};
            if (stop > 0 && start > stop) throw ArgError{fail(CUSTR_ERR_INVALID, "nvstrings::slice start cannot be greater than stop")};
            const int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            LAUNCH(k_slice, (n + 255) / 256, 256, 0, view_of(col), start, stop, step, lens.get(), (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            const int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
            BufPtr chars = dev_alloc((size_t)total);
            LAUNCH(k_slice, (n + 255) / 256, 256, 0, view_of(col), start, stop, step, (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, validity_copy(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

}  // extern "C"
