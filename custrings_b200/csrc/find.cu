// Literal find / contains / startswith / endswith / find_multiple and literal replace (single and multi target).
// Replaces NVStrings::find find.cu:75-120, rfind :163-199, find_multiple :202-233, contains :237-272,
// startswith/endswith :316-387, replace modify.cu:109-192 and replace(NVStrings&,NVStrings&) :263-299.
// Per-row logic lives in rowops.cuh (shared with the host simulation); kernels here are thread-per-row with the
// needle staged in shared memory.
#include "common.cuh"
#include "rowops.cuh"
#include "regex_bits.h"

namespace custr {

constexpr int FIND_THREADS = 256;
constexpr int NEEDLE_SMEM = 1024;

__device__ __forceinline__ const uint8_t* stage_needle(const uint8_t* g, int m, uint8_t* smem)
{
    if (m > NEEDLE_SMEM) return g;
    for (int i = threadIdx.x; i < m; i += blockDim.x) smem[i] = g[i];
    __syncthreads();
    return smem;
}

enum FindMode { FM_FIND = 0, FM_RFIND = 1, FM_CONTAINS = 2, FM_STARTS = 3, FM_ENDS = 4 };

__global__ void __launch_bounds__(FIND_THREADS)
k_find(ColView col, const uint8_t* __restrict__ needle, int m, int start, int end, int mode, int32_t* __restrict__ out_i,
       uint8_t* __restrict__ out_b, unsigned long long* __restrict__ total, const uint8_t* __restrict__ hits = nullptr)
{
    __shared__ uint8_t sm[NEEDLE_SMEM];
    const uint8_t* t = stage_needle(needle, m, sm);
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        int counted = 0;
        if (i < col.n) {
            bool ok = col.valid(i);
            const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
            int n = col.offsets[i + 1] - col.offsets[i];
            if (mode <= FM_RFIND) {
                // `hits` (find / rfind on a large column): rows the chain kernel found the needle nowhere in need no scan
                int r = !ok ? -2 : (hits && !hits[i]) ? -1 : row::find_chars(s, n, t, m, start, end, mode == FM_RFIND);
                out_i[i] = r;
                counted = r != -1;
            } else {
                bool r = false;
                if (ok) {
                    if (mode == FM_CONTAINS) r = m > 0 && row::find_bytes(s, 0, n, t, m) >= 0;
                    else if (mode == FM_STARTS) r = m <= n && row::bytes_equal(s, t, m);
                    else r = m <= n && row::bytes_equal(s + n - m, t, m);
                }
                out_b[i] = r;
                counted = r;
            }
        }
        unsigned b = __ballot_sync(0xffffffffu, counted);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(total, (unsigned long long)__popc(b));
    }
}

__global__ void __launch_bounds__(FIND_THREADS)
k_find_multiple(ColView col, ColView tg, int32_t* __restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        bool ok = col.valid(i);
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        int n = col.offsets[i + 1] - col.offsets[i];
        for (int t = 0; t < tg.n; ++t) {
            int r = -2;
            if (ok && tg.valid(t)) {
                int tb = tg.offsets[t];
                r = row::find_chars(s, n, (const uint8_t*)tg.chars + tb, tg.offsets[t + 1] - tb, 0, -1, false);
            }
            out[(size_t)i * tg.n + t] = r;
        }
    }
}

__global__ void k_count_not_minus1(const int32_t* __restrict__ v, int n, unsigned long long* total)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned b = __ballot_sync(0xffffffffu, i < n && v[i] != -1);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(total, (unsigned long long)__popc(b));
}

__global__ void __launch_bounds__(FIND_THREADS)
k_replace_literal(ColView col, const uint8_t* __restrict__ tgt, int m, const uint8_t* __restrict__ repl, int rl, int maxrepl,
                  int32_t* __restrict__ out_len, const int32_t* __restrict__ out_off, char* __restrict__ out_chars)
{
    __shared__ uint8_t sm[NEEDLE_SMEM];
    const uint8_t* t = stage_needle(tgt, m, sm);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        int n = col.offsets[i + 1] - col.offsets[i];
        int total = row::replace_literal(s, n, t, m, repl, rl, maxrepl, out_chars ? out_chars + out_off[i] : nullptr);
        if (!out_chars) out_len[i] = total;
    }
}

__global__ void __launch_bounds__(FIND_THREADS)
k_replace_literal_multi(ColView col, ColView tg, ColView rp, int32_t* __restrict__ out_len, const int32_t* __restrict__ out_off,
                        char* __restrict__ out_chars)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        if (!col.valid(i)) { if (!out_chars) out_len[i] = 0; continue; }
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        int n = col.offsets[i + 1] - col.offsets[i];
        int total = row::replace_literal_multi(s, n, tg, rp, out_chars ? out_chars + out_off[i] : nullptr);
        if (!out_chars) out_len[i] = total;
    }
}

static inline int row_grid(int n, int threads)
{
    int want = (n + threads - 1) / threads;
    int cap = num_sms() * 32;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

static unsigned long long fetch(unsigned long long* d)
{
    unsigned long long h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h;
}

// regex.cu: literal `contains` through the bit-stream chain kernel (rows holding a NUL byte are left to the caller)
bool literal_contains_chain(const custr_column* col, const char* literal, uint8_t* out_dev, unsigned long long* total, int32_t** dirty_rows,
                            unsigned int** dirty_count, BufPtr& keep_rows, BufPtr& keep_count);

// byte-compare `contains` for the listed rows only
__global__ void __launch_bounds__(FIND_THREADS)
k_contains_rows(ColView col, const uint8_t* __restrict__ needle, int m, const int32_t* __restrict__ rows, const unsigned int* __restrict__ nrows_ptr,
                uint8_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    const int nrows = (int)*nrows_ptr;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nrows; k += gridDim.x * blockDim.x) {
        const int i = rows[k];
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const bool r = m > 0 && row::find_bytes(s, 0, col.offsets[i + 1] - col.offsets[i], needle, m) >= 0;
        out[i] = r;
        if (r) atomicAdd(total, 1ull);
    }
}

static int find_family(const custr_column* col, const char* str, int start, int end, int mode, int32_t* out_i, uint8_t* out_b,
                       int devmem, int null_rc)
{
    if (!col || !str || (!out_i && !out_b)) return null_rc;
    int32_t n = col->n;
    if (n == 0) return 0;
    if (start < 0) start = 0;
    int m = (int)strlen(str);
    BufPtr d_needle = upload(str, m + 1);
    Scratch<unsigned long long> total(1);
    CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
    unsigned long long cnt;
    if (out_i) {
        ResultBuf<int32_t> out(out_i, n, devmem);
        // large column: one coalesced pass of the chain kernel decides which rows hold the needle at all (most do not, and a
        // row without it is -1 for any start / end); only those rows are scanned for the character position
        BufPtr hits;
        if (m > 0 && col->nbytes >= (1 << 20)) {
            hits = dev_alloc((size_t)n);
            Scratch<unsigned long long> nhit(1);
            CUSTR_CUDA(cudaMemsetAsync(nhit.get(), 0, 8, g_stream));
            int32_t* dirty_rows = nullptr;
            unsigned int* dirty_count = nullptr;
            BufPtr keep_rows, keep_count;
            if (literal_contains_chain(col, str, (uint8_t*)hits->ptr, nhit.get(), &dirty_rows, &dirty_count, keep_rows, keep_count))
                LAUNCH(k_contains_rows, 64, FIND_THREADS, 0, view_of(col), (const uint8_t*)d_needle->ptr, m, (const int32_t*)dirty_rows,
                       (const unsigned int*)dirty_count, (uint8_t*)hits->ptr, nhit.get());
            else
                hits = nullptr;
        }
        LAUNCH(k_find, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), (const uint8_t*)d_needle->ptr, m, start, end, mode,
               out.dev, (uint8_t*)nullptr, total.get(), hits ? (const uint8_t*)hits->ptr : (const uint8_t*)nullptr);
        cnt = fetch(total.get());
        out.finish();
    } else {
        ResultBuf<uint8_t> out(out_b, n, devmem);
        if (mode == FM_CONTAINS && m > 0 && col->nbytes >= (1 << 20)) {  // large column: the chain kernel reads every byte once, coalesced
            int32_t* dirty_rows = nullptr;
            unsigned int* dirty_count = nullptr;
            BufPtr keep_rows, keep_count;
            if (literal_contains_chain(col, str, out.dev, total.get(), &dirty_rows, &dirty_count, keep_rows, keep_count)) {
                LAUNCH(k_contains_rows, 64, FIND_THREADS, 0, view_of(col), (const uint8_t*)d_needle->ptr, m, (const int32_t*)dirty_rows,
                       (const unsigned int*)dirty_count, out.dev, total.get());
                cnt = fetch(total.get());
                out.finish();
                return (int)cnt;
            }
        }
        LAUNCH(k_find, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), (const uint8_t*)d_needle->ptr, m, start, end, mode,
               (int32_t*)nullptr, out.dev, total.get());
        cnt = fetch(total.get());
        out.finish();
    }
    return (int)cnt;
}

// find_from (find.cu:123-160): per-row start / end character positions
__global__ void __launch_bounds__(FIND_THREADS)
k_find_from(ColView col, const uint8_t* __restrict__ needle, int m, const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
            int32_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    __shared__ uint8_t sm[NEEDLE_SMEM];
    const uint8_t* t = stage_needle(needle, m, sm);
    for (int base = blockIdx.x * blockDim.x; base < col.n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        int counted = 0;
        if (i < col.n) {
            int r = -2;
            if (col.valid(i)) {
                const int pos = starts ? starts[i] : 0;
                const int end = ends ? ends[i] : pos - 1;  // count = end - pos < 0: to the end of the string
                r = row::find_chars((const uint8_t*)col.chars + col.offsets[i], col.offsets[i + 1] - col.offsets[i], t, m, pos, end, false);
            }
            out[i] = r;
            counted = r != -1;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, counted);
        if ((threadIdx.x & 31) == 0 && mk) atomicAdd(total, (unsigned long long)__popc(mk));
    }
}

// match_strings (find.cu:276-313): row-wise equality of two columns; two nulls are equal, a null and a string are not
__global__ void __launch_bounds__(FIND_THREADS)
k_match_strings(ColView a, ColView b, uint8_t* __restrict__ out, unsigned long long* __restrict__ total)
{
    for (int base = blockIdx.x * blockDim.x; base < a.n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        bool r = false;
        if (i < a.n) {
            const bool va = a.valid(i), vb = b.valid(i);
            if (va && vb) {
                const int na = a.offsets[i + 1] - a.offsets[i], nb = b.offsets[i + 1] - b.offsets[i];
                r = na == nb && row::bytes_equal((const uint8_t*)a.chars + a.offsets[i], (const uint8_t*)b.chars + b.offsets[i], na);
            } else
                r = va == vb;
            out[i] = r;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, r);
        if ((threadIdx.x & 31) == 0 && mk) atomicAdd(total, (unsigned long long)__popc(mk));
    }
}

static BufPtr validity_copy(const custr_column* col)
{
    if (!col->validity) return nullptr;
    BufPtr v = dev_alloc((col->n + 7) / 8);
    custr_create_offsets(col, nullptr, nullptr, (uint8_t*)v->ptr, 1);
    return v;
}

}  // namespace custr

using namespace custr;

extern "C" {

int custr_find(const custr_column* col, const char* str, int32_t start, int32_t end, int32_t* results, int devmem)
{
    return guarded([&] { return find_family(col, str, start, end, FM_FIND, results, nullptr, devmem, 0); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}
int custr_rfind(const custr_column* col, const char* str, int32_t start, int32_t end, int32_t* results, int devmem)
{
    return guarded([&] { return find_family(col, str, start, end, FM_RFIND, results, nullptr, devmem, 0); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}
int custr_find_from(const custr_column* col, const char* str, const int32_t* starts, const int32_t* ends, int32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !str || !results) return 0;  // find.cu:126-127
            const int32_t n = col->n;
            if (n == 0) return 0;
            const int m = (int)strlen(str);
            BufPtr d_needle = upload(str, m + 1);
            BufPtr ds, de;  // host arrays are staged when devmem == 0
            if (!devmem && starts) { ds = upload(starts, sizeof(int32_t) * (size_t)n); starts = (const int32_t*)ds->ptr; }
            if (!devmem && ends) { de = upload(ends, sizeof(int32_t) * (size_t)n); ends = (const int32_t*)de->ptr; }
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            ResultBuf<int32_t> out(results, n, devmem);
            LAUNCH(k_find_from, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), (const uint8_t*)d_needle->ptr, m, starts, ends, out.dev,
                   total.get());
            const long long cnt = fetch(total.get());
            out.finish();
            return (int)cnt;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_match_strings(const custr_column* col, const custr_column* other, uint8_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!results) return -1;
            if (!col || !other) return fail(CUSTR_ERR_ARG, "match_strings: null column");
            const int32_t n = col->n;
            if (n == 0) return 0;
            if (other->n != n) return fail(CUSTR_ERR_INVALID, "sizes must match");
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            ResultBuf<uint8_t> out(results, n, devmem);
            LAUNCH(k_match_strings, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), view_of(other), out.dev, total.get());
            const long long cnt = fetch(total.get());
            out.finish();
            return (int)cnt;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_contains(const custr_column* col, const char* str, uint8_t* results, int devmem)
{
    return guarded([&] { return find_family(col, str, 0, -1, FM_CONTAINS, nullptr, results, devmem, -1); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}
int custr_startswith(const custr_column* col, const char* str, uint8_t* results, int devmem)
{
    return guarded([&] { return find_family(col, str, 0, -1, FM_STARTS, nullptr, results, devmem, 0); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}
int custr_endswith(const custr_column* col, const char* str, uint8_t* results, int devmem)
{
    return guarded([&] { return find_family(col, str, 0, -1, FM_ENDS, nullptr, results, devmem, 0); }, (int)CUSTR_ERR_ARG,
                   (int)CUSTR_ERR_CUDA);
}

int custr_find_multiple(const custr_column* col, const custr_column* targets, int32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !targets || !results) return 0;
            int32_t n = col->n, m = targets->n;
            if (n == 0 || m == 0) return 0;
            ResultBuf<int32_t> out(results, (size_t)n * m, devmem);
            LAUNCH(k_find_multiple, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), view_of(targets), out.dev);
            // the reference counts only the first `n` entries of the n*m result (find.cu:226) — kept
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            LAUNCH(k_count_not_minus1, (n + 255) / 256, 256, 0, (const int32_t*)out.dev, n, total.get());
            unsigned long long cnt = fetch(total.get());
            out.finish();
            return (int)cnt;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_replace(const custr_column* col, const char* str, const char* repl, int32_t maxrepl)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "replace: null column")};
            if (!str || !*str) throw ArgError{fail(CUSTR_ERR_INVALID, "replace parameter cannot be null or empty")};
            if (!repl) repl = "";
            int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            int m = (int)strlen(str), rl = (int)strlen(repl);
            if (maxrepl < 0 && !bits::g_force_generic) {  // every occurrence: bit-stream splice (replace_bits.cuh), no per-row walk
                BufPtr chars, off;
                int64_t total = 0;
                if (bits::replace_literal_flat(col, str, m, repl, rl, chars, off, total))
                    return make_column(chars, off, validity_copy(col), n, col->nulls, total);
            }
            BufPtr d_t = upload(str, m + 1), d_r = upload(repl, rl + 1);
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            LAUNCH(k_replace_literal, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), (const uint8_t*)d_t->ptr, m,
                   (const uint8_t*)d_r->ptr, rl, maxrepl, lens.get(), (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
            if (total > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "replace: result exceeds 2 GiB of chars")};
            BufPtr chars = dev_alloc((size_t)total);
            LAUNCH(k_replace_literal, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), (const uint8_t*)d_t->ptr, m,
                   (const uint8_t*)d_r->ptr, rl, maxrepl, (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, validity_copy(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_replace_multi(const custr_column* col, const custr_column* targets, const custr_column* repls)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col || !targets || !repls) throw ArgError{fail(CUSTR_ERR_ARG, "replace: null column")};
            if (targets->n == 0 || repls->n == 0)
                throw ArgError{fail(CUSTR_ERR_INVALID, "replace targets and repls parameters cannot be empty")};
            if (repls->n > 1 && repls->n != targets->n)
                throw ArgError{fail(CUSTR_ERR_INVALID, "replace targets and replacement sizes must match")};
            int32_t n = col->n;
            if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
            Scratch<int32_t> lens((size_t)n + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            LAUNCH(k_replace_literal_multi, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), view_of(targets),
                   view_of(repls), lens.get(), (const int32_t*)nullptr, (char*)nullptr);
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
            if (total > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "replace: result exceeds 2 GiB of chars")};
            BufPtr chars = dev_alloc((size_t)total);
            LAUNCH(k_replace_literal_multi, row_grid(n, FIND_THREADS), FIND_THREADS, 0, view_of(col), view_of(targets),
                   view_of(repls), (int32_t*)nullptr, (const int32_t*)off->ptr, (char*)chars->ptr);
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return make_column(chars, off, validity_copy(col), n, col->nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

}  // extern "C"
