// split / split_record / tokenize / token_count.
// Replaces NVStrings::split split.cu:734-822 (delimiter) and :863-956 (whitespace), split_record :125-223 and
// :270-430, NVText::tokenize text/tokens.cu:123-155, token_count :337-361.
//
// Shape of the computation: the reference rescans every row once per output column and builds each
// column through create_from_index (3 more passes + an allocation), and allocates one object per ROW for
// split_record.  Here every variant is: count pass -> one scan -> length pass -> scan(s) -> one copy pass, and
// split_record / tokenize return ONE flat token column (+ row offsets) instead of N objects.
#include "common.cuh"
#include "rowops.cuh"
#include "regex_bits.h"
#include <cub/cub.cuh>

namespace custr {

constexpr int SPLIT_THREADS = 256;
constexpr int DELIM_SMEM = 256;

struct SplitParams {
    const uint8_t* delim;    // device copy of the delimiter bytes (split) — nullptr => whitespace
    int m;                   // delimiter byte length
    int limit;               // max tokens (maxsplit+1) or 0
    const uint32_t* set;     // tokenize: packed delimiter chars (nullptr => whitespace)
    int set_count;
    int mode;                // 0 = split(delim), 1 = split(ws), 2 = tokenize
    int record;              // split_record flavour of the whitespace placeholder ("" instead of null)
    int right;               // rsplit / rsplit_record (split.cu:435-735,960-1160)
    int ncols;               // rsplit(whitespace): column count of the whole call (see row::rwsplit_column)
};

__device__ __forceinline__ int row_token_count(const SplitParams& P, const uint8_t* s, int n)
{
    if (P.mode == 0) return row::split_count(s, n, P.delim, P.m, P.limit);
    if (P.mode == 1) return row::wsplit_count(s, n, P.limit);
    row::TokenWalk w;
    w.init(s, n, row::DelimSet{P.set, P.set_count});
    int b, e, c = 0;
    while (w.next(b, e)) ++c;
    return c;
}

// generic walker: calls f(k, begin, end, is_null) for each token of the row (k ascending, descending for the right-to-left forms)
template <typename F>
__device__ __forceinline__ void row_walk(const SplitParams& P, const uint8_t* s, int n, int dcount, F f)
{
    int b, e;
    if (P.right && P.mode == 0) {
        row::rsplit_walk(s, n, P.delim, P.m, dcount, [&](int k, int tb, int te) {
            if (tb >= te) tb = te = 0;  // empty (not null) token
            f(k, tb, te, false);
        });
    } else if (P.right && P.mode == 1 && P.record) {
        row::rwsplit_record_walk(s, n, dcount, P.limit, [&](int k, int tb, int te) { f(k, tb, te, false); });
    } else if (P.right && P.mode == 1) {
        for (int k = 0; k < dcount; ++k) {
            const bool ok = row::rwsplit_column(s, n, dcount, P.ncols, P.limit, k, b, e);
            f(k, ok ? b : 0, ok ? e : 0, !ok);
        }
    } else if (P.mode == 0) {
        row::SplitWalk w;
        w.init(s, n, P.delim, P.m, dcount);
        for (int k = 0; w.next(b, e); ++k) f(k, b, e, false);
    } else if (P.mode == 1) {
        row::WsWalk w;
        w.init(s, n, P.limit);
        bool ph;
        for (int k = 0; k < dcount && w.next(b, e, ph); ++k) f(k, b, e, ph && !P.record);
    } else {
        row::TokenWalk w;
        w.init(s, n, row::DelimSet{P.set, P.set_count});
        for (int k = 0; k < dcount && w.next(b, e); ++k) f(k, b, e, false);
    }
}

__global__ void __launch_bounds__(SPLIT_THREADS)
k_token_counts(ColView col, SplitParams P, int32_t* __restrict__ counts)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        int c = 0;
        if (col.valid(i)) c = row_token_count(P, (const uint8_t*)col.chars + col.offsets[i], col.offsets[i + 1] - col.offsets[i]);
        counts[i] = c;
    }
}

// column-major: lens[c*(n+1)+row], valid[c*n+row]
__global__ void __launch_bounds__(SPLIT_THREADS)
k_split_lengths(ColView col, SplitParams P, const int32_t* __restrict__ counts, int ncols, int32_t* __restrict__ lens,
                uint8_t* __restrict__ valid)
{
    const size_t n = col.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        int dcount = counts[i];
        int wrote = 0;
        if (dcount > 0) {
            const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
            row_walk(P, s, col.offsets[i + 1] - col.offsets[i], dcount, [&](int k, int b, int e, bool is_null) {
                lens[(size_t)k * (n + 1) + i] = is_null ? 0 : e - b;
                valid[(size_t)k * n + i] = !is_null;
                wrote = wrote > k + 1 ? wrote : k + 1;
            });
        }
        for (int k = wrote; k < ncols; ++k) { lens[(size_t)k * (n + 1) + i] = 0; valid[(size_t)k * n + i] = 0; }
    }
}

__global__ void __launch_bounds__(SPLIT_THREADS)
k_split_copy(ColView col, SplitParams P, const int32_t* __restrict__ counts, const ColumnOut* __restrict__ outs)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        int dcount = counts[i];
        if (dcount <= 0) continue;
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        row_walk(P, s, col.offsets[i + 1] - col.offsets[i], dcount, [&](int k, int b, int e, bool is_null) {
            if (is_null) return;
            char* o = outs[k].chars + outs[k].offsets[i];
            for (int j = b; j < e; ++j) *o++ = (char)s[j];
        });
    }
}

// row-major (split_record / tokenize): token lengths at tlens[row_off[i] + k]
__global__ void __launch_bounds__(SPLIT_THREADS)
k_record_lengths(ColView col, SplitParams P, const int32_t* __restrict__ row_off, int32_t* __restrict__ tlens)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        int first = row_off[i], dcount = row_off[i + 1] - first;
        if (dcount <= 0) continue;
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        int wrote = 0;
        row_walk(P, s, col.offsets[i + 1] - col.offsets[i], dcount, [&](int k, int b, int e, bool) { tlens[first + k] = e - b; wrote = wrote > k + 1 ? wrote : k + 1; });
        for (int k = wrote; k < dcount; ++k) tlens[first + k] = 0;
    }
}

__global__ void __launch_bounds__(SPLIT_THREADS)
k_record_copy(ColView col, SplitParams P, const int32_t* __restrict__ row_off, const int32_t* __restrict__ tok_off,
              char* __restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        int first = row_off[i], dcount = row_off[i + 1] - first;
        if (dcount <= 0) continue;
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        row_walk(P, s, col.offsets[i + 1] - col.offsets[i], dcount, [&](int k, int b, int e, bool) {
            char* o = out + tok_off[first + k];
            for (int j = b; j < e; ++j) *o++ = (char)s[j];
        });
    }
}

static inline int row_grid(int n)
{
    int want = (n + SPLIT_THREADS - 1) / SPLIT_THREADS;
    int cap = num_sms() * 32;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

struct ParamHolder {
    SplitParams P{};
    BufPtr delim_buf, set_buf;
};

static void make_split_params(ParamHolder& h, const char* delimiter, int maxsplit, bool record, bool right = false)
{
    h.P.limit = maxsplit > 0 ? maxsplit + 1 : 0;
    h.P.record = record;
    h.P.right = right;
    if (!delimiter) { h.P.mode = 1; return; }
    h.P.mode = 0;
    h.P.m = (int)strlen(delimiter);
    h.delim_buf = upload(delimiter, h.P.m + 1);
    h.P.delim = (const uint8_t*)h.delim_buf->ptr;
}

static void make_token_params(ParamHolder& h, const char* delimiter)
{
    h.P.mode = 2;
    if (!delimiter) return;  // whitespace
    std::vector<uint32_t> set;
    const uint8_t* p = (const uint8_t*)delimiter;
    size_t n = strlen(delimiter), i = 0;
    while (i < n) {
        int w = 1 + ((p[i] & 0xF0) == 0xF0) + ((p[i] & 0xE0) == 0xE0) + ((p[i] & 0xC0) == 0xC0);
        uint32_t c = p[i];
        for (int k = 1; k < w && i + k < n; ++k) c = (c << 8) | p[i + k];
        set.push_back(c);
        i += w;
    }
    set.push_back(0xFFFFFFFFu);  // keeps the pointer non-null for an empty set ("" delimits nothing)
    h.set_buf = upload(set.data(), set.size() * 4);
    h.P.set = (const uint32_t*)h.set_buf->ptr;
    h.P.set_count = (int)set.size() - 1;
}

static int max_of(const int32_t* d, int n)
{
    Scratch<int32_t> out(1);
    size_t tmp_bytes = 0;
    cub::DeviceReduce::Max(nullptr, tmp_bytes, d, out.get(), n, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes);
    CUSTR_CUDA(cub::DeviceReduce::Max(tmp->ptr, tmp_bytes, d, out.get(), n, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    int32_t h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, out.get(), 4, cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h;
}

static int split_columns(const custr_column* col, const char* delimiter, int maxsplit, custr_column** out, int cap, bool right = false)
{
    if (!col || (!out && cap > 0)) return fail(CUSTR_ERR_ARG, "split: null argument");
    int32_t n = col->n;
    if (n == 0) return 0;
    ParamHolder h;
    make_split_params(h, delimiter, maxsplit, false, right);
    Scratch<int32_t> counts((size_t)n + 1);
    LAUNCH(k_token_counts, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, counts.get());
    int ncols = max_of(counts.get(), n);
    h.P.ncols = ncols;
    if (ncols == 0) {  // every row null: one all-null column (split.cu:756-757)
        if (cap > 0) out[0] = all_null_column(n);
        return 1;
    }
    Scratch<int32_t> lens((size_t)ncols * (n + 1));
    Scratch<uint8_t> valid((size_t)ncols * n);
    for (int c = 0; c < ncols; ++c) CUSTR_CUDA(cudaMemsetAsync(lens.get() + (size_t)c * (n + 1) + n, 0, 4, g_stream));
    LAUNCH(k_split_lengths, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, (const int32_t*)counts.get(), ncols, lens.get(), valid.get());
    std::vector<BufPtr> offs(ncols), chars(ncols);
    std::vector<int64_t> totals(ncols);
    std::vector<ColumnOut> outs(ncols);
    for (int c = 0; c < ncols; ++c) {
        offs[c] = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
        totals[c] = scan_lengths_to_offsets(lens.get() + (size_t)c * (n + 1), (int32_t*)offs[c]->ptr, n);
        chars[c] = dev_alloc((size_t)totals[c]);
        outs[c] = ColumnOut{(char*)chars[c]->ptr, (const int32_t*)offs[c]->ptr};
    }
    BufPtr d_outs = upload(outs.data(), sizeof(ColumnOut) * ncols);
    LAUNCH(k_split_copy, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, (const int32_t*)counts.get(), (const ColumnOut*)d_outs->ptr);
    for (int c = 0; c < ncols; ++c) {
        BufPtr bits = dev_alloc((n + 7) / 8);
        pack_bits(valid.get() + (size_t)c * n, (uint8_t*)bits->ptr, n);
        int32_t nulls = count_zero_bits((const uint8_t*)bits->ptr, 0, n);
        custr_column* r = make_column(chars[c], offs[c], bits, n, nulls, totals[c]);
        if (c < cap) out[c] = r;
        else custr_column_free(r);
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return ncols;
}

// shared by split_record and tokenize: flat token column + row offsets
static custr_column* flat_tokens(const custr_column* col, const ParamHolder& h, int32_t* row_offsets_user, int devmem, int* total_out)
{
    int32_t n = col->n;
    Scratch<int32_t> counts((size_t)n + 1);
    CUSTR_CUDA(cudaMemsetAsync(counts.get() + n, 0, 4, g_stream));
    LAUNCH(k_token_counts, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, counts.get());
    Scratch<int32_t> row_off((size_t)n + 1);
    int64_t ntok = scan_lengths_to_offsets(counts.get(), row_off.get(), n);
    Scratch<int32_t> tlens((size_t)ntok + 1);
    CUSTR_CUDA(cudaMemsetAsync(tlens.get() + ntok, 0, 4, g_stream));
    if (ntok) LAUNCH(k_record_lengths, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, (const int32_t*)row_off.get(), tlens.get());
    BufPtr tok_off = dev_alloc(sizeof(int32_t) * (size_t)(ntok + 1));
    int64_t total = scan_lengths_to_offsets(tlens.get(), (int32_t*)tok_off->ptr, (int32_t)ntok);
    BufPtr chars = dev_alloc((size_t)total);
    if (ntok) LAUNCH(k_record_copy, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, (const int32_t*)row_off.get(),
                     (const int32_t*)tok_off->ptr, (char*)chars->ptr);
    if (row_offsets_user) {
        cudaMemcpyKind kind = devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        CUSTR_CUDA(cudaMemcpyAsync(row_offsets_user, row_off.get(), sizeof(int32_t) * (size_t)(n + 1), kind, g_stream));
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    *total_out = (int)ntok;
    return make_column(chars, tok_off, nullptr, (int32_t)ntok, 0, total);
}

// partition / rpartition (split.cu:1165-1262,1268-1372): three strings per row, row-major in ONE column of 3n rows:
// [left, delimiter, right] around the first (partition) or last (rpartition) delimiter; without a delimiter
// [row, "", ""] (partition) or ["", "", row] (rpartition); a null row gives three nulls.
__global__ void __launch_bounds__(SPLIT_THREADS)
k_partition(ColView col, const uint8_t* __restrict__ d, int m, int right, int32_t* __restrict__ lens, const int32_t* __restrict__ off,
            char* __restrict__ out, uint8_t* __restrict__ valid)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        const bool ok = col.valid(i);
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const int n = col.offsets[i + 1] - col.offsets[i];
        int b[3] = {0, 0, 0}, e[3] = {0, 0, 0};
        bool delim_piece = false;
        if (ok) {
            const int p = right ? row::rfind_bytes(s, 0, n, d, m) : row::find_bytes(s, 0, n, d, m);
            if (p >= 0) { e[0] = p; delim_piece = true; b[2] = p + m; e[2] = n; }
            else if (right) { e[2] = n; }
            else { e[0] = n; }
        }
        for (int k = 0; k < 3; ++k) {
            const int len = (k == 1) ? (delim_piece ? m : 0) : e[k] - b[k];
            if (!out) { lens[3 * (size_t)i + k] = len; valid[3 * (size_t)i + k] = ok; }
            else {
                char* o = out + off[3 * (size_t)i + k];
                if (k == 1) for (int j = 0; j < len; ++j) o[j] = (char)d[j];
                else for (int j = 0; j < len; ++j) o[j] = (char)s[b[k] + j];
            }
        }
    }
}

static custr_column* partition_rows(const custr_column* col, const char* delimiter, bool right)
{
    const int32_t n = col->n;
    if (n == 0) return custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
    if ((int64_t)n * 3 > 0x7fffffff) throw ArgError{fail(CUSTR_ERR_INVALID, "partition: too many rows")};
    const int m = (int)strlen(delimiter);
    BufPtr d = upload(delimiter, m);
    const int32_t n3 = 3 * n;
    Scratch<int32_t> lens((size_t)n3 + 1);
    Scratch<uint8_t> valid((size_t)n3);
    CUSTR_CUDA(cudaMemsetAsync(lens.get() + n3, 0, 4, g_stream));
    LAUNCH(k_partition, row_grid(n), SPLIT_THREADS, 0, view_of(col), (const uint8_t*)d->ptr, m, right ? 1 : 0, lens.get(), (const int32_t*)nullptr,
           (char*)nullptr, valid.get());
    BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n3 + 1));
    const int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n3);
    BufPtr chars = dev_alloc((size_t)total);
    LAUNCH(k_partition, row_grid(n), SPLIT_THREADS, 0, view_of(col), (const uint8_t*)d->ptr, m, right ? 1 : 0, (int32_t*)nullptr,
           (const int32_t*)off->ptr, (char*)chars->ptr, (uint8_t*)nullptr);
    BufPtr bits;
    int32_t nulls = 0;
    if (col->nulls) {
        bits = dev_alloc((n3 + 7) / 8);
        pack_bits(valid.get(), (uint8_t*)bits->ptr, n3);
        nulls = 3 * col->nulls;
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return make_column(chars, off, bits, n3, nulls, total);
}

}  // namespace custr

using namespace custr;

extern "C" {

int custr_split(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** out, int32_t cap)
{
    return guarded([&] { return split_columns(col, delimiter, maxsplit, out, cap); }, (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_rsplit(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** out, int32_t cap)
{
    return guarded([&] { return split_columns(col, delimiter, maxsplit, out, cap, true); }, (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_split_record(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** tokens, int32_t* row_offsets,
                       int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !tokens) return fail(CUSTR_ERR_ARG, "split_record: null argument");
            // one ASCII delimiter byte, no split limit: bit-stream compaction (split_bits.cuh), no per-row walk
            if (delimiter && delimiter[0] && !delimiter[1] && (unsigned char)delimiter[0] < 0x80 && maxsplit < 0 && !bits::g_force_generic) {
                BufPtr chars, off, row_off;
                int64_t ntok = 0, nbytes = 0;
                if (bits::split_record_flat(col, (uint8_t)delimiter[0], chars, off, row_off, ntok, nbytes)) {
                    if (row_offsets) {
                        CUSTR_CUDA(cudaMemcpyAsync(row_offsets, row_off->ptr, sizeof(int32_t) * (size_t)(col->n + 1),
                                                   devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, g_stream));
                        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                    }
                    *tokens = make_column(chars, off, nullptr, (int32_t)ntok, 0, nbytes);
                    return (int)ntok;
                }
            }
            ParamHolder h;
            make_split_params(h, delimiter, maxsplit, true);
            int total = 0;
            *tokens = flat_tokens(col, h, row_offsets, devmem, &total);
            return total;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_rsplit_record(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** tokens, int32_t* row_offsets,
                        int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !tokens) return fail(CUSTR_ERR_ARG, "rsplit_record: null argument");
            ParamHolder h;
            make_split_params(h, delimiter, maxsplit, true, true);
            int total = 0;
            *tokens = flat_tokens(col, h, row_offsets, devmem, &total);
            return total;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_partition(const custr_column* col, const char* delimiter, int right)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "partition: null column")};
            if (!delimiter || !*delimiter) throw ArgError{fail(CUSTR_ERR_INVALID, "partition: delimiter is null or empty")};
            return partition_rows(col, delimiter, right != 0);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_tokenize(const custr_column* col, const char* delimiter)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col) throw ArgError{fail(CUSTR_ERR_ARG, "tokenize: null column")};
            // whitespace or a few ASCII delimiter bytes: bit-stream compaction, no per-row walk (tokenize_bits.cuh)
            bool ascii_set = delimiter != nullptr && strlen(delimiter) >= 1 && strlen(delimiter) <= 8;
            for (const char* p = delimiter; ascii_set && *p; ++p) ascii_set = (unsigned char)*p < 0x80;
            if ((!delimiter || ascii_set) && !bits::g_force_generic) {
                BufPtr chars, off;
                int64_t ntok = 0, nbytes = 0;
                if (bits::tokenize_flat(col, (const uint8_t*)delimiter, delimiter ? (int)strlen(delimiter) : 0, chars, off, ntok, nbytes))
                    return make_column(chars, off, nullptr, (int32_t)ntok, 0, nbytes);
            }
            ParamHolder h;
            make_token_params(h, delimiter);
            int total = 0;
            return flat_tokens(col, h, nullptr, 0, &total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

int custr_token_count(const custr_column* col, const char* delimiter, uint32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !results) return fail(CUSTR_ERR_ARG, "token_count: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            ParamHolder h;
            make_token_params(h, delimiter);
            ResultBuf<uint32_t> out(results, n, devmem);
            LAUNCH(k_token_counts, row_grid(n), SPLIT_THREADS, 0, view_of(col), h.P, (int32_t*)out.dev);
            out.finish();
            return 0;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

}  // extern "C"
