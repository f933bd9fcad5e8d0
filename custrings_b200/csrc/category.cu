// NVCategory key build: keys = sorted distinct strings (null first), values = int32 index of each row's key.
// Replaces NVCategory::create_from_strings NVCategory.cu:327-356 -> NVCategoryImpl_init :220-304 (a string-compare
// merge sort over all N rows + un-sort + unique), get_keys :724-750, get_values :866-878, values_cptr :880-883.
//
// B200 design: only the K distinct keys are ever compared as strings.
//   1. 64-bit hash per row                                  (reads chars once, coalesced per warp)
//   2. radix sort (hash,row) pairs                          (cub, 8 passes over 12 B/row)
//   3. adjacent compare: group heads + EXACT byte check inside equal-hash runs (collision => re-hash, new seed)
//   4. scan heads -> group ids, K representatives
//   5. merge-sort the K representatives with the reference's comparator (custring.inl:240-261, null first)
//   6. scatter ranks back: values[row] = rank[group(row)]
// Observable contract is identical to the reference (keys order, null key first, values); the hash is internal.
#include "common.cuh"
#include "rowops.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <mutex>

namespace custr {

constexpr int CAT_THREADS = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// one warp per row would waste lanes on 16-byte keys: thread-per-row, 8 bytes at a time
__global__ void __launch_bounds__(CAT_THREADS)
k_hash_rows(ColView col, uint64_t seed, uint64_t* __restrict__ hashes, int32_t* __restrict__ rows)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        uint64_t h = 0;
        if (col.valid(i)) {
            const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
            int n = col.offsets[i + 1] - col.offsets[i];
            h = seed ^ ((uint64_t)n * 0x9E3779B97F4A7C15ULL);
            int k = 0;
            for (; k + 8 <= n; k += 8) {
                uint64_t v = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) v |= (uint64_t)s[k + j] << (8 * j);
                h = mix64(h ^ v) + 0x9E3779B97F4A7C15ULL;
            }
            uint64_t v = 0;
            for (int j = 0; k + j < n; ++j) v |= (uint64_t)s[k + j] << (8 * j);
            h = mix64(h ^ v);
            if (h == 0) h = 1;  // 0 is reserved for null rows
        }
        hashes[i] = h;
        rows[i] = i;
    }
}

__device__ __forceinline__ bool rows_equal(const ColView& col, int a, int b)
{
    bool va = col.valid(a), vb = col.valid(b);
    if (!va || !vb) return va == vb;
    int ao = col.offsets[a], an = col.offsets[a + 1] - ao;
    int bo = col.offsets[b], bn = col.offsets[b + 1] - bo;
    if (an != bn) return false;
    return row::bytes_equal((const uint8_t*)col.chars + ao, (const uint8_t*)col.chars + bo, an);
}

__global__ void __launch_bounds__(CAT_THREADS)
k_group_heads(ColView col, const uint64_t* __restrict__ hashes, const int32_t* __restrict__ rows, int32_t* __restrict__ heads,
              int* __restrict__ collision)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < col.n; j += gridDim.x * blockDim.x) {
        int head = 1;
        if (j > 0 && hashes[j] == hashes[j - 1]) {
            head = 0;
            if (!rows_equal(col, rows[j], rows[j - 1])) *collision = 1;
        }
        heads[j] = head;
    }
}

// gid = inclusive_scan(heads) - 1 (in place array `gids`); representatives[gid] = row at each head
__global__ void __launch_bounds__(CAT_THREADS)
k_representatives(const int32_t* __restrict__ heads, const int32_t* __restrict__ gids, const int32_t* __restrict__ rows, int n,
                  int32_t* __restrict__ reps, int32_t* __restrict__ group_ids)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        if (heads[j]) { int g = gids[j] - 1; reps[g] = rows[j]; group_ids[g] = g; }
}

// ---- hash-table path: few distinct keys (the usual shape of categorical data) ------------------------------------------------
// Pass 1 hashes every row and inserts the hash into an open-addressing table (linear probing; a plain read first, the CAS
// only while the slot still looks empty, so the ~N lookups of K << N distinct keys are cache hits, not contended atomics).
// The thread that wins a slot registers its row as the key's representative.  Pass 2 (after the K representatives were
// sorted with the reference's comparator) compares every row with its representative BYTE FOR BYTE — a 64-bit collision
// or a table that fills beyond half sends the whole build to the sort path below — and writes values[row] = rank.
constexpr int HT_BITS = 20;  // 1 Mi slots: up to 512 Ki distinct keys
__global__ void __launch_bounds__(CAT_THREADS)
k_ht_insert(ColView col, uint64_t seed, unsigned long long* __restrict__ tab, int32_t* __restrict__ slot_group, int32_t* __restrict__ reps,
            int32_t* __restrict__ nkeys, int32_t* __restrict__ row_slot, int32_t* __restrict__ null_rep)
{
    const uint32_t mask = (1u << HT_BITS) - 1u;
    const int32_t max_keys = 1 << (HT_BITS - 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        if (*(volatile int32_t*)nkeys > max_keys) { row_slot[i] = -2; continue; }  // too many distinct keys: the host switches paths
        if (!col.valid(i)) {
            row_slot[i] = -1;
            if (*(volatile int32_t*)null_rep < 0) atomicCAS(null_rep, -1, i);
            continue;
        }
        const uint8_t* s = (const uint8_t*)col.chars + col.offsets[i];
        const int n = col.offsets[i + 1] - col.offsets[i];
        uint64_t h = seed ^ ((uint64_t)n * 0x9E3779B97F4A7C15ULL);
        int k = 0;
        for (; k + 8 <= n; k += 8) {
            uint64_t v = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) v |= (uint64_t)s[k + j] << (8 * j);
            h = mix64(h ^ v) + 0x9E3779B97F4A7C15ULL;
        }
        uint64_t v = 0;
        for (int j = 0; k + j < n; ++j) v |= (uint64_t)s[k + j] << (8 * j);
        h = mix64(h ^ v);
        if (h == 0) h = 1;  // 0 = empty slot
        uint32_t slot = (uint32_t)(h >> 20) & mask;
        for (int probe = 0;; ++probe) {
            unsigned long long cur = *(volatile unsigned long long*)(tab + slot);
            if (cur == 0) {
                cur = atomicCAS(tab + slot, 0ull, (unsigned long long)h);
                if (cur == 0) {  // this thread created the key
                    const int g = atomicAdd(nkeys, 1);
                    if (g < max_keys) reps[g] = i;
                    slot_group[slot] = g;
                    cur = h;
                }
            }
            if (cur == h) { row_slot[i] = (int32_t)slot; break; }
            slot = (slot + 1) & mask;
            if ((probe & 63) == 63 && *(volatile int32_t*)nkeys > max_keys) { row_slot[i] = -2; break; }
        }
    }
}

__global__ void __launch_bounds__(CAT_THREADS)
k_ht_values(ColView col, const int32_t* __restrict__ row_slot, const int32_t* __restrict__ slot_group, const int32_t* __restrict__ reps,
            const int32_t* __restrict__ rank, int null_group, int32_t* __restrict__ values, int* __restrict__ collision)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < col.n; i += gridDim.x * blockDim.x) {
        const int slot = row_slot[i];
        if (slot < 0) { values[i] = slot == -1 ? rank[null_group] : 0; if (slot == -2) *collision = 1; continue; }
        const int g = slot_group[slot];
        if (!rows_equal(col, i, reps[g])) *collision = 1;  // two different strings with one 64-bit hash
        values[i] = rank[g];
    }
}

struct KeyLess {
    ColView col;
    __device__ bool operator()(const int32_t& a, const int32_t& b) const
    {
        bool va = col.valid(a), vb = col.valid(b);
        if (!va || !vb) return !va && vb;  // null sorts first
        int ao = col.offsets[a], bo = col.offsets[b];
        return row::compare_bytes((const uint8_t*)col.chars + ao, col.offsets[a + 1] - ao, (const uint8_t*)col.chars + bo,
                                  col.offsets[b + 1] - bo) < 0;
    }
};

__global__ void k_iota(int32_t* __restrict__ a, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
// after the sort: sorted_reps[r] is the representative of group sorted_groups[r]
__global__ void k_scatter_reps(const int32_t* __restrict__ sorted_reps, const int32_t* __restrict__ sorted_groups, int k, int32_t* __restrict__ rep_of_group)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < k) rep_of_group[sorted_groups[r]] = sorted_reps[r];
}
__global__ void k_rank_of_group(const int32_t* __restrict__ sorted_groups, int k, int32_t* __restrict__ rank)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < k) rank[sorted_groups[r]] = r;
}

__global__ void __launch_bounds__(CAT_THREADS)
k_scatter_values(const int32_t* __restrict__ rows, const int32_t* __restrict__ gids, const int32_t* __restrict__ rank, int n,
                 int32_t* __restrict__ values)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) values[rows[j]] = rank[gids[j] - 1];
}

// concat helper
__global__ void k_concat_offsets(const int32_t* __restrict__ src, int n, int32_t first_off, int32_t base_bytes, int32_t* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) dst[i] = src[i] - first_off + base_bytes;
}
__global__ void k_valid_flags(ColView c, uint8_t* __restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.n) flags[i] = c.valid(i);
}

// binary search of each local key in the (sorted, distinct) global keys
__global__ void k_lookup_keys(ColView local, ColView global, int32_t* __restrict__ map)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= local.n) return;
    if (!local.valid(i)) { map[i] = (global.n > 0 && !global.valid(0)) ? 0 : -1; return; }  // the null key is key 0 wherever it exists
    const uint8_t* s = (const uint8_t*)local.chars + local.offsets[i];
    int n = local.offsets[i + 1] - local.offsets[i];
    int lo = 0, hi = global.n - 1, found = -1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int c;
        if (!global.valid(mid)) c = 1;
        else {
            int go = global.offsets[mid];
            c = row::compare_bytes(s, n, (const uint8_t*)global.chars + go, global.offsets[mid + 1] - go);
        }
        if (c == 0) { found = mid; break; }
        if (c < 0) hi = mid - 1; else lo = mid + 1;
    }
    map[i] = found;
}
__global__ void k_remap_values(const int32_t* __restrict__ in, const int32_t* __restrict__ map, int n, int32_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] < 0 ? in[i] : map[in[i]];  // -1 = "no key" (after remove_keys / set_keys) stays -1
}
// used[v] = 1 for every value v in [0, k); *bad counts the values outside [lo, k)
__global__ void k_mark_used(const int32_t* __restrict__ vals, int n, int k, int lo, uint8_t* __restrict__ used, int* __restrict__ bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = vals[i];
    if (v < lo || v >= k) atomicAdd(bad, 1);
    else if (v >= 0 && used) used[v] = 1;
}

// byte length of every key, -1 for the null key (payload of the sharded key exchange)
__global__ void k_key_lengths(ColView keys, int32_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < keys.n) out[i] = keys.valid(i) ? keys.offsets[i + 1] - keys.offsets[i] : -1;
}

static inline int row_grid(int n)
{
    int want = (n + CAT_THREADS - 1) / CAT_THREADS;
    int cap = num_sms() * 32;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

static custr_column* concat_columns(const custr_column* const* cols, int ncols)
{
    int64_t rows = 0, bytes = 0;
    int32_t nulls = 0;
    for (int k = 0; k < ncols; ++k) { rows += cols[k]->n; bytes += cols[k]->nbytes; nulls += cols[k]->nulls; }
    if (rows > 0x7fffffffLL || bytes > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "category: more than 2^31 rows or bytes")};
    BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(rows + 1)), chars = dev_alloc((size_t)bytes);
    Scratch<uint8_t> flags((size_t)rows + 1);
    int64_t r0 = 0, b0 = 0;
    for (int k = 0; k < ncols; ++k) {
        const custr_column* c = cols[k];
        if (c->nbytes) CUSTR_CUDA(cudaMemcpyAsync((char*)chars->ptr + b0, c->chars + c->first_off, (size_t)c->nbytes, cudaMemcpyDeviceToDevice, g_stream));
        LAUNCH(k_concat_offsets, (c->n + 256) / 256, 256, 0, c->offsets, c->n, c->first_off, (int32_t)b0, (int32_t*)off->ptr + r0);
        if (c->n) LAUNCH(k_valid_flags, (c->n + 255) / 256, 256, 0, view_of(c), flags.get() + r0);
        r0 += c->n;
        b0 += c->nbytes;
    }
    BufPtr bits;
    if (nulls) { bits = dev_alloc((rows + 7) / 8); pack_bits(flags.get(), (uint8_t*)bits->ptr, (int32_t)rows); }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return make_column(chars, off, bits, (int32_t)rows, nulls, bytes);
}

// false: more than 512 Ki distinct keys or a hash collision -> the caller takes the sort path
static bool build_category_hashed(const custr_column* col, custr_category* cat)
{
    const int32_t n = col->n;
    const size_t slots = (size_t)1 << HT_BITS;
    const int32_t max_keys = 1 << (HT_BITS - 1);
    BufPtr tab = dev_alloc(sizeof(unsigned long long) * slots);
    Scratch<int32_t> slot_group(slots), reps((size_t)max_keys + 1), row_slot(n), counters(2);
    Scratch<int> collision(1);
    CUSTR_CUDA(cudaMemsetAsync(tab->ptr, 0, sizeof(unsigned long long) * slots, g_stream));
    const int32_t init[2] = {0, -1};  // nkeys, null representative
    CUSTR_CUDA(cudaMemcpyAsync(counters.get(), init, sizeof(init), cudaMemcpyHostToDevice, g_stream));
    CUSTR_CUDA(cudaMemsetAsync(collision.get(), 0, sizeof(int), g_stream));
    LAUNCH(k_ht_insert, row_grid(n), CAT_THREADS, 0, view_of(col), 0x243F6A8885A308D3ULL, (unsigned long long*)tab->ptr, slot_group.get(),
           reps.get(), counters.get(), row_slot.get(), counters.get() + 1);
    int32_t h_cnt[2] = {0, -1};
    CUSTR_CUDA(cudaMemcpyAsync(h_cnt, counters.get(), sizeof(h_cnt), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    int32_t nkeys = h_cnt[0];
    if (nkeys > max_keys) return false;
    int null_group = 0;
    if (h_cnt[1] >= 0) {  // the null key: one more representative (it sorts first)
        null_group = nkeys;
        CUSTR_CUDA(cudaMemcpyAsync(reps.get() + nkeys, &h_cnt[1], sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
        ++nkeys;
    }
    Scratch<int32_t> groups(nkeys ? nkeys : 1), rank(nkeys ? nkeys : 1);
    if (nkeys) {
        LAUNCH(k_iota, (nkeys + 255) / 256, 256, 0, groups.get(), nkeys);
        size_t merge_bytes = 0;
        KeyLess less{view_of(col)};
        cub::DeviceMergeSort::SortPairs(nullptr, merge_bytes, reps.get(), groups.get(), nkeys, less, g_stream);
        BufPtr mtmp = dev_alloc(merge_bytes);
        CUSTR_CUDA(cub::DeviceMergeSort::SortPairs(mtmp->ptr, merge_bytes, reps.get(), groups.get(), nkeys, less, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        LAUNCH(k_rank_of_group, (nkeys + 255) / 256, 256, 0, (const int32_t*)groups.get(), nkeys, rank.get());
    }
    // reps is now in key order; the value pass needs the representative of a GROUP: keep an unsorted copy
    // (sorted reps feed the keys gather, group -> representative goes through rank)
    Scratch<int32_t> rep_of_group(nkeys ? nkeys : 1);
    if (nkeys) LAUNCH(k_scatter_reps, (nkeys + 255) / 256, 256, 0, (const int32_t*)reps.get(), (const int32_t*)groups.get(), nkeys, rep_of_group.get());
    LAUNCH(k_ht_values, row_grid(n), CAT_THREADS, 0, view_of(col), (const int32_t*)row_slot.get(), (const int32_t*)slot_group.get(),
           (const int32_t*)rep_of_group.get(), (const int32_t*)rank.get(), null_group, (int32_t*)cat->values_buf->ptr, collision.get());
    int hit = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&hit, collision.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    if (hit) return false;
    cat->keys = custr_gather(col, reps.get(), nkeys, 1);
    if (!cat->keys) throw CudaError{cudaErrorUnknown};
    cat->has_null_key = col->nulls > 0;
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return true;
}

static custr_category* build_category(const custr_column* col)
{
    const int32_t n = col->n;
    custr_category* cat = new custr_category;
    std::unique_ptr<custr_category, void (*)(custr_category*)> guard(cat, [](custr_category* c) { custr_category_free(c); });
    cat->n = n;
    cat->values_buf = dev_alloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    if (n == 0) {
        cat->keys = custr_create_from_offsets(nullptr, 0, nullptr, nullptr, 0, 0);
        return guard.release();
    }
    if (build_category_hashed(col, cat)) return guard.release();
    Scratch<uint64_t> h_in(n), h_out(n);
    Scratch<int32_t> r_in(n), r_out(n), heads(n), gids(n);
    Scratch<int> collision(1);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, h_in.get(), h_out.get(), r_in.get(), r_out.get(), n, 0, 64, g_stream);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, heads.get(), gids.get(), n, g_stream);
    BufPtr tmp = dev_alloc(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
    uint64_t seed = 0x243F6A8885A308D3ULL;
    for (int attempt = 0;; ++attempt) {
        LAUNCH(k_hash_rows, row_grid(n), CAT_THREADS, 0, view_of(col), seed, h_in.get(), r_in.get());
        CUSTR_CUDA(cub::DeviceRadixSort::SortPairs(tmp->ptr, sort_bytes, h_in.get(), h_out.get(), r_in.get(), r_out.get(), n, 0, 64, g_stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUSTR_CUDA(cudaMemsetAsync(collision.get(), 0, sizeof(int), g_stream));
        LAUNCH(k_group_heads, row_grid(n), CAT_THREADS, 0, view_of(col), (const uint64_t*)h_out.get(), (const int32_t*)r_out.get(),
               heads.get(), collision.get());
        int hit = 0;
        CUSTR_CUDA(cudaMemcpyAsync(&hit, collision.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
        if (!hit) break;
        if (attempt == 3) throw ArgError{fail(CUSTR_ERR_INVALID, "category: persistent 64-bit hash collisions")};
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    CUSTR_CUDA(cub::DeviceScan::InclusiveSum(tmp->ptr, scan_bytes, heads.get(), gids.get(), n, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    int32_t nkeys = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&nkeys, gids.get() + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    Scratch<int32_t> reps(nkeys), groups(nkeys), rank(nkeys);
    LAUNCH(k_representatives, row_grid(n), CAT_THREADS, 0, (const int32_t*)heads.get(), (const int32_t*)gids.get(),
           (const int32_t*)r_out.get(), n, reps.get(), groups.get());
    size_t merge_bytes = 0;
    KeyLess less{view_of(col)};
    cub::DeviceMergeSort::SortPairs(nullptr, merge_bytes, reps.get(), groups.get(), nkeys, less, g_stream);
    BufPtr mtmp = dev_alloc(merge_bytes);
    CUSTR_CUDA(cub::DeviceMergeSort::SortPairs(mtmp->ptr, merge_bytes, reps.get(), groups.get(), nkeys, less, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LAUNCH(k_rank_of_group, (nkeys + 255) / 256, 256, 0, (const int32_t*)groups.get(), nkeys, rank.get());
    LAUNCH(k_scatter_values, row_grid(n), CAT_THREADS, 0, (const int32_t*)r_out.get(), (const int32_t*)gids.get(),
           (const int32_t*)rank.get(), n, (int32_t*)cat->values_buf->ptr);
    cat->keys = custr_gather(col, reps.get(), nkeys, 1);
    if (!cat->keys) throw CudaError{cudaErrorUnknown};
    cat->has_null_key = col->nulls > 0;
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return guard.release();
}

}  // namespace custr

using namespace custr;

extern "C" {

custr_category* custr_category_create(const custr_column* const* cols, int32_t ncols)
{
    return guarded(
        [&]() -> custr_category* {
            if (!cols || ncols <= 0) throw ArgError{fail(CUSTR_ERR_ARG, "category: no input columns")};
            for (int k = 0; k < ncols; ++k)
                if (!cols[k]) throw ArgError{fail(CUSTR_ERR_ARG, "category: null input column")};
            if (ncols == 1) return build_category(cols[0]);
            std::unique_ptr<custr_column> all(concat_columns(cols, ncols));
            return build_category(all.get());
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}

void custr_category_free(custr_category* cat)
{
    if (!cat) return;
    custr_column_free(cat->keys);
    delete cat;
}

uint32_t custr_category_size(const custr_category* cat) { return cat ? (uint32_t)cat->n : 0; }
uint32_t custr_category_keys_size(const custr_category* cat) { return cat && cat->keys ? (uint32_t)cat->keys->n : 0; }

custr_column* custr_category_keys(const custr_category* cat)
{
    if (!cat || !cat->keys) return nullptr;
    return custr_slice_rows(cat->keys, 0, cat->keys->n);  // shares the immutable buffers
}

int custr_category_values(const custr_category* cat, int32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!cat || !results) return fail(CUSTR_ERR_ARG, "get_values: null argument");
            if (cat->n == 0) return 0;
            cudaMemcpyKind kind = devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            CUSTR_CUDA(cudaMemcpyAsync(results, cat->values_buf->ptr, sizeof(int32_t) * (size_t)cat->n, kind, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return cat->n;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

const int32_t* custr_category_values_cptr(const custr_category* cat) { return cat ? (const int32_t*)cat->values_buf->ptr : nullptr; }

custr_category* custr_category_remap_to_union(const custr_category* cat, const custr_column* all_keys)
{
    return guarded(
        [&]() -> custr_category* {
            if (!cat || !all_keys) throw ArgError{fail(CUSTR_ERR_ARG, "remap_to_union: null argument")};
            std::unique_ptr<custr_category, void (*)(custr_category*)> uni(build_category(all_keys), [](custr_category* c) { custr_category_free(c); });
            custr_category* out = new custr_category;
            out->n = cat->n;
            out->values_buf = dev_alloc(sizeof(int32_t) * (size_t)(cat->n ? cat->n : 1));
            out->keys = uni->keys;
            uni->keys = nullptr;
            out->has_null_key = out->keys->nulls > 0;
            int32_t k = cat->keys->n;
            if (k && cat->n) {
                Scratch<int32_t> map(k);
                LAUNCH(k_lookup_keys, (k + 255) / 256, 256, 0, view_of(cat->keys), view_of(out->keys), map.get());
                LAUNCH(k_remap_values, (cat->n + 255) / 256, 256, 0, (const int32_t*)cat->values_buf->ptr, (const int32_t*)map.get(),
                       cat->n, (int32_t*)out->values_buf->ptr);
                CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            }
            return out;
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}

// NVCategory::create_from_categories / merge_and_remap (NVCategory.cu:430-514,1339-1345; sorted != 0): keys = sorted
// distinct union of every input's keys, values = the inputs' values remapped and appended in input order.
// NVCategory::merge_category (NVCategory.cu:1223-1336; sorted == 0, exactly two inputs): keys = first input's keys followed by
// the second input's keys that are new (in their own order), first values unchanged, second values remapped and appended.
custr_category* custr_category_merge(const custr_category* const* cats, int32_t ncats, int sorted)
{
    return guarded(
        [&]() -> custr_category* {
            if (!cats || ncats <= 0) throw ArgError{fail(CUSTR_ERR_ARG, "category merge: no inputs")};
            if (!sorted && ncats != 2) throw ArgError{fail(CUSTR_ERR_ARG, "merge_category takes exactly two categories")};
            int64_t total = 0;
            std::vector<const custr_column*> keycols;
            for (int c = 0; c < ncats; ++c) {
                if (!cats[c] || !cats[c]->keys) throw ArgError{fail(CUSTR_ERR_ARG, "category merge: null input")};
                total += cats[c]->n;
                keycols.push_back(cats[c]->keys);
            }
            if (total > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "category merge: more than 2^31 values")};
            custr_category* out = new custr_category;
            std::unique_ptr<custr_category, void (*)(custr_category*)> guard(out, [](custr_category* c) { custr_category_free(c); });
            out->n = (int32_t)total;
            out->values_buf = dev_alloc(sizeof(int32_t) * (size_t)(total ? total : 1));
            std::vector<BufPtr> maps(ncats);
            if (sorted) {
                std::unique_ptr<custr_column> all(concat_columns(keycols.data(), ncats));
                std::unique_ptr<custr_category, void (*)(custr_category*)> uni(build_category(all.get()), [](custr_category* c) { custr_category_free(c); });
                out->keys = uni->keys;
                uni->keys = nullptr;
                for (int c = 0; c < ncats; ++c) {
                    const int32_t k = cats[c]->keys->n;
                    maps[c] = dev_alloc(sizeof(int32_t) * (size_t)(k ? k : 1));
                    if (k) LAUNCH(k_lookup_keys, (k + 255) / 256, 256, 0, view_of(cats[c]->keys), view_of(out->keys), (int32_t*)maps[c]->ptr);
                }
            } else {
                const custr_column *k1 = cats[0]->keys, *k2 = cats[1]->keys;
                const int32_t n1 = k1->n, n2 = k2->n;
                maps[1] = dev_alloc(sizeof(int32_t) * (size_t)(n2 ? n2 : 1));
                std::vector<int32_t> h_map((size_t)n2), fresh;
                // k1 may itself be the result of a merge_category (keys appended, NOT sorted; its null key anywhere), so it is
                // never binary-searched directly: the lookup goes through a sorted view of k1 (a category built over k1's keys:
                // sorted distinct keys + values[i] = rank of k1[i]) and the rank is mapped back to k1's own index.
                std::vector<int32_t> rank_to_k1;
                std::unique_ptr<custr_category, void (*)(custr_category*)> sorted1(nullptr, [](custr_category* c) { custr_category_free(c); });
                if (n1 && n2) {
                    sorted1.reset(build_category(k1));
                    std::vector<int32_t> rank((size_t)n1);
                    CUSTR_CUDA(cudaMemcpyAsync(rank.data(), sorted1->values_buf->ptr, sizeof(int32_t) * (size_t)n1, cudaMemcpyDeviceToHost, g_stream));
                    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                    rank_to_k1.assign((size_t)sorted1->keys->n, -1);
                    for (int32_t i = n1 - 1; i >= 0; --i) rank_to_k1[(size_t)rank[i]] = i;  // first occurrence wins
                }
                if (n2) {
                    if (sorted1) LAUNCH(k_lookup_keys, (n2 + 255) / 256, 256, 0, view_of(k2), view_of(sorted1->keys), (int32_t*)maps[1]->ptr);
                    else CUSTR_CUDA(cudaMemsetAsync(maps[1]->ptr, 0xff, sizeof(int32_t) * (size_t)n2, g_stream));
                    CUSTR_CUDA(cudaMemcpyAsync(h_map.data(), maps[1]->ptr, sizeof(int32_t) * (size_t)n2, cudaMemcpyDeviceToHost, g_stream));
                    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                }
                // the new keys are appended in SORTED order (the reference picks them out of the sorted keys2 + keys1 sequence,
                // NVCategory.cu:1262-1290), which is k2's own order only while k2 is a freshly built, sorted key set
                std::vector<int32_t> order2((size_t)n2);
                for (int32_t i = 0; i < n2; ++i) order2[(size_t)i] = i;
                if (n2 > 1) {
                    std::unique_ptr<custr_category, void (*)(custr_category*)> sorted2(build_category(k2), [](custr_category* c) { custr_category_free(c); });
                    std::vector<int32_t> rank2((size_t)n2);
                    CUSTR_CUDA(cudaMemcpyAsync(rank2.data(), sorted2->values_buf->ptr, sizeof(int32_t) * (size_t)n2, cudaMemcpyDeviceToHost, g_stream));
                    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                    std::stable_sort(order2.begin(), order2.end(), [&](int32_t a, int32_t b) { return rank2[(size_t)a] < rank2[(size_t)b]; });
                }
                for (int32_t i : order2) {  // key sets are small (dictionary entries, not rows): host bookkeeping
                    if (h_map[(size_t)i] >= 0) h_map[(size_t)i] = rank_to_k1[(size_t)h_map[(size_t)i]];
                    else { h_map[(size_t)i] = n1 + (int32_t)fresh.size(); fresh.push_back(i); }
                }
                if (n2) CUSTR_CUDA(cudaMemcpyAsync(maps[1]->ptr, h_map.data(), sizeof(int32_t) * (size_t)n2, cudaMemcpyHostToDevice, g_stream));
                std::unique_ptr<custr_column> added(custr_gather(k2, fresh.data(), (int32_t)fresh.size(), 0));
                if (!added) throw ArgError{CUSTR_ERR_INVALID};
                const custr_column* both[2] = {k1, added.get()};
                out->keys = concat_columns(both, 2);
                CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            }
            int64_t pos = 0;
            for (int c = 0; c < ncats; ++c) {
                const int32_t n = cats[c]->n;
                if (n) {
                    if (maps[c])
                        LAUNCH(k_remap_values, (n + 255) / 256, 256, 0, (const int32_t*)cats[c]->values_buf->ptr, (const int32_t*)maps[c]->ptr, n,
                               (int32_t*)out->values_buf->ptr + pos);
                    else
                        CUSTR_CUDA(cudaMemcpyAsync((int32_t*)out->values_buf->ptr + pos, cats[c]->values_buf->ptr, sizeof(int32_t) * (size_t)n,
                                                   cudaMemcpyDeviceToDevice, g_stream));
                }
                pos += n;
            }
            out->has_null_key = out->keys->nulls > 0;
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return guard.release();
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}
}  // extern "C"

// ---- key-set algebra on an existing category: add_keys_and_remap NVCategory.cu:1375-1480,
//      remove_keys_and_remap :1482-1565, remove_unused_keys_and_remap :1567-1706, set_keys_and_remap :1708-1820, gather :1142,
//      gather_and_remap :1084, gather_strings :1011.  Key sets are dictionary-sized: their bookkeeping runs on the host, only the
//      remap of the n values is a device pass.
namespace {
using CatPtr = std::unique_ptr<custr_category, void (*)(custr_category*)>;
CatPtr own(custr_category* c) { return CatPtr(c, [](custr_category* x) { custr_category_free(x); }); }
std::vector<int32_t> to_host32(const void* dev, size_t n)
{
    std::vector<int32_t> h(n);
    if (n) {
        CUSTR_CUDA(cudaMemcpyAsync(h.data(), dev, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, g_stream));
        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    }
    return h;
}
// index of every row of `keys` in the sorted distinct column `sorted` (-1: absent)
std::vector<int32_t> lookup_host(const custr_column* keys, const custr_column* sorted)
{
    const int32_t k = keys->n;
    if (!k) return {};
    Scratch<int32_t> map((size_t)k);
    LAUNCH(k_lookup_keys, (k + 255) / 256, 256, 0, view_of(keys), view_of(sorted), map.get());
    return to_host32(map.get(), (size_t)k);
}
custr_category* with_keys_and_map(const custr_category* cat, custr_column* new_keys, const std::vector<int32_t>& map)
{
    custr_category* out = new custr_category;
    out->keys = new_keys;
    out->n = cat->n;
    out->has_null_key = new_keys && new_keys->nulls > 0;
    out->values_buf = dev_alloc(sizeof(int32_t) * (size_t)(cat->n ? cat->n : 1));
    if (cat->n) {
        BufPtr d_map = upload(map.empty() ? (const void*)&cat->n : (const void*)map.data(), sizeof(int32_t) * (map.empty() ? 1 : map.size()));
        LAUNCH(k_remap_values, (cat->n + 255) / 256, 256, 0, (const int32_t*)cat->values_buf->ptr, (const int32_t*)d_map->ptr, cat->n,
               (int32_t*)out->values_buf->ptr);
        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    }
    return out;
}
// keys of `cat` in sorted order minus the flagged ones -> new category (values of dropped keys become -1)
custr_category* drop_keys(const custr_category* cat, const custr_category* sorted_self, const std::vector<uint8_t>& drop_rank)
{
    const int32_t k = cat->keys->n;
    const std::vector<int32_t> rank = to_host32(sorted_self->values_buf->ptr, (size_t)k);  // old key index -> sorted rank
    std::vector<int32_t> keep, new_of_rank(drop_rank.size(), -1);
    for (size_t r = 0; r < drop_rank.size(); ++r)
        if (!drop_rank[r]) { new_of_rank[r] = (int32_t)keep.size(); keep.push_back((int32_t)r); }
    std::vector<int32_t> map((size_t)k);
    for (int32_t i = 0; i < k; ++i) map[(size_t)i] = new_of_rank[(size_t)rank[(size_t)i]];
    custr_column* nk = custr_gather(sorted_self->keys, keep.data(), (int32_t)keep.size(), 0);
    if (!nk) throw ArgError{CUSTR_ERR_INVALID};
    return with_keys_and_map(cat, nk, map);
}
}  // namespace

extern "C" {
// op: 0 add_keys_and_remap, 1 remove_keys_and_remap, 2 set_keys_and_remap, 3 remove_unused_keys_and_remap (strs unused)
custr_category* custr_category_keys_op(const custr_category* cat, const custr_column* strs, int op)
{
    return guarded(
        [&]() -> custr_category* {
            if (!cat || !cat->keys || (op != 3 && !strs) || op < 0 || op > 3) throw ArgError{fail(CUSTR_ERR_ARG, "category keys op: bad argument")};
            const custr_column* keys = cat->keys;
            if (op == 0) {
                const custr_column* both[2] = {keys, strs};
                std::unique_ptr<custr_column> all(concat_columns(both, 2));
                CatPtr uni = own(build_category(all.get()));
                const std::vector<int32_t> map = lookup_host(keys, uni->keys);
                custr_column* nk = uni->keys;
                uni->keys = nullptr;
                return with_keys_and_map(cat, nk, map);
            }
            if (op == 2) {
                CatPtr uni = own(build_category(strs));
                const std::vector<int32_t> map = lookup_host(keys, uni->keys);
                custr_column* nk = uni->keys;
                uni->keys = nullptr;
                return with_keys_and_map(cat, nk, map);
            }
            CatPtr self = own(build_category(keys));  // sorted view of the own keys (they may be unsorted after merge_category)
            std::vector<uint8_t> drop((size_t)self->keys->n, 0);
            if (op == 1) {
                CatPtr rem = own(build_category(strs));
                const std::vector<int32_t> hit = lookup_host(self->keys, rem->keys);
                for (size_t r = 0; r < drop.size(); ++r) drop[r] = hit[r] >= 0;
            } else {
                const int32_t k = keys->n;
                Scratch<uint8_t> used((size_t)(k ? k : 1));
                Scratch<int> bad(1);
                CUSTR_CUDA(cudaMemsetAsync(used.get(), 0, (size_t)(k ? k : 1), g_stream));
                CUSTR_CUDA(cudaMemsetAsync(bad.get(), 0, sizeof(int), g_stream));
                if (cat->n) LAUNCH(k_mark_used, (cat->n + 255) / 256, 256, 0, (const int32_t*)cat->values_buf->ptr, cat->n, k, -1, used.get(), bad.get());
                std::vector<uint8_t> h_used((size_t)k);
                if (k) CUSTR_CUDA(cudaMemcpyAsync(h_used.data(), used.get(), (size_t)k, cudaMemcpyDeviceToHost, g_stream));
                const std::vector<int32_t> rank = to_host32(self->values_buf->ptr, (size_t)k);
                std::fill(drop.begin(), drop.end(), 1);
                for (int32_t i = 0; i < k; ++i)
                    if (h_used[(size_t)i]) drop[(size_t)rank[(size_t)i]] = 0;
            }
            return drop_keys(cat, self.get(), drop);
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}

// remap == 0: NVCategory::gather — same keys, values = pos (each in [-1, keys)); remap != 0: gather_and_remap — keys = the keys
// that pos uses (sorted), values = pos remapped onto them (each in [0, keys)).  Out-of-range positions: CUSTR_ERR_INVALID
// (the reference throws std::out_of_range).
custr_category* custr_category_gather(const custr_category* cat, const int32_t* pos, int32_t count, int devmem, int remap)
{
    return guarded(
        [&]() -> custr_category* {
            if (!cat || !cat->keys || count < 0 || (count && !pos)) throw ArgError{fail(CUSTR_ERR_ARG, "category gather: bad argument")};
            const int32_t k = cat->keys->n;
            BufPtr vals = dev_alloc(sizeof(int32_t) * (size_t)(count ? count : 1));
            if (count)
                CUSTR_CUDA(cudaMemcpyAsync(vals->ptr, pos, sizeof(int32_t) * (size_t)count, devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g_stream));
            Scratch<uint8_t> used((size_t)(k ? k : 1));
            Scratch<int> bad(1);
            CUSTR_CUDA(cudaMemsetAsync(used.get(), 0, (size_t)(k ? k : 1), g_stream));
            CUSTR_CUDA(cudaMemsetAsync(bad.get(), 0, sizeof(int), g_stream));
            // (the reference's gather documents -1 as allowed but compares against an unsigned key count, NVCategory.cu:1156-1163: -1 is
            // rejected there too — same here)
            if (count) LAUNCH(k_mark_used, (count + 255) / 256, 256, 0, (const int32_t*)vals->ptr, count, k, 0, used.get(), bad.get());
            int h_bad = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h_bad, bad.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
            std::vector<uint8_t> h_used((size_t)k);
            if (k) CUSTR_CUDA(cudaMemcpyAsync(h_used.data(), used.get(), (size_t)k, cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            if (h_bad) throw ArgError{fail(CUSTR_ERR_INVALID, "category gather: position out of range")};
            custr_category tmp;  // the gathered values over the old keys
            tmp.keys = cat->keys;
            tmp.values_buf = vals;
            tmp.n = count;
            struct Unhook { custr_category& t; ~Unhook() { t.keys = nullptr; } } unhook{tmp};
            if (!remap) {
                custr_category* out = new custr_category;
                out->keys = custr_slice_rows(cat->keys, 0, k);
                out->values_buf = vals;
                out->n = count;
                out->has_null_key = cat->has_null_key;
                return out;
            }
            CatPtr self = own(build_category(cat->keys));
            const std::vector<int32_t> rank = to_host32(self->values_buf->ptr, (size_t)k);
            std::vector<uint8_t> drop((size_t)self->keys->n, 1);
            for (int32_t i = 0; i < k; ++i)
                if (h_used[(size_t)i]) drop[(size_t)rank[(size_t)i]] = 0;
            return drop_keys(&tmp, self.get(), drop);
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}

// NVCategory::gather_strings: the key strings at pos[i] (each in [0, keys))
custr_column* custr_category_gather_strings(const custr_category* cat, const int32_t* pos, int32_t count, int devmem)
{
    return guarded(
        [&]() -> custr_column* {
            if (!cat || !cat->keys || count < 0 || (count && !pos)) throw ArgError{fail(CUSTR_ERR_ARG, "gather_strings: bad argument")};
            BufPtr vals = dev_alloc(sizeof(int32_t) * (size_t)(count ? count : 1));
            if (count)
                CUSTR_CUDA(cudaMemcpyAsync(vals->ptr, pos, sizeof(int32_t) * (size_t)count, devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g_stream));
            Scratch<int> bad(1);
            CUSTR_CUDA(cudaMemsetAsync(bad.get(), 0, sizeof(int), g_stream));
            if (count) LAUNCH(k_mark_used, (count + 255) / 256, 256, 0, (const int32_t*)vals->ptr, count, cat->keys->n, 0, (uint8_t*)nullptr, bad.get());
            int h_bad = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h_bad, bad.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            if (h_bad) throw ArgError{fail(CUSTR_ERR_INVALID, "gather_strings: position out of range")};
            custr_column* out = custr_gather(cat->keys, (const int32_t*)vals->ptr, count, 1);
            if (!out) throw ArgError{CUSTR_ERR_INVALID};
            return out;
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}
}  // extern "C"

extern "C" {

// ---- multi-GPU: row-sharded dictionary build with the key exchange over NCCL ------------------------------
// The only collective of the whole path.  NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is
// the library torch already loaded, otherwise the system one), so libcustr.so itself has no hard dependency on it.
}  // extern "C"
#include <dlfcn.h>
#if __has_include(<nccl.h>)
#include <nccl.h>
#else
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
enum { ncclUint8 = 1, ncclInt32 = 2, ncclInt64 = 4 };
#endif
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
const NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
    });
    return api;
}
void nccl_check(ncclResult_t r, const char* what)
{
    if (r != 0) {
        const NcclApi& a = nccl();
        throw custr::ArgError{custr::fail(CUSTR_ERR_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"))};
    }
}
}  // namespace
struct custr_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};
extern "C" {

int custr_comm_unique_id(void* id128)
{
    return guarded(
        [&]() -> int {
            if (!id128) return fail(CUSTR_ERR_ARG, "comm_unique_id: null argument");
            if (!nccl().ok) return fail(CUSTR_ERR_CUDA, "NCCL library (libnccl.so.2) not found");
            ncclUniqueId id;
            nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
            memcpy(id128, &id, sizeof(id));
            return CUSTR_OK;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_comm* custr_comm_create(int rank, int world, const void* id128)
{
    return guarded(
        [&]() -> custr_comm* {
            if (!id128 || world < 1 || rank < 0 || rank >= world) throw ArgError{fail(CUSTR_ERR_ARG, "comm_create: bad argument")};
            if (!nccl().ok) throw ArgError{fail(CUSTR_ERR_CUDA, "NCCL library (libnccl.so.2) not found")};
            ncclUniqueId id;
            memcpy(&id, id128, sizeof(id));
            std::unique_ptr<custr_comm> c(new custr_comm);
            c->rank = rank;
            c->world = world;
            nccl_check(nccl().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
            return c.release();
        },
        (custr_comm*)nullptr, (custr_comm*)nullptr);
}

void custr_comm_destroy(custr_comm* c)
{
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

// Every rank: dictionary of its own rows -> all-gather of (key count, key bytes), then of the max-padded key lengths and key
// bytes (two ncclAllGather over NVLink; K keys x ~16 B per rank: latency-bound) -> union column of every rank's keys -> global
// sorted key set + remap of the local values (create_from_categories math, NVCategory.cu:430-514).  Every rank returns the
// same keys; values index into them.  timings_ms (optional, 3 floats): local build, key exchange, union + remap.
custr_category* custr_category_create_sharded(custr_comm* comm, const custr_column* col, float* timings_ms)
{
    return guarded(
        [&]() -> custr_category* {
            if (!comm || !col) throw ArgError{fail(CUSTR_ERR_ARG, "category_create_sharded: null argument")};
            cudaEvent_t ev[4];
            for (auto& e : ev) CUSTR_CUDA(cudaEventCreate(&e));
            struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]); } } evg{ev};
            CUSTR_CUDA(cudaEventRecord(ev[0], g_stream));
            std::unique_ptr<custr_category, void (*)(custr_category*)> local(build_category(col), [](custr_category* c) { custr_category_free(c); });
            CUSTR_CUDA(cudaEventRecord(ev[1], g_stream));
            const int world = comm->world;
            const custr_column* lk = local->keys;
            const int32_t k = lk->n;
            // (1) sizes
            long long h_meta[2] = {k, lk->nbytes};
            Scratch<long long> d_meta(2), d_metas(2 * (size_t)world);
            CUSTR_CUDA(cudaMemcpyAsync(d_meta.get(), h_meta, sizeof(h_meta), cudaMemcpyHostToDevice, g_stream));
            nccl_check(nccl().AllGather(d_meta.get(), d_metas.get(), 2, ncclInt64, comm->comm, g_stream), "ncclAllGather(sizes)");
            std::vector<long long> metas(2 * (size_t)world);
            CUSTR_CUDA(cudaMemcpyAsync(metas.data(), d_metas.get(), sizeof(long long) * metas.size(), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            long long max_k = 1, max_b = 1, tot_k = 0, tot_b = 0;
            for (int r = 0; r < world; ++r) {
                max_k = std::max(max_k, metas[2 * r]);
                max_b = std::max(max_b, metas[2 * r + 1]);
                tot_k += metas[2 * r];
                tot_b += metas[2 * r + 1];
            }
            if (tot_b > 0x7fffffffLL || tot_k > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "category_create_sharded: key union exceeds 2 GiB")};
            // (2) payload: lengths (-1 = the null key) and key bytes, max-padded
            Scratch<int32_t> d_len((size_t)max_k), d_lens((size_t)max_k * world);
            Scratch<uint8_t> d_chr((size_t)max_b), d_chrs((size_t)max_b * world);
            if (k) LAUNCH(k_key_lengths, (k + 255) / 256, 256, 0, view_of(lk), (int32_t*)d_len.get());
            if (lk->nbytes) CUSTR_CUDA(cudaMemcpyAsync(d_chr.get(), lk->chars + lk->first_off, (size_t)lk->nbytes, cudaMemcpyDeviceToDevice, g_stream));
            nccl_check(nccl().AllGather(d_len.get(), d_lens.get(), (size_t)max_k, ncclInt32, comm->comm, g_stream), "ncclAllGather(key lengths)");
            nccl_check(nccl().AllGather(d_chr.get(), d_chrs.get(), (size_t)max_b, ncclUint8, comm->comm, g_stream), "ncclAllGather(key bytes)");
            std::vector<int32_t> lens((size_t)max_k * world);
            std::vector<uint8_t> chrs((size_t)max_b * world);
            CUSTR_CUDA(cudaMemcpyAsync(lens.data(), d_lens.get(), sizeof(int32_t) * lens.size(), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaMemcpyAsync(chrs.data(), d_chrs.get(), chrs.size(), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaEventRecord(ev[2], g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            // (3) union column of every rank's keys in rank order (a few thousand dictionary entries: host bookkeeping)
            std::vector<char> all_chars((size_t)(tot_b ? tot_b : 1));
            std::vector<int32_t> all_off((size_t)tot_k + 1, 0);
            std::vector<uint8_t> all_valid((size_t)(tot_k + 7) / 8 + 1, 0);
            size_t row = 0, pos = 0;
            int nulls = 0;
            for (int r = 0; r < world; ++r) {
                size_t src = 0;
                for (long long i = 0; i < metas[2 * r]; ++i, ++row) {
                    const int32_t len = lens[(size_t)r * max_k + i];
                    if (len < 0) ++nulls;
                    else {
                        all_valid[row >> 3] |= (uint8_t)(1u << (row & 7));
                        memcpy(all_chars.data() + pos, chrs.data() + (size_t)r * max_b + src, (size_t)len);
                        pos += (size_t)len;
                        src += (size_t)len;
                    }
                    all_off[row + 1] = (int32_t)pos;
                }
            }
            std::unique_ptr<custr_column> all_keys(custr_create_from_offsets(all_chars.data(), (int32_t)tot_k, all_off.data(), all_valid.data(), nulls, 0));
            if (!all_keys) throw ArgError{CUSTR_ERR_INVALID};
            custr_category* out = custr_category_remap_to_union(local.get(), all_keys.get());
            if (!out) throw ArgError{CUSTR_ERR_INVALID};
            CUSTR_CUDA(cudaEventRecord(ev[3], g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            if (timings_ms)
                for (int i = 0; i < 3; ++i) CUSTR_CUDA(cudaEventElapsedTime(&timings_ms[i], ev[i], ev[i + 1]));
            return out;
        },
        (custr_category*)nullptr, (custr_category*)nullptr);
}

}  // extern "C"
