// Column life-cycle, import/export and the cheap per-row attributes.
// Replaces (reference, rapidsai/custrings): NVStrings::create_from_offsets NVStrings.cu:109-119 ->
// NVStringsImpl.cu:328-444, create_from_array NVStrings.cu:74-86, create_offsets :402-482,
// set_null_bitarray :493-544, byte_count/len attrs.cu:32-110, hash convert.cu:34-63 + custring.inl:164-232.
// The reference rebuilds a pointer-per-string object heap on import (3 passes + memmove per row); here the
// Arrow triple IS the storage, so import is two memcpys and export is one.
#include "common.cuh"
#include <vector>
#include <string>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "device_utils.cuh"
#include <cub/cub.cuh>
#include <thrust/sort.h>
#include <thrust/sequence.h>
#include <thrust/execution_policy.h>
#include <thrust/system/cuda/execution_policy.h>

namespace custr {

thread_local std::string g_error;
thread_local cudaStream_t g_stream = 0;
std::atomic<long long> g_launches{0};

// Big-block cache in front of the driver's pool.  The pool keeps freed memory (release threshold = max), but it hands a freed
// 1 GB block to the next 400 MB request and then has to stitch the 1 GB request that follows together from fragments — a
// remap that idles the stream for 2-4 ms (CUSTR_TRACE=2 showed replace_re 2.2 ms in a fresh process and 4.4-4.9 ms after other
// calls had populated the pool, all of it inside cudaMallocAsync).  Calls on the same column ask for the same sizes again and
// again (result chars, span streams, offsets), so blocks >= 32 MiB are kept here per (device, stream) and reused for requests of
// the same size or up to 1/8 smaller; at most 8 blocks / 6 GiB, oldest evicted to the pool.
namespace {
struct BigBlock { void* ptr; size_t cap; int dev; cudaStream_t stream; };
struct BigCache {
    std::vector<BigBlock> blocks;  // oldest first
    size_t total = 0;
    ~BigCache() { for (auto& b : blocks) cudaFreeAsync(b.ptr, b.stream); }
};
thread_local BigCache g_big;
constexpr size_t BIG_MIN = 32ull << 20, BIG_TOTAL = 6ull << 30;
constexpr size_t BIG_COUNT = 8;
}  // namespace

// custr_release_cached_memory: hand this thread's cached big blocks and the pool's unused memory back to the driver
void release_cached_memory()
{
    for (auto& b : g_big.blocks) cudaFreeAsync(b.ptr, b.stream);
    g_big.blocks.clear();
    g_big.total = 0;
    cudaStreamSynchronize(g_stream);
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

DeviceBuf::DeviceBuf(size_t n) : bytes(n), cap(n)
{
    static std::atomic<unsigned long long> pools_ready{0};  // one bit per device ordinal
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(pools_ready.load(std::memory_order_relaxed) & bit)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;  // keep freed memory cached in the pool instead of returning it to the driver
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        pools_ready.fetch_or(bit, std::memory_order_relaxed);
    }
    if (n >= BIG_MIN) {
        int best = -1;
        for (int i = 0; i < (int)g_big.blocks.size(); ++i) {
            const BigBlock& b = g_big.blocks[i];
            if (b.dev == dev && b.stream == g_stream && b.cap >= n && b.cap - n <= n / 8 && (best < 0 || b.cap < g_big.blocks[best].cap)) best = i;
        }
        if (best >= 0) {
            ptr = g_big.blocks[best].ptr;
            cap = g_big.blocks[best].cap;
            g_big.total -= cap;
            g_big.blocks.erase(g_big.blocks.begin() + best);
            return;
        }
    }
    cudaError_t e = cudaMallocAsync(&ptr, n, g_stream);
    if (e != cudaSuccess && !g_big.blocks.empty()) {  // give the cached blocks back and retry
        for (auto& b : g_big.blocks) cudaFreeAsync(b.ptr, b.stream);
        g_big.blocks.clear();
        g_big.total = 0;
        cudaGetLastError();
        e = cudaMallocAsync(&ptr, n, g_stream);
    }
    if (e != cudaSuccess) {
        g_error = std::string("device allocation of ") + std::to_string(n) + " bytes failed: " + cudaGetErrorString(e);
        cudaGetLastError();
        throw CudaError{e};
    }
}
DeviceBuf::~DeviceBuf()
{
    if (!ptr) return;
    int dev = 0;
    if (cap >= BIG_MIN && cap <= BIG_TOTAL / 2 && cudaGetDevice(&dev) == cudaSuccess) {
        g_big.blocks.push_back(BigBlock{ptr, cap, dev, g_stream});
        g_big.total += cap;
        while (g_big.blocks.size() > BIG_COUNT || g_big.total > BIG_TOTAL) {
            cudaFreeAsync(g_big.blocks.front().ptr, g_big.blocks.front().stream);
            g_big.total -= g_big.blocks.front().cap;
            g_big.blocks.erase(g_big.blocks.begin());
        }
        return;
    }
    cudaFreeAsync(ptr, g_stream);
}

// CUSTR_TRACE=1: synchronise at every point and print host time deltas.  CUSTR_TRACE=2: record a CUDA event at every point (no
// synchronisation, the call runs as it does untraced) and print the device time between consecutive points when a point
// whose name ends in '.' is reached.
void trace_point(const char* what)
{
    static const int mode = getenv("CUSTR_TRACE") ? atoi(getenv("CUSTR_TRACE")) : 0;
    if (!mode) return;
    if (mode == 1) {
        static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
        cudaStreamSynchronize(g_stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[custr trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
        return;
    }
    static thread_local std::vector<std::pair<std::string, cudaEvent_t>> evs;
    static thread_local std::chrono::steady_clock::time_point h0;
    if (evs.empty()) h0 = std::chrono::steady_clock::now();
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, g_stream);
    evs.emplace_back(what, e);
    const size_t len = strlen(what);
    if (len && what[len - 1] == '.') {
        cudaEventSynchronize(e);
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        for (size_t i = 1; i < evs.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, evs[i - 1].second, evs[i].second);
            fprintf(stderr, "[custr trace/dev] %-28s %8.3f ms\n", evs[i].first.c_str(), ms);
        }
        fprintf(stderr, "[custr trace/dev] host wall first..last point %8.3f ms\n", host_ms);
        {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaMemPool_t pool;
            uint64_t reserved = 0, used = 0, thr = 0;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            fprintf(stderr, "[custr trace/dev] pool reserved %.1f MB, used %.1f MB, release threshold %s\n", reserved / 1e6, used / 1e6,
                    thr == UINT64_MAX ? "max" : std::to_string(thr).c_str());
        }
        for (auto& p : evs) cudaEventDestroy(p.second);
        evs.clear();
    }
}
int num_sms()
{
    static std::atomic<int> sms[64];  // per device (custr_set_device may switch devices inside one process)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int v = sms[dev].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

BufPtr upload(const void* host, size_t bytes)
{
    BufPtr b = dev_alloc(bytes);
    if (bytes) CUSTR_CUDA(cudaMemcpyAsync(b->ptr, host, bytes, cudaMemcpyHostToDevice, g_stream));
    return b;
}

// ------------------------------------------------------------------------------------------------ kernels
__global__ void k_rebase_offsets(int32_t* offsets, int32_t n1, int32_t base)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) offsets[i] -= base;
}

// lengths with nulls forced to zero; flags a null row that carries bytes
__global__ void k_valid_lengths(ColView c, int32_t* lengths, int* dirty)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int len = c.offsets[i + 1] - c.offsets[i];
    if (!c.valid(i)) {
        if (len) *dirty = 1;
        len = 0;
    }
    lengths[i] = len;
}

// one warp per row copy (rows are ~100 B; coalesced 32-lane byte copy)
__global__ void k_gather_rows(const char* __restrict__ src, const int32_t* __restrict__ src_off,
                              const int32_t* __restrict__ rows, char* __restrict__ dst,
                              const int32_t* __restrict__ dst_off, int32_t n)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int r = rows ? rows[warp] : warp;
    int d0 = dst_off[warp], len = dst_off[warp + 1] - d0;
    if (len <= 0) return;
    const char* s = src + src_off[r];
    for (int j = lane; j < len; j += 32) dst[d0 + j] = s[j];
}

__global__ void k_pack_bits(const uint8_t* __restrict__ flags, uint8_t* __restrict__ bits, int32_t n)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int nb = (n + 7) >> 3;
    if (b >= nb) return;
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) {
        int i = b * 8 + k;
        if (i < n && flags[i]) v |= 1u << k;
    }
    bits[b] = (uint8_t)v;
}

// number of set bits among the n validity bits that start at bit vbit0: one byte per thread and step, one atomic per warp
__global__ void k_count_valid(const uint8_t* __restrict__ bits, int32_t vbit0, int32_t n, unsigned long long* total)
{
    const long long first = vbit0, last = (long long)vbit0 + n;  // bit range [first, last)
    const long long byte0 = first >> 3, byte1 = (last + 7) >> 3;
    unsigned cnt = 0;
    for (long long b = byte0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; b < byte1; b += (long long)gridDim.x * blockDim.x) {
        unsigned v = bits[b];
        const long long lo = b << 3;
        if (lo < first) v &= 0xFFu << (first - lo);
        if (lo + 8 > last) v &= 0xFFu >> (lo + 8 - last);
        cnt += __popc(v);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

// validity re-aligned to bit 0; rows past n are 0; optionally clears empties
__global__ void k_export_validity(ColView c, uint8_t* __restrict__ out, int empty_is_null, unsigned long long* cleared)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int nb = (c.n + 7) >> 3;
    if (b >= nb) return;
    unsigned v = 0;
    int zeros = 0;
    for (int k = 0; k < 8; ++k) {
        int i = b * 8 + k;
        if (i >= c.n) break;
        bool ok = c.valid(i);
        if (ok && empty_is_null && c.offsets[i + 1] == c.offsets[i]) ok = false;
        if (ok) v |= 1u << k; else ++zeros;
    }
    out[b] = (uint8_t)v;
    if (zeros && cleared) atomicAdd(cleared, (unsigned long long)zeros);
}

__global__ void k_byte_count(ColView c, int32_t* __restrict__ lengths, unsigned long long* total)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (i < c.n) {
        bool ok = c.valid(i);
        len = ok ? c.offsets[i + 1] - c.offsets[i] : 0;
        if (lengths) lengths[i] = ok ? len : -1;
    }
    for (int o = 16; o; o >>= 1) len += __shfl_down_sync(0xffffffffu, len, o);
    if ((threadIdx.x & 31) == 0 && len) atomicAdd(total, (unsigned long long)len);
}

__global__ void k_char_len(ColView c, int32_t* __restrict__ lengths)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    if (!c.valid(i)) { lengths[i] = -1; return; }
    int b = c.offsets[i], e = c.offsets[i + 1], cnt = 0;
    for (int j = b; j < e; ++j) cnt += ((uint8_t)c.chars[j] & 0xC0) != 0x80;
    lengths[i] = cnt;
}

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// MurmurHash3_x86_32, seed 31 (published algorithm; the reference's variant is custring.inl:164-232)
__global__ void k_murmur3(ColView c, uint32_t* __restrict__ out, unsigned long long* nonzero)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h = 0;
    if (i < c.n && c.valid(i)) {
        const uint8_t* p = (const uint8_t*)c.chars + c.offsets[i];
        int len = c.offsets[i + 1] - c.offsets[i];
        h = 31u;
        int nblocks = len >> 2;
        for (int b = 0; b < nblocks; ++b) {
            uint32_t k = p[4 * b] | (p[4 * b + 1] << 8) | (p[4 * b + 2] << 16) | ((uint32_t)p[4 * b + 3] << 24);
            k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u;
            h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
        }
        const uint8_t* t = p + 4 * nblocks;
        uint32_t k = 0;
        int rem = len & 3;
        if (rem == 3) k ^= t[2] << 16;
        if (rem >= 2) k ^= t[1] << 8;
        if (rem >= 1) { k ^= t[0]; k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u; h ^= k; }
        h ^= (uint32_t)len;
        h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    }
    if (i < c.n) out[i] = h;
    unsigned m = __ballot_sync(0xffffffffu, h != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nonzero, (unsigned long long)__popc(m));
}

__global__ void k_gather_lengths(ColView c, const int32_t* __restrict__ idx, int32_t m, int32_t* __restrict__ lengths,
                                 uint8_t* __restrict__ valid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int r = idx[i];
    bool ok = r >= 0 && r < c.n && c.valid(r);
    lengths[i] = ok ? c.offsets[r + 1] - c.offsets[r] : 0;
    valid[i] = ok;
}

__global__ void k_clamp_rows(int32_t* idx, int32_t m, int32_t n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && (idx[i] < 0 || idx[i] >= n)) idx[i] = 0;
}

// ---- create_from_index (NVStrings.cu:88-107, NVStringsImpl.cu:209-325): (device pointer, byte length) pairs -> column
struct IndexPair { const char* ptr; size_t len; };  // layout of std::pair<const char*, size_t> / thrust::pair
__global__ void k_index_lengths(const IndexPair* __restrict__ pairs, int32_t n, const int32_t* __restrict__ order, int32_t* __restrict__ lengths,
                                uint8_t* __restrict__ valid, int* __restrict__ too_long)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const IndexPair p = pairs[order ? order[i] : i];
    if (p.ptr && p.len > 0x7fffffffULL) *too_long = 1;
    lengths[i] = p.ptr ? (int32_t)p.len : 0;
    valid[i] = p.ptr != nullptr;
}
// one warp per 32 rows: short rows are copied by their own lane, long rows by the whole warp (coalesced)
__global__ void k_index_copy(const IndexPair* __restrict__ pairs, int32_t n, const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
                             char* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const char* src = nullptr;
    int len = 0, dst = 0;
    if (i < n) {
        const IndexPair p = pairs[order ? order[i] : i];
        src = p.ptr;
        len = p.ptr ? (int)p.len : 0;
        dst = offsets[i];
    }
    unsigned big = __ballot_sync(0xffffffffu, len > 64);
    if (len <= 64)
        for (int k = 0; k < len; ++k) out[dst + k] = src[k];
    while (big) {
        const int l = __ffs(big) - 1;
        big &= big - 1;
        const char* s = (const char*)__shfl_sync(0xffffffffu, (unsigned long long)src, l);
        const int n_ = __shfl_sync(0xffffffffu, len, l), d = __shfl_sync(0xffffffffu, dst, l);
        for (int k = lane; k < n_; k += 32) out[d + k] = s[k];
    }
}
// sort order for create_from_index's `stype` (NVStringsImpl.cu:255-268): null < non-null, then by length and / or bytes
struct IndexLess {
    const IndexPair* pairs;
    int stype;
    __device__ bool operator()(int32_t a, int32_t b) const
    {
        const IndexPair l = pairs[a], r = pairs[b];
        if (!l.ptr || !r.ptr) return r.ptr != nullptr;
        int diff = 0;
        if (stype & 1) diff = (int)(unsigned int)(l.len - r.len);
        if (diff == 0 && (stype & 2)) {
            const size_t m = l.len < r.len ? l.len : r.len;
            for (size_t k = 0; k < m && diff == 0; ++k) diff = (int)(uint8_t)l.ptr[k] - (int)(uint8_t)r.ptr[k];
            if (diff == 0) diff = l.len < r.len ? -1 : (l.len > r.len ? 1 : 0);
        }
        return diff < 0;
    }
};

// ------------------------------------------------------------------------------------------------ helpers
static inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

struct WidenLen {
    __host__ __device__ __forceinline__ long long operator()(int32_t v) const { return (long long)v; }
};

int64_t scan_lengths_to_offsets(const int32_t* lengths, int32_t* offsets, int32_t n)
{
    // The total is accumulated in 64 bits FIRST (a column that grows past 2 GiB must be rejected, not wrapped: the int32
    // scan below would otherwise hand the write pass offsets that alias), then offsets[0..n-1] = exclusive scan,
    // offsets[n] = total.  lengths must have n+1 readable entries (callers allocate n+1 and zero the last).
    Scratch<long long> wide(1);
    cub::TransformInputIterator<long long, WidenLen, const int32_t*> it(lengths, WidenLen{});
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceReduce::Sum(nullptr, tmp_bytes, it, wide.get(), n, g_stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, lengths, offsets, n + 1, g_stream);
    BufPtr tmp = dev_alloc(tmp_bytes > tmp2 ? tmp_bytes : tmp2);
    CUSTR_CUDA(cub::DeviceReduce::Sum(tmp->ptr, tmp_bytes, it, wide.get(), n, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    long long total = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&total, wide.get(), sizeof(total), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    if (total > 0x7fffffffLL) throw ArgError{fail(CUSTR_ERR_INVALID, "result exceeds 2 GiB of chars / 2^31 entries (int32 offsets)")};
    CUSTR_CUDA(cub::DeviceScan::ExclusiveSum(tmp->ptr, tmp2, lengths, offsets, n + 1, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return total;
}

void pack_bits(const uint8_t* flags, uint8_t* bits, int32_t n)
{
    if (n > 0) LAUNCH(k_pack_bits, blocks_for((n + 7) / 8, 256), 256, 0, flags, bits, n);
}

int32_t count_zero_bits(const uint8_t* bits, int32_t vbit0, int32_t n)
{
    if (!bits || n == 0) return 0;
    Scratch<unsigned long long> total(1);
    CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
    LAUNCH(k_count_valid, blocks_for((n + 7) / 8 + 1, 256 * 8), 256, 0, bits, vbit0, n, total.get());
    unsigned long long h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, total.get(), 8, cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return n - (int32_t)h;
}

custr_column* make_column(BufPtr chars, BufPtr offsets, BufPtr validity, int32_t n, int32_t nulls, int64_t nbytes)
{
    custr_column* c = new custr_column;
    c->chars_buf = chars;
    c->offsets_buf = offsets;
    c->validity_buf = (nulls > 0) ? validity : nullptr;
    c->chars = chars ? (const char*)chars->ptr : nullptr;
    c->offsets = (const int32_t*)offsets->ptr;
    c->validity = c->validity_buf ? (const uint8_t*)c->validity_buf->ptr : nullptr;
    c->n = n;
    c->nulls = nulls;
    c->nbytes = nbytes;
    c->first_off = 0;
    return c;
}

std::vector<custr_column*> assemble_columns(int32_t n, int ncols, int32_t* lens, const uint8_t* valid,
                                            void (*copy)(const ColumnOut* d_outs, void* ctx), void* ctx)
{
    std::vector<BufPtr> offs(ncols), chars(ncols);
    std::vector<int64_t> totals(ncols);
    std::vector<ColumnOut> outs(ncols);
    for (int c = 0; c < ncols; ++c) {
        CUSTR_CUDA(cudaMemsetAsync(lens + (size_t)c * (n + 1) + n, 0, sizeof(int32_t), g_stream));
        offs[c] = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
        totals[c] = scan_lengths_to_offsets(lens + (size_t)c * (n + 1), (int32_t*)offs[c]->ptr, n);
        chars[c] = dev_alloc((size_t)totals[c]);
        outs[c] = ColumnOut{(char*)chars[c]->ptr, (const int32_t*)offs[c]->ptr};
    }
    BufPtr d_outs = upload(outs.data(), sizeof(ColumnOut) * (size_t)ncols);
    copy((const ColumnOut*)d_outs->ptr, ctx);
    std::vector<custr_column*> result;
    for (int c = 0; c < ncols; ++c) {
        BufPtr bits = dev_alloc((n + 7) / 8);
        pack_bits(valid + (size_t)c * n, (uint8_t*)bits->ptr, n);
        int32_t nulls = count_zero_bits((const uint8_t*)bits->ptr, 0, n);
        result.push_back(make_column(chars[c], offs[c], bits, n, nulls, totals[c]));
    }
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return result;
}

custr_column* all_null_column(int32_t n)
{
    BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
    CUSTR_CUDA(cudaMemsetAsync(off->ptr, 0, sizeof(int32_t) * (size_t)(n + 1), g_stream));
    BufPtr val = dev_alloc((n + 7) / 8);
    CUSTR_CUDA(cudaMemsetAsync(val->ptr, 0, (n + 7) / 8, g_stream));
    return make_column(dev_alloc(1), off, val, n, n, 0);
}

// compacting copy: rows (or all rows when rows==nullptr) of `src` with nulls forced empty
static custr_column* compact_copy(const custr_column* src)
{
    int32_t n = src->n;
    Scratch<int32_t> lens((size_t)n + 1);
    Scratch<int> dirty(1);
    CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
    CUSTR_CUDA(cudaMemsetAsync(dirty.get(), 0, sizeof(int), g_stream));
    LAUNCH(k_valid_lengths, blocks_for(n, 256), 256, 0, view_of(src), lens.get(), dirty.get());
    BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
    int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);
    BufPtr chars = dev_alloc((size_t)total);
    LAUNCH(k_gather_rows, blocks_for((int64_t)n * 32, 256), 256, 0, src->chars, src->offsets, (const int32_t*)nullptr,
           (char*)chars->ptr, (const int32_t*)off->ptr, n);
    BufPtr val;
    if (src->validity) {
        val = dev_alloc((n + 7) / 8);
        LAUNCH(k_export_validity, blocks_for((n + 7) / 8, 256), 256, 0, view_of(src), (uint8_t*)val->ptr, 0,
               (unsigned long long*)nullptr);
    }
    return make_column(chars, off, val, n, src->nulls, total);
}

static bool null_rows_carry_bytes(const custr_column* c)
{
    if (!c->validity || c->n == 0) return false;
    Scratch<int32_t> lens((size_t)c->n + 1);
    Scratch<int> dirty(1);
    CUSTR_CUDA(cudaMemsetAsync(dirty.get(), 0, sizeof(int), g_stream));
    LAUNCH(k_valid_lengths, blocks_for(c->n, 256), 256, 0, view_of(c), lens.get(), dirty.get());
    int h = 0;
    CUSTR_CUDA(cudaMemcpyAsync(&h, dirty.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    return h != 0;
}

static custr_column* create_from_offsets_impl(const char* chars, int32_t count, const int32_t* offsets,
                                              const uint8_t* validity, int32_t nulls, int devmem, bool adopt)
{
    if (count < 0 || (count > 0 && !offsets)) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_offsets: null offsets")};
    if (count == 0) {
        BufPtr off = dev_alloc(sizeof(int32_t));
        CUSTR_CUDA(cudaMemsetAsync(off->ptr, 0, sizeof(int32_t), g_stream));
        return make_column(dev_alloc(1), off, nullptr, 0, 0, 0);
    }
    cudaMemcpyKind kind = devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int32_t ends[2];
    if (devmem) {
        CUSTR_CUDA(cudaMemcpyAsync(&ends[0], offsets, sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream));
        CUSTR_CUDA(cudaMemcpyAsync(&ends[1], offsets + count, sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream));
        CUSTR_CUDA(cudaStreamSynchronize(g_stream));
    } else {
        ends[0] = offsets[0];
        ends[1] = offsets[count];
    }
    if (ends[1] < ends[0]) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_offsets: offsets not ascending")};
    int64_t nbytes = (int64_t)ends[1] - ends[0];
    if (nbytes > 0 && !chars) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_offsets: null chars")};
    bool use_mask = validity != nullptr && nulls != 0;

    custr_column* c = new custr_column;
    std::unique_ptr<custr_column> guard(c);
    c->n = count;
    c->nbytes = nbytes;
    if (adopt) {
        c->chars = chars;
        c->offsets = offsets;
        c->validity = use_mask ? validity : nullptr;
        c->first_off = ends[0];
    } else {
        c->offsets_buf = dev_alloc(sizeof(int32_t) * (size_t)(count + 1));
        CUSTR_CUDA(cudaMemcpyAsync(c->offsets_buf->ptr, offsets, sizeof(int32_t) * (size_t)(count + 1), kind, g_stream));
        if (ends[0] != 0)
            LAUNCH(k_rebase_offsets, blocks_for(count + 1, 256), 256, 0, (int32_t*)c->offsets_buf->ptr, count + 1, ends[0]);
        c->chars_buf = dev_alloc((size_t)nbytes);
        if (nbytes) CUSTR_CUDA(cudaMemcpyAsync(c->chars_buf->ptr, chars + ends[0], (size_t)nbytes, kind, g_stream));
        if (use_mask) {
            c->validity_buf = dev_alloc((count + 7) / 8);
            CUSTR_CUDA(cudaMemcpyAsync(c->validity_buf->ptr, validity, (count + 7) / 8, kind, g_stream));
        }
        c->chars = (const char*)c->chars_buf->ptr;
        c->offsets = (const int32_t*)c->offsets_buf->ptr;
        c->validity = use_mask ? (const uint8_t*)c->validity_buf->ptr : nullptr;
        c->first_off = 0;
    }
    if (use_mask) {
        c->nulls = count_zero_bits(c->validity, 0, count);
        if (c->nulls == 0) { c->validity = nullptr; c->validity_buf = nullptr; }
        else if (null_rows_carry_bytes(c)) {
            custr_column* fixed = compact_copy(c);
            return fixed;  // guard frees the raw view
        }
    }
    if (!devmem) CUSTR_CUDA(cudaStreamSynchronize(g_stream));  // host buffers may be reused by the caller
    return guard.release();
}

}  // namespace custr

using namespace custr;

extern "C" {

const char* custr_last_error(void) { return g_error.c_str(); }
const char* custr_version(void) { return "custrings_b200 0.1 (sm_100a)"; }

int custr_set_device(int device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(CUSTR_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return CUSTR_OK;
}
void custr_set_stream(void* s) { g_stream = (cudaStream_t)s; }
int custr_sync(void)
{
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) return fail(CUSTR_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    return CUSTR_OK;
}
long long custr_launch_count(void) { return g_launches.load(); }

custr_column* custr_create_from_offsets(const char* chars, int32_t count, const int32_t* offsets, const uint8_t* validity,
                                        int32_t nulls, int devmem)
{
    return guarded([&] { return create_from_offsets_impl(chars, count, offsets, validity, nulls, devmem, false); },
                   (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_adopt_device(const char* chars, int32_t count, const int32_t* offsets, const uint8_t* validity, int32_t nulls)
{
    return guarded([&] { return create_from_offsets_impl(chars, count, offsets, validity, nulls, 1, true); },
                   (custr_column*)nullptr, (custr_column*)nullptr);
}

// ---- CUDA IPC (reference cpp/include/ipc_transfer.h:31-100, NVStrings::create_ipc_transfer / create_from_ipc): hand a column
// to another process on the same GPU without a copy through the host.  The column's three buffers are packed into ONE plain
// cudaMalloc allocation (stream-ordered pool memory cannot be exported with cudaIpcGetMemHandle): [chars | offsets | validity]
int custr_ipc_export(const custr_column* col, custr_ipc_handle* out)
{
    return guarded(
        [&]() -> int {
            if (!col || !out) return fail(CUSTR_ERR_ARG, "ipc_export: null argument");
            memset(out, 0, sizeof(*out));
            const int32_t n = col->n;
            const size_t chars_b = ((size_t)col->nbytes + 15) & ~(size_t)15, off_b = (sizeof(int32_t) * (size_t)(n + 1) + 15) & ~(size_t)15;
            const size_t val_b = col->validity ? (size_t)(n + 7) / 8 : 0;
            char* base = nullptr;
            CUSTR_CUDA(cudaMalloc(&base, chars_b + off_b + val_b + 16));
            std::shared_ptr<void> owner(base, [](void* p) { cudaFree(p); });
            // normalised copy: offsets rebased to 0, validity re-aligned to bit 0 (the export helper does both)
            const int rc = custr_create_offsets(col, base, (int32_t*)(base + chars_b), val_b ? (uint8_t*)(base + chars_b + off_b) : nullptr, 1);
            if (rc < 0) return rc;
            const int nulls = col->nulls;
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            cudaIpcMemHandle_t h;
            CUSTR_CUDA(cudaIpcGetMemHandle(&h, base));
            static_assert(sizeof(h) <= sizeof(out->handle), "handle size");
            memcpy(out->handle, &h, sizeof(h));
            out->n = n;
            out->nulls = nulls;
            out->chars_bytes = col->nbytes;
            out->offsets_at = (int64_t)chars_b;
            out->validity_at = val_b ? (int64_t)(chars_b + off_b) : -1;
            col->ipc_owner = owner;
            return CUSTR_OK;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_ipc_import(const custr_ipc_handle* in)
{
    return guarded(
        [&]() -> custr_column* {
            if (!in) throw ArgError{fail(CUSTR_ERR_ARG, "ipc_import: null argument")};
            cudaIpcMemHandle_t h;
            memcpy(&h, in->handle, sizeof(h));
            void* base = nullptr;
            CUSTR_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            std::shared_ptr<void> owner(base, [](void* p) { cudaIpcCloseMemHandle(p); });
            const char* b = (const char*)base;
            custr_column* c = create_from_offsets_impl(b, in->n, (const int32_t*)(b + in->offsets_at),
                                                       in->validity_at >= 0 ? (const uint8_t*)(b + in->validity_at) : nullptr, in->nulls, 1, true);
            c->ipc_owner = owner;
            return c;
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_create_from_array(const char* const* strs, uint32_t count)
{
    return guarded(
        [&]() -> custr_column* {
            if (count && !strs) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_array: null array")};
            std::vector<int32_t> off(count + 1, 0);
            std::vector<uint8_t> val((count + 7) / 8, 0);
            int32_t nulls = 0;
            size_t total = 0;
            for (uint32_t i = 0; i < count; ++i) {
                if (strs[i]) { total += strlen(strs[i]); val[i >> 3] |= 1u << (i & 7); }
                else ++nulls;
                if (total > 0x7fffffffULL) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_array: more than 2 GiB of chars")};
                off[i + 1] = (int32_t)total;
            }
            std::vector<char> chars(total ? total : 1);
            for (uint32_t i = 0; i < count; ++i)
                if (strs[i]) memcpy(chars.data() + off[i], strs[i], off[i + 1] - off[i]);
            return create_from_offsets_impl(chars.data(), (int32_t)count, off.data(), val.data(), nulls, 0, false);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

// NVStrings::create_from_index (NVStrings.cu:88-107): `pairs` = count (pointer, byte length) pairs laid out like
// std::pair<const char*, size_t>, in device memory when devmem != 0 else on the host; the POINTERS always address device memory
// (the reference dereferences them in device code either way, NVStringsImpl.cu:232-238,272-279).  Null pointer = null row.
// stype: 0 none, 1 length, 2 name, 3 both (NVStrings::sorttype).  A bad device pointer is reported as CUSTR_ERR_INVALID with the
// reference's message ("nvstrings::create_from_index bad_device_ptr"); like there, the CUDA context is unusable afterwards.
custr_column* custr_create_from_index(const void* pairs, uint32_t count, int devmem, int stype)
{
    return guarded(
        [&]() -> custr_column* {
            if (count && !pairs) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_index: null array")};
            if (count > 0x7fffffffu) throw ArgError{fail(CUSTR_ERR_ARG, "create_from_index: too many rows")};
            const int32_t n = (int32_t)count;
            BufPtr staged;
            const IndexPair* d_pairs = (const IndexPair*)pairs;
            if (!devmem && n) { staged = upload(pairs, sizeof(IndexPair) * (size_t)n); d_pairs = (const IndexPair*)staged->ptr; }
            BufPtr order;
            if (stype && n > 1) {
                order = dev_alloc(sizeof(int32_t) * (size_t)n);
                thrust::sequence(thrust::cuda::par.on(g_stream), (int32_t*)order->ptr, (int32_t*)order->ptr + n);
                thrust::sort(thrust::cuda::par.on(g_stream), (int32_t*)order->ptr, (int32_t*)order->ptr + n, IndexLess{d_pairs, stype});
                g_launches.fetch_add(2, std::memory_order_relaxed);
            }
            const int32_t* d_order = order ? (const int32_t*)order->ptr : nullptr;
            Scratch<int32_t> lens((size_t)n + 1);
            Scratch<uint8_t> ok((size_t)n + 1);
            Scratch<int> too_long(1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + n, 0, sizeof(int32_t), g_stream));
            CUSTR_CUDA(cudaMemsetAsync(too_long.get(), 0, sizeof(int), g_stream));
            if (n) LAUNCH(k_index_lengths, blocks_for(n, 256), 256, 0, d_pairs, n, d_order, lens.get(), ok.get(), too_long.get());
            int h_long = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h_long, too_long.get(), sizeof(int), cudaMemcpyDeviceToHost, g_stream));
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(n + 1));
            int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, n);  // syncs
            if (h_long) throw ArgError{fail(CUSTR_ERR_INVALID, "create_from_index: a string is longer than 2 GiB")};
            BufPtr chars = dev_alloc((size_t)total);
            if (n && total) {
                LAUNCH(k_index_copy, blocks_for(n, 256), 256, 0, d_pairs, n, d_order, (const int32_t*)off->ptr, (char*)chars->ptr);
                const cudaError_t e = cudaStreamSynchronize(g_stream);
                if (e == cudaErrorIllegalAddress) throw ArgError{fail(CUSTR_ERR_INVALID, "nvstrings::create_from_index bad_device_ptr")};
                CUSTR_CUDA(e);
            }
            BufPtr val = dev_alloc((n + 7) / 8 + 1);
            pack_bits(ok.get(), (uint8_t*)val->ptr, n);
            const int32_t nulls = count_zero_bits((const uint8_t*)val->ptr, 0, n);
            return make_column(chars, off, val, n, nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

void custr_column_free(custr_column* col) { delete col; }
uint32_t custr_size(const custr_column* col) { return col ? (uint32_t)col->n : 0; }
int64_t custr_chars_bytes(const custr_column* col) { return col ? col->nbytes : 0; }
int32_t custr_null_count(const custr_column* col) { return col ? col->nulls : 0; }
void custr_release_cached_memory(void) { custr::release_cached_memory(); }
const char* custr_chars_ptr(const custr_column* col) { return col->chars + col->first_off; }
const int32_t* custr_offsets_ptr(const custr_column* col) { return col->offsets; }
const uint8_t* custr_validity_ptr(const custr_column* col) { return col->validity; }

int custr_create_offsets(const custr_column* col, char* chars, int32_t* offsets, uint8_t* validity, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col) return fail(CUSTR_ERR_ARG, "create_offsets: null column");
            int32_t n = col->n;
            cudaMemcpyKind kind = devmem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            if (chars && col->nbytes)
                CUSTR_CUDA(cudaMemcpyAsync(chars, col->chars + col->first_off, (size_t)col->nbytes, kind, g_stream));
            if (offsets) {
                if (col->first_off == 0)
                    CUSTR_CUDA(cudaMemcpyAsync(offsets, col->offsets, sizeof(int32_t) * (size_t)(n + 1), kind, g_stream));
                else {
                    Scratch<int32_t> tmp((size_t)n + 1);
                    CUSTR_CUDA(cudaMemcpyAsync(tmp.get(), col->offsets, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyDeviceToDevice, g_stream));
                    LAUNCH(k_rebase_offsets, blocks_for(n + 1, 256), 256, 0, tmp.get(), n + 1, col->first_off);
                    CUSTR_CUDA(cudaMemcpyAsync(offsets, tmp.get(), sizeof(int32_t) * (size_t)(n + 1), kind, g_stream));
                    CUSTR_CUDA(cudaStreamSynchronize(g_stream));
                }
            }
            if (validity && n) {
                ResultBuf<uint8_t> out(validity, (n + 7) / 8, devmem);
                LAUNCH(k_export_validity, blocks_for((n + 7) / 8, 256), 256, 0, view_of(col), out.dev, 0, (unsigned long long*)nullptr);
                out.finish();
            }
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return 0;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_set_null_bitarray(const custr_column* col, uint8_t* bitarray, int empty_is_null, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !bitarray) return fail(CUSTR_ERR_ARG, "set_null_bitarray: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            ResultBuf<uint8_t> out(bitarray, (n + 7) / 8, devmem);
            Scratch<unsigned long long> cleared(1);
            CUSTR_CUDA(cudaMemsetAsync(cleared.get(), 0, 8, g_stream));
            LAUNCH(k_export_validity, blocks_for((n + 7) / 8, 256), 256, 0, view_of(col), out.dev, empty_is_null, cleared.get());
            out.finish();
            unsigned long long h = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h, cleared.get(), 8, cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return (int)h;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int64_t custr_byte_count(const custr_column* col, int32_t* lengths, int devmem)
{
    return guarded(
        [&]() -> int64_t {
            if (!col) return fail(CUSTR_ERR_ARG, "byte_count: null column");
            int32_t n = col->n;
            if (n == 0) return 0;
            if (!lengths) return col->nbytes;
            ResultBuf<int32_t> out(lengths, n, devmem);
            Scratch<unsigned long long> total(1);
            CUSTR_CUDA(cudaMemsetAsync(total.get(), 0, 8, g_stream));
            LAUNCH(k_byte_count, blocks_for(n, 256), 256, 0, view_of(col), out.dev, total.get());
            out.finish();
            return col->nbytes;
        },
        (int64_t)CUSTR_ERR_ARG, (int64_t)CUSTR_ERR_CUDA);
}

int custr_len(const custr_column* col, int32_t* lengths, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !lengths) return fail(CUSTR_ERR_ARG, "len: null argument");
            int32_t n = col->n;
            if (n == 0) return 0;
            ResultBuf<int32_t> out(lengths, n, devmem);
            LAUNCH(k_char_len, blocks_for(n, 256), 256, 0, view_of(col), out.dev);
            out.finish();
            return n - col->nulls;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

int custr_hash(const custr_column* col, uint32_t* results, int devmem)
{
    return guarded(
        [&]() -> int {
            if (!col || !results || col->n == 0) return fail(CUSTR_ERR_ARG, "hash: null argument or empty column");
            int32_t n = col->n;
            ResultBuf<uint32_t> out(results, n, devmem);
            Scratch<unsigned long long> nz(1);
            CUSTR_CUDA(cudaMemsetAsync(nz.get(), 0, 8, g_stream));
            LAUNCH(k_murmur3, blocks_for(n, 256), 256, 0, view_of(col), out.dev, nz.get());
            out.finish();
            unsigned long long h = 0;
            CUSTR_CUDA(cudaMemcpyAsync(&h, nz.get(), 8, cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            return (int)h;
        },
        (int)CUSTR_ERR_ARG, (int)CUSTR_ERR_CUDA);
}

custr_column* custr_slice_rows(const custr_column* col, int32_t first, int32_t last)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col || first < 0 || last < first || last > col->n) throw ArgError{fail(CUSTR_ERR_ARG, "slice_rows: bad range")};
            custr_column* v = new custr_column(*col);  // shares the buffers
            v->item_bounds = nullptr;                  // derived indexes belong to the parent's row range
            v->item_bounds_count = 0;
            v->n = last - first;
            v->offsets = col->offsets + first;
            v->vbit0 = col->vbit0 + first;
            int32_t ends[2] = {0, 0};
            CUSTR_CUDA(cudaMemcpyAsync(&ends[0], col->offsets + first, sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaMemcpyAsync(&ends[1], col->offsets + last, sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
            v->first_off = ends[0];
            v->nbytes = ends[1] - ends[0];
            v->nulls = col->validity ? count_zero_bits(col->validity, v->vbit0, v->n) : 0;
            if (v->nulls == 0) { v->validity = nullptr; v->vbit0 = 0; }
            return v;
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

custr_column* custr_gather(const custr_column* col, const int32_t* indices, int32_t count, int devmem)
{
    return guarded(
        [&]() -> custr_column* {
            if (!col || count < 0 || (count && !indices)) throw ArgError{fail(CUSTR_ERR_ARG, "gather: bad argument")};
            BufPtr idx;
            const int32_t* d_idx = indices;
            if (!devmem) { idx = upload(indices, sizeof(int32_t) * (size_t)count); d_idx = (const int32_t*)idx->ptr; }
            Scratch<int32_t> lens((size_t)count + 1);
            Scratch<uint8_t> ok((size_t)count + 1);
            CUSTR_CUDA(cudaMemsetAsync(lens.get() + count, 0, sizeof(int32_t), g_stream));
            if (count) LAUNCH(k_gather_lengths, blocks_for(count, 256), 256, 0, view_of(col), d_idx, count, lens.get(), ok.get());
            BufPtr off = dev_alloc(sizeof(int32_t) * (size_t)(count + 1));
            int64_t total = scan_lengths_to_offsets(lens.get(), (int32_t*)off->ptr, count);
            BufPtr chars = dev_alloc((size_t)total);
            // rows with invalid indices have zero length, so clamping them to row 0 is harmless
            Scratch<int32_t> safe((size_t)count + 1);
            if (count) {
                CUSTR_CUDA(cudaMemcpyAsync(safe.get(), d_idx, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToDevice, g_stream));
                LAUNCH(k_clamp_rows, blocks_for(count, 256), 256, 0, safe.get(), count, col->n);
                if (col->n > 0)
                    LAUNCH(k_gather_rows, blocks_for((int64_t)count * 32, 256), 256, 0, col->chars, col->offsets,
                           (const int32_t*)safe.get(), (char*)chars->ptr, (const int32_t*)off->ptr, count);
            }
            BufPtr val = dev_alloc((count + 7) / 8 + 1);
            pack_bits(ok.get(), (uint8_t*)val->ptr, count);
            int32_t nulls = count_zero_bits((const uint8_t*)val->ptr, 0, count);
            return make_column(chars, off, val, count, nulls, total);
        },
        (custr_column*)nullptr, (custr_column*)nullptr);
}

}  // extern "C"
