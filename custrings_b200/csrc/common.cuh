// Shared internals of libcustr.so: the device column object, stream/launch bookkeeping, error plumbing.
// Data layout in HBM: chars uint8[], offsets int32[n+1] (absolute into chars, offsets[0] may be
// non-zero for row-slice views), validity bits LSB-first starting at bit `vbit0` (NULL = all valid).
// Invariant kept by every constructor: null rows have zero length ("normalized" column).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <atomic>
#include <type_traits>
#include "../../include/custr.h"

namespace custr {

extern thread_local std::string g_error;
extern thread_local cudaStream_t g_stream;
extern std::atomic<long long> g_launches;

inline int fail(int code, const std::string& msg) { g_error = msg; return code; }

#define CUSTR_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            custr::g_error = std::string(#call) + ": " + cudaGetErrorName(e__) + " - " +         \
                             cudaGetErrorString(e__);                                            \
            throw custr::CudaError{e__};                                                         \
        }                                                                                        \
    } while (0)

struct CudaError { cudaError_t err; };
struct ArgError { int code; };

// every kernel launch of this library goes through LAUNCH so that custr_launch_count() is honest
#define LAUNCH(kernel, grid, block, smem, ...)                                  \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), custr::g_stream>>>(__VA_ARGS__);      \
        custr::g_launches.fetch_add(1, std::memory_order_relaxed);              \
        CUSTR_CUDA(cudaGetLastError());                                         \
    } while (0)

// Stream-ordered device buffer (cudaMallocAsync from the default pool; pool keeps freed memory cached).
struct DeviceBuf {
    void* ptr = nullptr;
    size_t bytes = 0;   // requested size
    size_t cap = 0;     // size of the underlying block (>= bytes when it came from the big-block cache)
    explicit DeviceBuf(size_t n);
    ~DeviceBuf();
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
};
using BufPtr = std::shared_ptr<DeviceBuf>;
inline BufPtr dev_alloc(size_t n) { return std::make_shared<DeviceBuf>(n ? n : 1); }

template <typename T>
struct Scratch {  // typed RAII scratch
    BufPtr buf;
    explicit Scratch(size_t count) : buf(dev_alloc(count * sizeof(T))) {}
    T* get() const { return (T*)buf->ptr; }
};

int num_sms();
void release_cached_memory();
// CUSTR_TRACE=1: synchronise the stream and print the host time since the previous trace point (development aid)
void trace_point(const char* what);

}  // namespace custr

struct custr_column {
    const char* chars = nullptr;
    const int32_t* offsets = nullptr;
    const uint8_t* validity = nullptr;  // nullptr => no nulls
    int32_t vbit0 = 0;                  // bit offset of row 0 inside validity
    int32_t n = 0;
    int32_t nulls = 0;
    int32_t first_off = 0;              // offsets[0]
    int64_t nbytes = 0;                 // offsets[n] - offsets[0]
    custr::BufPtr chars_buf, offsets_buf, validity_buf;  // owners (shared between views); empty when adopted
    // derived, pattern-independent index built lazily by the bitstream regex tier and kept with the (immutable) column:
    // first row of every 32 KiB work item (4 bytes per 32 KiB of chars)
    mutable custr::BufPtr item_bounds;
    mutable int32_t item_bounds_count = 0;
    mutable int32_t item_bounds_bytes = 0;   // item size the index was built for
    // CUDA IPC (custr_ipc_export / custr_ipc_import): the exported copy lives as long as the column it was made from; an
    // imported column keeps the opened mapping until it is freed
    mutable std::shared_ptr<void> ipc_owner;
};

struct custr_category {
    custr_column* keys = nullptr;  // sorted distinct keys (null key first if any null rows)
    custr::BufPtr values_buf;      // int32[n]
    int32_t n = 0;
    bool has_null_key = false;
};

namespace custr {

// device-side view passed by value to kernels
struct ColView {
    const char* __restrict__ chars;
    const int32_t* __restrict__ offsets;
    const uint8_t* __restrict__ validity;
    int32_t vbit0;
    int32_t n;
    __host__ __device__ __forceinline__ bool valid(int i) const
    {
        if (!validity) return true;
        int b = vbit0 + i;
        return (validity[b >> 3] >> (b & 7)) & 1;
    }
};
inline ColView view_of(const custr_column* c) { return ColView{c->chars, c->offsets, c->validity, c->vbit0, c->n}; }

// column.cu helpers used by the other translation units
custr_column* make_column(BufPtr chars, BufPtr offsets, BufPtr validity, int32_t n, int32_t nulls, int64_t nbytes);
// Build a column from per-row (valid, length) already scanned into offsets; chars buffer filled by caller.
custr_column* all_null_column(int32_t n);
// exclusive scan of int32 lengths[n] into offsets[n+1] (offsets[n] = total); returns total (sync D2H)
int64_t scan_lengths_to_offsets(const int32_t* lengths, int32_t* offsets, int32_t n);
// pack a bool/uint8 array [n] into LSB-first bits
void pack_bits(const uint8_t* flags, uint8_t* bits, int32_t n);
int32_t count_zero_bits(const uint8_t* bits, int32_t vbit0, int32_t n);
// copy results to host or leave on device according to devmem
template <typename T>
struct ResultBuf {  // device staging for a caller result array that may live on the host
    T* dev;
    T* user;
    bool devmem;
    size_t count;
    BufPtr tmp;
    ResultBuf(T* user_, size_t count_, int devmem_) : user(user_), devmem(devmem_ != 0), count(count_)
    {
        if (devmem) dev = user;
        else { tmp = dev_alloc(count * sizeof(T)); dev = (T*)tmp->ptr; }
    }
    void finish()
    {
        if (!devmem && count) {
            CUSTR_CUDA(cudaMemcpyAsync(user, dev, count * sizeof(T), cudaMemcpyDeviceToHost, g_stream));
            CUSTR_CUDA(cudaStreamSynchronize(g_stream));
        }
    }
};

// upload a small host blob
BufPtr upload(const void* host, size_t bytes);

// Column-major result assembly shared by the "one output column per token / match / group" operations:
// lens[c*(n+1)+row] and valid[c*n+row] are filled by the caller's length pass; `copy` must launch the kernel that writes
// the bytes of (row, c) at outs[c].chars + outs[c].offsets[row].  Returns the ncols new columns.
struct ColumnOut { char* chars; const int32_t* offsets; };
std::vector<custr_column*> assemble_columns(int32_t n, int ncols, int32_t* lens, const uint8_t* valid,
                                            void (*copy)(const ColumnOut* d_outs, void* ctx), void* ctx);

// generic guard for the extern "C" layer
template <typename F, typename R>
R guarded(F&& f, R on_arg, R on_cuda)
{
    try { return f(); }
    catch (const CudaError&) { return on_cuda; }
    catch (const ArgError& e) {  // integer entry points report the code that was raised (CUSTR_ERR_INVALID stays distinguishable from a null argument)
        if constexpr (std::is_integral<R>::value) return e.code ? (R)e.code : on_arg;
        else return on_arg;
    }
    catch (const std::bad_alloc&) { g_error = "host allocation failed"; return on_cuda; }
}

}  // namespace custr
