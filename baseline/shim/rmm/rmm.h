// Minimal RMM shim for building the reference custrings CUDA sources without librmm (absent from this image):
// allocation through the stream-ordered CUDA pool, which behaves like RMM's pool mode.  Baseline infrastructure only.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <sstream>
#include <iostream>
#include <stdexcept>
#include <locale>
typedef enum { RMM_SUCCESS = 0, RMM_ERROR_CUDA_ERROR, RMM_ERROR_INVALID_ARGUMENT, RMM_ERROR_NOT_INITIALIZED, RMM_ERROR_OUT_OF_MEMORY, RMM_ERROR_UNKNOWN, RMM_ERROR_IO } rmmError_t;
template <typename T>
inline rmmError_t RMM_ALLOC(T** p, size_t sz, cudaStream_t s)
{
    cudaError_t e = cudaMallocAsync((void**)p, sz ? sz : 1, s);
    return e == cudaSuccess ? RMM_SUCCESS : (e == cudaErrorMemoryAllocation ? RMM_ERROR_OUT_OF_MEMORY : RMM_ERROR_CUDA_ERROR);
}
inline rmmError_t RMM_FREE(void* p, cudaStream_t s) { return cudaFreeAsync(p, s) == cudaSuccess ? RMM_SUCCESS : RMM_ERROR_CUDA_ERROR; }
