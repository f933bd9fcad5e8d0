// malloc-backed stand-in for RMM (un-vendored dependency of the reference; allocation only).
#pragma once
#include <cuda_runtime.h>
typedef enum { RMM_SUCCESS = 0, RMM_ERROR_CUDA_ERROR, RMM_ERROR_INVALID_ARGUMENT, RMM_ERROR_NOT_INITIALIZED, RMM_ERROR_OUT_OF_MEMORY } rmmError_t;
template <typename T> inline rmmError_t RMM_ALLOC(T** p, size_t sz, cudaStream_t) { *p = (T*)malloc(sz ? sz : 1); return RMM_SUCCESS; }
inline rmmError_t RMM_FREE(void* p, cudaStream_t) { free(p); return RMM_SUCCESS; }
