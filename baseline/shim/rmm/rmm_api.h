#pragma once
#include <cuda_runtime.h>
#ifdef __cplusplus
static inline int rmmGetInfo(size_t* f, size_t* t, cudaStream_t) { return (int)cudaMemGetInfo(f, t); }
#endif
