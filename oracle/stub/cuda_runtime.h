// Host-only stand-in for <cuda_runtime.h>, used ONLY to compile the reference
// (/root/reference/cpp/src) as plain C++ with Thrust's CPP backend so that it can
// serve as the parity oracle. Test infrastructure, not product.
#pragma once
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cmath>
#include <sys/types.h>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <locale>
#define __device__
#define __host__
#define __global__
struct int2 { int x, y; };
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorIllegalAddress = 700 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t = 0) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return 0; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorName(cudaError_t) { return "cpu"; }
inline const char* cudaGetErrorString(cudaError_t) { return "cpu"; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return 1; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return 1; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return 1; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)1 << 40; return 0; }
inline unsigned int atomicAdd(unsigned int* a, unsigned int v) { unsigned int o = *a; *a += v; return o; }
