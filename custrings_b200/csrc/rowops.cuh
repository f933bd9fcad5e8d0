// Per-row literal operations (find / replace / split / tokenize) as __host__ __device__ functions.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"
