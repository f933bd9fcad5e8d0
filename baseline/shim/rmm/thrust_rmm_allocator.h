#pragma once
#include <memory>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/system/cuda/execution_policy.h>
#include <thrust/system_error.h>
#include "rmm.h"
namespace rmm {
template <typename T> using device_vector = thrust::device_vector<T>;
struct exec_policy_t {
    auto on(cudaStream_t s) const { return thrust::cuda::par.on(s); }
};
inline std::unique_ptr<exec_policy_t> exec_policy(cudaStream_t = 0) { return std::unique_ptr<exec_policy_t>(new exec_policy_t); }
}  // namespace rmm
