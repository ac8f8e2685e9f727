// common.cuh -- shared helpers for the ecord_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/ecord_b200.h"

#define ECORD_RET_IF(cond, code) do { if (cond) return (code); } while (0)
#define ECORD_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (int)_e; } while (0)
// after a kernel launch: report launch-configuration errors without synchronising
#define ECORD_LAUNCH_CHECK() do { cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) return (int)_e; } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr float kLeaky = 0.01f;       // nn.LeakyReLU default slope
constexpr float kLnEps = 1e-5f;       // nn.LayerNorm default eps

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : kLeaky * x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// plane offset of unit (g,t): graph_ptr[g]*T + t*n_g   (see include/ecord_b200.h)
__device__ __forceinline__ long long unit_base(const int32_t *graph_ptr, int g, int t, int T, int &n) {
    const int lo = graph_ptr[g];
    n = graph_ptr[g + 1] - lo;
    return (long long)lo * T + (long long)t * n;
}

// accurate, deterministic activations (same formulas on every path so modes are comparable)
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
