// common.cuh -- shared helpers for the ecord_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/ecord_b200.h"

#define ECORD_RET_IF(cond, code) do { if (cond) return (code); } while (0)
#define ECORD_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (int)_e; } while (0)
// after a kernel launch: report launch-configuration errors without synchronising
#define ECORD_LAUNCH_CHECK() do { cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) return (int)_e; } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): the kernels of one rollout step form a strictly serial chain, and each
// has a few microseconds of work that does not depend on its predecessor (barrier / TMEM set-up, staging constant
// weights into shared memory).  Launched with the attribute below, a kernel may become resident while its predecessor
// drains; pdl_wait() blocks until the predecessor has completed and its memory is visible, so every read or write of
// step data must come after it.  Outside a PDL launch both device calls are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const int enabled = [] { const char *e = getenv("ECORD_PDL"); return e ? atoi(e) : 1; }();   // ECORD_PDL=0: plain launches (A/B)
    cfg.attrs = attr; cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define ECORD_PDL_LAUNCH(...) do { cudaError_t _e = launch_pdl(__VA_ARGS__); if (_e != cudaSuccess) return (int)_e; } while (0)

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr float kLeaky = 0.01f;       // nn.LeakyReLU default slope
constexpr float kLnEps = 1e-5f;       // nn.LayerNorm default eps

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : kLeaky * x; }

// ---- Philox4x32-10 (counter-based; stateless) for tau > 0 sampling ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float uniform01_open(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }


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

// packed fp32 pairs (sm_100 FFMA2)
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long abs2(unsigned long long v) {
    float a, b;
    unpack2(v, a, b);
    return pack2(fabsf(a), fabsf(b));
}

// plane offset of unit (g,t): graph_ptr[g]*T + t*n_g   (see include/ecord_b200.h)
__device__ __forceinline__ long long unit_base(const int32_t *graph_ptr, int g, int t, int T, int &n) {
    const int lo = graph_ptr[g];
    n = graph_ptr[g + 1] - lo;
    return (long long)lo * T + (long long)t * n;
}

// accurate, deterministic activations (same formulas on every path so modes are comparable)
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// u = LReLU(msg_in([last_sel (32); glob0; 0])) for one row, computed by one warp (lane holds last_sel[lane]).
// wt = transposed msg_in weight in shared memory [34][64], wb = bias [64].  Outputs u[lane], u[lane+32].
// (_ActorBase._update_internal_state, actors/_base.py:247-269; rnn_decoder.py:121-124)
constexpr int kMsgPitch = 65;         // row pitch of the transposed msg_in weight in shared memory: 64 outputs + 1 (see msg_in_load)
constexpr int kMsgWtFloats = 34 * kMsgPitch;
__device__ __forceinline__ void msg_in_row(const float *wt, const float *wb, float x, float glob0, int lane, float &u0,
                                           float &u1) {
    float a0 = wb[lane], a1 = wb[lane + 32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float xi = __shfl_sync(0xffffffffu, x, i);
        a0 = fmaf(wt[i * kMsgPitch + lane], xi, a0);
        a1 = fmaf(wt[i * kMsgPitch + lane + 32], xi, a1);
    }
    a0 = fmaf(wt[32 * kMsgPitch + lane], glob0, a0);     // + wt[33][.] * 0: the 2nd global feature is identically 0 (quirk C)
    a1 = fmaf(wt[32 * kMsgPitch + lane + 32], glob0, a1);
    u0 = lrelu(a0);
    u1 = lrelu(a1);
}
// msg_w [64][34] is read in memory order (whole lines per warp) and transposed on the way into shared memory; the pitch of
// 65 makes both the transposing stores (lanes <-> consecutive inputs of one output row) and the row reads conflict-free.
// (A gather in output order would touch 32 different lines per load instruction in every CTA's set-up.)
__device__ __forceinline__ void msg_in_load(const float *__restrict__ msg_w, const float *__restrict__ msg_b, float *wt,
                                            float *wb) {
    for (int i = threadIdx.x; i < 34 * 64; i += blockDim.x) wt[(i % 34) * kMsgPitch + (i / 34)] = msg_w[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) wb[i] = msg_b[i];
}
__device__ __forceinline__ void msg_in_store(const ecord_policy &p, int m, int lane, float u0, float u1) {
    p.u[(long long)m * 64 + lane] = u0;
    p.u[(long long)m * 64 + lane + 32] = u1;
    if (p.u_bf16) {
        __nv_bfloat16 *ub = reinterpret_cast<__nv_bfloat16 *>(p.u_bf16);
        const __nv_bfloat16 h0 = __float2bfloat16(u0), h1 = __float2bfloat16(u1);
        ub[(long long)m * 64 + lane] = h0;
        ub[(long long)m * 64 + lane + 32] = h1;
        if (p.split) {                  // ECORD_GEMM_BF16X3: the bf16 remainders, one plane ([rows_pad][64]) further
            ub += (long long)p.rows_pad * 64;
            ub[(long long)m * 64 + lane] = __float2bfloat16(u0 - __bfloat162float(h0));
            ub[(long long)m * 64 + lane + 32] = __float2bfloat16(u1 - __bfloat162float(h1));
        }
    }
}
