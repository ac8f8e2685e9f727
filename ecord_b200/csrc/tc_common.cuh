// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (gemm_tc.cu, tail_tc.cu, qtc.cu): mbarrier, TMA,
// UMMA descriptors, tcgen05.mma / commit / ld, and the host-side tensor-map encoder.
#pragma once
#include "common.cuh"
#include <math.h>
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched at run time (no link-time libcuda dependency)

namespace {

constexpr int H = ECORD_DIM_H;
constexpr int BM = 128;          // rows per tile (UMMA_M)
constexpr int BK = 64;           // bf16 elements per k-chunk = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 576;    // 18 warps: TMA, MMA, 16 epilogue (four per TMEM lane quarter)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// round-to-nearest onto the tf32 grid (10 explicit mantissa bits); the tensor core ignores the 13 low bits
__host__ __device__ __forceinline__ float tf32_rn(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
    union { float f; uint32_t u; } v; v.f = x; v.u = (v.u + 0x1000u) & 0xffffe000u; return v.f;
#endif
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float *v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// MUFU.TANH (rel. error ~2^-11): the bf16 path rounds its outputs to 8 mantissa bits anyway
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Gate functions of the split (bf16x3) kernels' epilogues (GRU cell, tanh(h') for the projection, tanh(P) in the tail): exp through MUFU.EX2 (2 ulp) and an approximate reciprocal -- fp32-grade (~3e-7
// absolute on sigmoid / tanh) at a third of the instructions of expf / tanhf.  A/B against the library functions
// (-DECORD_GATES_LIBM) on the reference fixtures of tests/test_bestcut_gpu.py: best cut and per-trajectory best unchanged, whole
// trajectories identical 0.975 / 0.950 / 0.900 / 0.31 / 0.63 against 0.975 / 0.950 / 0.912 / 0.31 / 0.50 (ER200 / ER500 / BA500 /
// ER2000 / ER10000): differences of single trajectories either way, the size of fp32 rounding; GRU kernel 79.6 -> 75.2 us at ER500.
// (MUFU.TANH, 2^-11, is NOT fp32-grade: it costs a fifth of the trajectories, oracle/ablate_precision.py.)
#ifndef ECORD_GATES_LIBM
__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) {
    const float e = __expf(-2.0f * fabsf(x));                   // in (0, 1]
    return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
#else
__device__ __forceinline__ float gate_sigmoid(float x) { return sigmoidf_acc(x); }
__device__ __forceinline__ float gate_tanh(float x) { return tanhf(x); }
#endif

struct TcParams {
    // GRU
    const float *hidden;          // [2][Mp][1024] fp32
    __nv_bfloat16 *hidden_bf16;   // [2][Mp][1024]
    __nv_bfloat16 *th_bf16;       // [Mp][1024]
    const float *bih, *bhh;       // [3072]
    // PROJ
    const float *p1_b;            // [512]
    float *y1;                    // [Mp][512]
    int rows, rows_pad;
    const int32_t *step_base;     // device step counter (or NULL)
    int step_off;
    int debug;                    // diagnostics (ECORD_TC_DEBUG, bits): 1 no MMA issue, 2 no epilogue math, 4 no TMA (rejected), 8 prologue +
                                  // teardown only, 32 hi planes only, 64 every cluster fetches tile (0, 0), 128 no epilogue global traffic
    int sched_full, sched_kn, sched_narrow;      // GRU tile schedule (gemm_tc.cu: tc2_tile)
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encoder(PFN_encodeTiled *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return ECORD_E_DEVICE;
    *out = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

// 2-D bf16 row-major [rows][cols], box = [box_rows][box_cols]: 64 columns with the 128-byte swizzle, 32 with the 64-byte one
int make_map(PFN_encodeTiled enc, CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
             uint32_t box_cols = BK) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ECORD_E_DEVICE;
}

}  // namespace
