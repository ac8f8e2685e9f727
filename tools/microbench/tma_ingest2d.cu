// Microbenchmark (run on the B200 box): bytes per clock one SM takes in from L2 through 2-D TENSOR copies
// (cp.async.bulk.tensor.2d), as a function of the box row width (64-byte rows / SWIZZLE_64B against 128-byte rows /
// SWIZZLE_128B), the bytes in flight and the number of issuing threads.  The split GEMM kernels of gemm_tc.cu fetch
// [128 x 32] bf16 boxes (64-byte rows); this tells whether their ingest rate is a per-row or a per-byte limit.
//
//   source: bf16 [8192][1024] (2 KB row pitch, 16 MB: L2 resident), every CTA walks k over its own 128-row block
//   stage = NB boxes of [box_rows x BC columns]; ring of NS stages; lane 0 of warps 1.. issue (mode 0: the issuers split the
//   boxes of every stage; mode 1: issuer j takes stages j, j + issuers, ... whole), thread 0 recycles
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest2d tma_ingest2d.cu && ./tma_ingest2d
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

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
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int kMaxStages = 16;

// bc = box columns (bf16 elements), nb = boxes per stage, ns = stages, issuers = 1 or nb (one thread per box)
__global__ void __launch_bounds__(512, 1)
k_ingest2d(const __grid_constant__ CUtensorMap map, int bc, int nb, int ns, int issuers, int iters, int rows_total, int mode, int box_rows) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages];
    const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int box_bytes = box_rows * bc * 2, stage_bytes = nb * box_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(full0 + 8 * s, mode == 0 ? issuers : 1); mbar_init(empty0 + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int row0 = (blockIdx.x * 128) % rows_total;
    const int kcols = 1024 / bc;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0 && w >= 1 && w <= issuers) {                 // producers: warp 1 .. issuers, lane 0
        const int me = w - 1;
        for (int it = (mode == 1 ? me : 0); it < iters; it += (mode == 1 ? issuers : 1)) {
            const uint32_t s = it % ns, ph = (it / ns) & 1u;
            if (it >= ns) mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t dst = base + s * stage_bytes;
            const int per = mode == 1 ? nb : nb / issuers;
            const int b0 = mode == 1 ? 0 : me * per;
            mbar_expect_tx(full0 + 8 * s, per * box_bytes);
            for (int b = b0; b < b0 + per; ++b) {
                // box b of the stage: another row block, same k position
                const int r = (row0 + b * 2048 + (it / kcols) * 128) % rows_total;
                tma_load_2d(dst + b * box_bytes, &map, full0 + 8 * s, (it % kcols) * bc, r);
            }
        }
    } else if (threadIdx.x == 0) {                             // consumer: recycle a stage as soon as it has landed
        for (int it = 0; it < iters; ++it) {
            const uint32_t s = it % ns, ph = (it / ns) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
        }
    }
    __syncthreads();
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { printf("no encoder\n"); return 1; }
    PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
    const int rows = 8192, cols = 1024;
    void *src;
    cudaMalloc(&src, (size_t)rows * cols * 2);
    cudaMemset(src, 0, (size_t)rows * cols * 2);
    cudaFuncSetAttribute(k_ingest2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    struct Cfg { int bc, nb, ns, issuers, mode, box_rows; };
    const Cfg cfgs[] = {
        {32, 4, 5, 1, 0, 128}, {32, 4, 5, 4, 0, 128}, {32, 4, 5, 2, 1, 128}, {32, 4, 5, 5, 1, 128}, {32, 4, 6, 3, 1, 128}, {32, 4, 6, 6, 1, 128},
        {32, 4, 6, 1, 0, 112}, {32, 4, 6, 4, 0, 112}, {32, 4, 6, 6, 1, 112}, {32, 2, 12, 6, 1, 128}, {32, 2, 12, 12, 1, 128},
        {64, 2, 6, 1, 0, 128}, {64, 2, 6, 2, 0, 128}, {64, 2, 6, 3, 1, 128}, {64, 2, 6, 6, 1, 128}, {64, 4, 3, 4, 0, 128}, {64, 4, 3, 3, 1, 128},
        {64, 1, 12, 12, 1, 128}, {64, 1, 12, 6, 1, 128},
    };
    for (int ctas : {148}) {
        for (const Cfg &c : cfgs) {
            CUtensorMap m;
            const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
            const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            const cuuint32_t box[2] = {(cuuint32_t)c.bc, (cuuint32_t)c.box_rows};
            const cuuint32_t estr[2] = {1, 1};
            if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    c.bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
            const int stage_bytes = c.nb * c.box_rows * c.bc * 2;
            const size_t smem = (size_t)c.ns * stage_bytes + 1024;
            if (smem > 225 * 1024) { printf("skip (smem)\n"); continue; }
            const int iters = 3000;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            k_ingest2d<<<ctas, 512, smem>>>(m, c.bc, c.nb, c.ns, c.issuers, 200, rows, c.mode, c.box_rows);
            cudaEventRecord(a);
            k_ingest2d<<<ctas, 512, smem>>>(m, c.bc, c.nb, c.ns, c.issuers, iters, rows, c.mode, c.box_rows);
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double bytes = (double)iters * stage_bytes, clk = ms * 1e-3 * 1.965e9;
            printf("%3d CTAs  rows of %3d B  %d boxes/stage (%2d KB)  %2d stages (%3d KB in flight)  %2d issuer(s) %s, %3d-row boxes: %6.1f B/clk/SM  %5.2f TB/s\n",
                   ctas, c.bc * 2, c.nb, stage_bytes >> 10, c.ns, (c.ns * stage_bytes) >> 10, c.issuers, c.mode ? "by stage" : "by box  ", c.box_rows, bytes / clk,
                   bytes * ctas / (ms * 1e-3) / 1e12); fflush(stdout);
        }
    }
    return 0;
}
