// Microbenchmarks behind DESIGN.md's analysis of the Q kernel (run on the B200 box):
//   (1) tcgen05.ld throughput: bytes per clock per SM moved from tensor memory to registers (the Q kernel reads its
//       128 x 65 fp32 accumulator once per (vertex tile, trajectory));
//   (2) legacy mma.sync.m16n8k16 (f16 x f16 -> f32) throughput per SM, the alternative that keeps accumulators in registers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_rates tc_rates.cu && ./tc_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- (1) tcgen05.ld ----
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_ldtm(int iters, int cols, float *out, long long *cycles) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < cols; c += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(trow + (uint32_t)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += __uint_as_float(r[i] & 0x3f800000u);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(128u) : "memory");
}

// ---- (2) mma.sync m16n8k16 f16 -> f32 ----
__global__ void __launch_bounds__(512) k_hmma(int iters, float *out, long long *cycles) {
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u};
    float c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)       // 8 independent accumulators (one 16 x 64 block of the Q head)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    // tcgen05.ld: 4 CTAs per SM x 4 warps (the Q kernel's shape) and 1 CTA x 4 / 8 / 16 warps
    for (int ctas_per_sm : {1, 2, 4}) {
        k_ldtm<4><<<148 * ctas_per_sm, 128>>>(iters, 64, out, cyc);
        cudaDeviceSynchronize();
        k_ldtm<4><<<148 * ctas_per_sm, 128>>>(iters, 64, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double bytes_per_sm = (double)ctas_per_sm * 128 * 64 * 4 * iters;
        printf("tcgen05.ld x16, %d CTAs/SM x 4 warps: %.1f B/clk/SM (%s)\n", ctas_per_sm, bytes_per_sm / (double)h, cudaGetErrorString(e));
    }
    for (int warps : {8, 16}) {
        if (warps == 8) k_ldtm<8><<<148, 256>>>(iters, 64, out, cyc); else k_ldtm<16><<<148, 512>>>(iters, 64, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("tcgen05.ld x16, 1 CTA/SM x %d warps: %.1f B/clk/SM (%s)\n", warps, (double)warps * 32 * 64 * 4 * iters / (double)h, cudaGetErrorString(e));
    }
    for (int ctas_per_sm : {1, 2, 4}) {
        k_hmma<<<148 * ctas_per_sm, 512>>>(iters, out, cyc);
        cudaDeviceSynchronize();
        k_hmma<<<148 * ctas_per_sm, 512>>>(iters, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double macs_per_sm = (double)ctas_per_sm * 16 * 8.0 * iters * 16 * 8 * 16;
        printf("mma.sync m16n8k16 f16->f32, %d x 16 warps/SM: %.0f MAC/clk/SM (%s)\n", ctas_per_sm, macs_per_sm / (double)h, cudaGetErrorString(e));
    }
    return 0;
}
