// Microbenchmark behind DESIGN.md's analysis of the split GEMM kernels (run on the B200 box): how many bytes per clock
// can one SM take in from L2 through the bulk-copy (TMA) engine, and does a multicast copy -- one L2 read delivered to
// the same shared-memory offset of several CTAs of a cluster -- count against the receiving SM's limit?
//
//   mode 0  unicast : every CTA of a 4-CTA cluster fetches its own 16 KB per stage (distinct addresses)
//   mode 1  same    : every CTA fetches the SAME 16 KB as its cluster peers (unicast; does L2 / the crossbar merge them?)
//   mode 2  mcast   : every CTA fetches one 4 KB quarter and multicasts it to all four CTAs (each still receives 16 KB/stage)
//
// Each CTA runs a ring of kStages x 16 KB; one thread issues the copies and recycles a stage as soon as it has landed
// (nothing reads the data).  For modes 0/1 the ring is CTA-local; for mode 2 a stage may be refilled only when all four
// CTAs have seen it land, hence the remote arrivals on every peer's `empty` barrier.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest tma_ingest.cu && ./tma_ingest
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kStages = 6;
constexpr int kStageBytes = 16384;
constexpr int kCluster = 4;

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
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(64, 1)
k_ingest(const uint8_t *__restrict__ src, size_t src_bytes, int iters, int mode, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, mode == 2 ? kCluster : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int cluster_id = blockIdx.x / kCluster;
    const size_t stride = (size_t)kStageBytes * 37;            // successive stages of a CTA come from different lines / slices
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {                                    // producer
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            if (it >= kStages) mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t dst = smem_u32(smem) + s * kStageBytes;
            size_t off = ((size_t)(mode == 0 ? blockIdx.x : cluster_id) * 2654435761ull + (size_t)it * stride) % (src_bytes - kStageBytes);
            off &= ~(size_t)1023;
            mbar_expect_tx(full0 + 8 * s, kStageBytes);
            if (mode == 2) {
                const uint32_t q = kStageBytes / kCluster;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"(dst + rank * q), "l"(src + off + rank * q), "r"(q), "r"(full0 + 8 * s), "h"((uint16_t)0xF) : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src + off), "r"((uint32_t)kStageBytes), "r"(full0 + 8 * s) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {                            // consumer: recycle a stage as soon as it has landed
        for (int it = 0; it < iters; ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            if (mode == 2) {
                for (uint32_t r = 0; r < kCluster; ++r) mbar_arrive_remote(mapa_u32(empty0 + 8 * s, r));
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 32) cycles[blockIdx.x] = t1;
    if (threadIdx.x == 0) cycles[gridDim.x + blockIdx.x] = t0;
}

int main() {
    const size_t src_bytes = 96ull << 20;                      // L2-resident after the first pass
    uint8_t *src;
    long long *cyc;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    const int max_ctas = 148;
    cudaMalloc(&cyc, 2 * max_ctas * sizeof(long long));
    cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes);
    const int iters = 4000;
    const char *names[3] = {"unicast, distinct addresses", "unicast, same addresses in a cluster", "multicast (each CTA fetches a quarter)"};
    for (int ctas : {148, 72, 36}) {
        for (int mode = 0; mode < 3; ++mode) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            k_ingest<<<ctas, 64, kStages * kStageBytes>>>(src, src_bytes, 200, mode, cyc);        // warm-up (L2)
            cudaEventRecord(a);
            k_ingest<<<ctas, 64, kStages * kStageBytes>>>(src, src_bytes, iters, mode, cyc);
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double bytes_per_cta = (double)iters * kStageBytes;
            const double clk = ms * 1e-3 * 1.965e9;
            printf("%3d CTAs  %-44s %8.1f us  landed %6.1f B/clk/SM  (%5.2f TB/s into shared memory, %5.2f TB/s read from L2)\n", ctas,
                   names[mode], ms * 1e3, bytes_per_cta / clk, bytes_per_cta * ctas / (ms * 1e-3) / 1e12,
                   bytes_per_cta * ctas / (mode == 2 ? kCluster : 1) / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
