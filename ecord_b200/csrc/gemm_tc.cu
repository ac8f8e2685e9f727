// gemm_tc.cu -- the two dense contractions of the policy step on 5th-gen tensor cores (sm_100a):
//
//   GRU  :  [M, 64+1024] x [1088, 3072]  fused with the GRU cell   (models/rnn_decoder.py:369-381 -> nn.GRUCell)
//   PROJ :  [M, 1024]    x [1024, 512]   + bias                    (models/rnn_decoder.py:84-86, proj_rnn_to_gnn.1)
//
// Hand-written tcgen05 pipeline on CTA pairs (cta_group::2, k_gemm_tc2 below), persistent:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d -> shared-memory ring, mbarrier complete_tx; programmatic dependent
//               launch: set-up overlaps the predecessor
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.cta_group::2.kind::f16
//               (bf16 x bf16 -> fp32 accumulate in TMEM), tcgen05.commit releases stages / signals the epilogue
//   warps 2..17 epilogue: tcgen05.ld (32 lanes x 16 columns at a time) -> registers -> fused math -> global;
//               warp w reads TMEM lanes 32*(w%4).., the four warps of a lane quarter split the tile's columns
//
// GRU tile = 256 rows (128 per CTA) x W hidden units, W = 64, 32 or 16.  The weights are packed (ecord_pack_weights) in
// 16-unit sub-blocks so that a tile of k = W/16 sub-blocks is a CONTIGUOUS row range of each packed matrix:
//   Wh      [64 sub-blocks][r16 | z16 | hn16][1024]      48 rows per sub-block
//   Wu_main [64 sub-blocks][r16 | z16 | 0  ][64]         48 rows per sub-block (the zero rows initialise the hn columns)
//   Wu_in   [1024][64]                                    W_in: 16 rows per sub-block
// and ONE accumulator [128 x 64k] holds all four gate pre-activations of the tile:
//   columns [48 j, 48 j + 48): r | z | hn of sub-block j;   columns [48 k + 16 j, +16): in of sub-block j
//   u-chunks : A = u,          B = Wu_main rows (N = 48k, initialises) and Wu_in rows (N = 16k, a second MMA)
//   h-chunks : A = h chunk,    B = Wh rows (N = 48k), accumulates
// so gi_n and gh_n stay separate (n = tanh(gi_n + r * gh_n)) at no extra MMA work, and the epilogue applies the
// cell per element:  h' = (h - n) * z + n.  h is double-buffered: buffer (s & 1) holds the state after s steps.
// Why three widths: a launch is 16 x Mp/256 tiles of 64 units on 74 CTA pairs, and the last wave of such tiles is mostly
// empty (ER500: 320 tiles = 4.32 waves; ER10000: 80 tiles = 1.08 waves; 16 trajectories per GPU: 16 tiles on 74 pairs).
// The leftover tiles are therefore cut into 32- or 16-unit ones (launch_gru_tc_ex picks the width that minimises the
// makespan) -- the results do not depend on the width: every output element sees the same k-chunks in the same order.

#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {
// MODE 0 = GRU : see above; accumulator = 256 TMEM columns (W = 64), two accumulators = all 512 columns
// MODE 1 = PROJ: tile = 256 rows x 176 outputs; 16 chunks (N=176) from (th, W1); accumulator = 176 of 256 columns
//
// Persistent: grid = min(#tiles, #SMs); the accumulator is double-buffered in TMEM (tmem_full / tmem_empty barriers), so the
// epilogue of tile i (TMEM -> registers -> GRU cell -> global) runs while TMA + MMA already work on tile i+1.
template <int MODE> struct TcCfg;
template <> struct TcCfg<0> { static constexpr int N_MAIN = 192, B_ROWS = 256, ACC_COLS = 256, TILE_N = 64; };
// PROJ: 512 outputs as three column blocks of 176 (the last one runs 16 columns past the matrix: TMA zero-fills them, the
// epilogue drops them): 3 x 20 pair tiles fit one wave of 74 clusters, where 4 x 20 tiles of 128 needed two
template <> struct TcCfg<1> { static constexpr int N_MAIN = 176, B_ROWS = 176, ACC_COLS = 256, TILE_N = 176; };
// MODE 2 = PROJ when every CTA pair has at most ONE tile (the usual case: 3 x Mp/256 <= 74): the kernel is then a single
// latency-bound pass over 16 k-chunks, so it runs seven stages deep and its epilogue staging tile reuses the stage memory
// (all loads have landed and all MMAs have retired when the accumulator is complete).
template <> struct TcCfg<2> { static constexpr int N_MAIN = 176, B_ROWS = 176, ACC_COLS = 256, TILE_N = 176; };

__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// tensor maps of a launch, passed as one __grid_constant__ block.  GRU: u, a0 / a1 (h of either parity), and per tile width
// (index 0: 64 units, 1: the narrow width of this launch) b = Wh, wum = Wu_main, wui = Wu_in.  PROJ: a0 = tanh(h), b[0] = W1.
struct TcMaps { CUtensorMap u, a0, a1, b[2], wum[2], wui[2]; };

// =================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster (= the two SMs of a TPC) work on a 256-row tile.
// Each CTA owns 128 rows (its own A tile, accumulator and epilogue -- exactly as above) but holds only HALF of the
// weight tile: the leader's single tcgen05.mma.cta_group::2 (M = 256) reads A from both CTAs and the two B halves
// from the two shared memories.  Per k-chunk an SM therefore ingests 16 KB + 12 KB instead of 16 KB + 24 KB: the
// single-CTA kernel is bound by the L2 -> SM path (435 MB per GRU launch), not by the tensor pipe.
//   - TMA: every CTA loads its own A rows and its half of B; all four loads complete_tx on the LEADER's full barrier
//   - MMA: issued by the leader only; tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers
//   - tmem_empty lives in the leader and collects the 8 epilogue warps of both CTAs (remote mbarrier arrive)
// =================================================================================================
constexpr int kStages2 = 4;      // the pair kernel's stages are 28-32 KB (half of B per CTA): six fit in 192 KB
// kProducers TMA warps (warp 0 and, for kProducers > 1, the warps behind the epilogue; producer j takes the ring iterations
// j, j + kProducers, ...).  A lone thread issuing every stage of a ring of tensor copies sustains 24-26 B/clk/SM with 64-byte
// box rows and 31-34 B/clk/SM with 128-byte rows however deep the ring is, where two threads taking whole stages in turn reach
// 52-55 and 72 B/clk/SM (tools/microbench/tma_ingest2d.cu, measured on the B200).  In THESE kernels a second producer changes
// nothing (split GRU 97.2 vs 99 us): with the MMAs switched off the ring already moves 609 MB per launch in 50 us = 12.2 TB/s,
// the L2 -> SM delivery limit of the chip, whatever the ring depth (3, 4, 5 stages: 49.7 / 50.2 / 50.0 us) -- so one it is.
#ifndef ECORD_TC_PRODUCERS
#define ECORD_TC_PRODUCERS 1
#endif
constexpr int kProducers = ECORD_TC_PRODUCERS;
constexpr int kEpiThreads = kThreads - 64;                    // 16 epilogue warps
constexpr int kThreads2 = kThreads + 32 * (kProducers - 1);   // + the extra producer warps
constexpr int kHsPitch = 68;     // floats per row of the fp32 epilogue staging tile (64 + 4: 16-byte aligned rows, 4-way STS.128)
constexpr int kBsPitch = 72;     // bf16 per row of the two bf16 staging tiles (64 + 8)
constexpr int kStageTileBytes = BM * kHsPitch * 4 + 2 * BM * kBsPitch * 2;      // 34816 + 36864 bytes
// x ~ hi + lo, both bf16 (|x - hi - lo| <= 2^-17 |x|); four values -> packed hi words and lo words
__device__ __forceinline__ void split_bf16x4(const float4 v, uint2 &hi, uint2 &lo) {
    const __nv_bfloat16 h0 = __float2bfloat16(v.x), h1 = __float2bfloat16(v.y), h2 = __float2bfloat16(v.z), h3 = __float2bfloat16(v.w);
    const __nv_bfloat16 l0 = __float2bfloat16(v.x - __bfloat162float(h0)), l1 = __float2bfloat16(v.y - __bfloat162float(h1));
    const __nv_bfloat16 l2 = __float2bfloat16(v.z - __bfloat162float(h2)), l3 = __float2bfloat16(v.w - __bfloat162float(h3));
    hi.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    hi.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    lo.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    lo.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
}
__device__ __forceinline__ bool elect_one() {          // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// (collector hints -- .collector::a::fill / ::lastuse on the two MMAs that share A hi -- changed neither the kernel's time nor its
// shared-memory operand wavefronts, 7.37 M per launch with and without: measured and dropped)
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// NS = 1: bf16 operands (ECORD_GEMM_BF16).  NS = 3: every operand is a bf16 hi part plus a bf16 lo remainder
// (x ~ hi + lo to 2^-17) and the three products hi.hi + lo.hi + hi.lo are accumulated (ECORD_GEMM_BF16X3).  A stage of
// the NS = 3 pipeline holds all four operand tiles of a 32-element k-chunk -- A hi, A lo, B hi, B lo (64-byte rows,
// SWIZZLE_64B) -- so each byte is fetched once from L2 for its two (hi parts) or one (lo parts) products: the kernel is
// bound by the L2 -> SM path (three separate (A, B) fetches per k-chunk moved 914 MB per GRU launch at 9.2 TB/s, 75 % of
// what the L2 slices deliver; this layout moves two thirds of that).  The lo parts live one "plane" behind the hi parts
// in the same tensor maps (A: rows_pad rows further, B: the weight matrix's row count further).
// The epilogue of NS = 3 uses the accurate sigmoid / tanh and writes h' and tanh(h') as hi / lo pairs: with it the
// recurrent state tracks the reference's fp32 arithmetic to ~1e-6 instead of ~1e-3 (oracle/ablate_precision.py).
template <int NS> __host__ __device__ constexpr int tc2_bk() { return NS == 3 ? 32 : 64; }      // bf16 elements per k-chunk
template <int MODE, int NS> __host__ __device__ constexpr int tc2_stage_bytes() {
    return (NS == 3 ? 2 : 1) * (BM + TcCfg<MODE>::B_ROWS / 2) * tc2_bk<NS>() * 2;              // per CTA: own A rows, half of B
}
// pipeline depth: the GRU split kernel needs its bytes in flight (TMA latency x 45 B/clk of ingest per SM): three stages cost
// 10 us per launch against four; its epilogue stages one fp32 tile only (tanh is taken in the store pass), which pays for a fifth
#ifndef ECORD_GRU_STAGES
#define ECORD_GRU_STAGES 5
#endif
template <int MODE, int NS> __host__ __device__ constexpr int stages2() { return MODE == 2 ? 7 : (MODE == 0 && NS == 3) ? ECORD_GRU_STAGES : 4; }
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {      // K-major SWIZZLE_64B: 8-row groups 512 B apart
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// tile i of cluster `cluster_id`.  PROJ: pair tiles cluster_id + i * num_clusters of 256 rows x TILE_N columns.  GRU: first
// prm.sched_full rounds of 64-unit tiles, then the leftover 64-unit tiles cut into prm.sched_kn-sub-block ones
// (prm.sched_narrow of them), both dealt round-robin; nb = first 16-unit sub-block, k = sub-blocks (4, 2 or 1).
// GRU tile i of cluster `cluster_id`: first `full` rounds of 64-unit tiles, then the leftover 64-unit tiles cut into kn-sub-block
// ones (`narrow` of them), both dealt round-robin; nb = first 16-unit sub-block, k = sub-blocks (4, 2 or 1).  Host + device: the
// same code enumerates the tiles for ecord_gru_tile_schedule(), which tests/ check for exact coverage at any batch size.
__host__ __device__ __forceinline__ bool gru_tile(int i, int cluster_id, int num_clusters, int full, int kn, int narrow,
                                                  int &m0_pair, int &nb, int &k) {
    int id;
    k = 4;
    if (i < full) {
        id = cluster_id + i * num_clusters;
        nb = (id & 15) * 4;
    } else {
        const int j = cluster_id + (i - full) * num_clusters;
        if (j >= narrow) return false;
        k = kn;
        const int per = 4 / k;
        id = full * num_clusters + j / per;
        nb = (id & 15) * 4 + (j % per) * k;
    }
    m0_pair = (id >> 4) * (2 * BM);
    return true;
}
// tile i of cluster `cluster_id`.  PROJ: pair tiles cluster_id + i * num_clusters of 256 rows x TILE_N columns; GRU: gru_tile
template <int MODE>
__device__ __forceinline__ bool tc2_tile(int i, int cluster_id, int num_clusters, int num_tiles, int nblk, const TcParams &prm,
                                         int &m0_pair, int &nb, int &k) {
    if (MODE == 0) return gru_tile(i, cluster_id, num_clusters, prm.sched_full, prm.sched_kn, prm.sched_narrow, m0_pair, nb, k);
    const int tile = cluster_id + i * num_clusters;
    if (tile >= num_tiles) return false;
    m0_pair = (tile / nblk) * (2 * BM);
    nb = tile % nblk;
    k = 0;
    return true;
}

template <int MODE, int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
k_gemm_tc2(const __grid_constant__ TcMaps maps, TcParams prm, int num_tiles, int nblk) {
    // num_tiles = number of 256-row x TILE_N pair tiles (PROJ); the GRU's tiles come from prm.sched_* (tc2_tile)
    using C = TcCfg<MODE>;
    constexpr int BKK = tc2_bk<NS>(), ROWB = BKK * 2;                           // k-chunk width: elements, bytes per tile row
    constexpr int PARTS = NS == 3 ? 2 : 1;                                      // operand tiles per side: hi [, lo]
    constexpr int A_BYTES = BM * ROWB, B_BYTES = C::B_ROWS / 2 * ROWB;          // per CTA: own A rows, half of B
    constexpr int STAGE_BYTES = tc2_stage_bytes<MODE, NS>();                    // [A hi][A lo][B hi][B lo]
    constexpr int KCH_U = MODE == 0 ? ECORD_DIM_MSG / BKK : 0;                  // k-chunks of the GRU's input part (u, Wu)
    constexpr int NCHUNK = KCH_U + H / BKK;
    constexpr int B_PLANE = MODE == 0 ? 3 * H : ECORD_DIM_P1;                   // rows between the hi and lo parts of B
    constexpr int NST = stages2<MODE, NS>();
    constexpr int TMEM_COLS = 2 * C::ACC_COLS;          // 512 for both modes

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NST + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[2][4][64];       // GRU gate biases of the tile in flight (r, z combined; hn; in)

    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // epilogue staging tiles: behind the stages, or (MODE 2: one tile per CTA) on top of them
    uint8_t *stage_gen = smem_raw + (tiles - smem_u32(smem_raw)) + (MODE == 2 ? 0 : NST * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NST]);
    const uint32_t tfull0 = smem_u32(&bars[2 * NST]), tempty0 = smem_u32(&bars[2 * NST + 2]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 2 * (kEpiThreads / 32)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                    // both CTAs' barriers are initialised before anything remote touches them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    if (prm.debug & 8) num_tiles = 0;       // diagnostics: prologue + teardown only
    pdl_launch_dependents();
    pdl_wait();                             // everything below reads / writes step data
    const int cur = ((prm.step_base ? *prm.step_base : 0) + prm.step_off) & 1;

    if (warp == 0 || warp >= kThreads / 32) {
        // ===================== TMA producers (both CTAs): warp 0 and the warps behind the epilogue =====================
        // the whole warp walks the loop converged (uniform registers, no divergence bookkeeping); one elected lane issues
        {
            const CUtensorMap *mapA = (MODE == 0) ? (cur ? &maps.a1 : &maps.a0) : &maps.a0;
            if (lane == 0) {
                tmap_prefetch(mapA); tmap_prefetch(&maps.b[0]);
                if (MODE == 0) { tmap_prefetch(&maps.u); tmap_prefetch(&maps.wum[0]); tmap_prefetch(&maps.wui[0]); }
            }
            const uint32_t lfull0 = mapa_u32(full0, 0);            // the leader's full barriers
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - kThreads / 32 + 1);     // this producer's turn in the ring
            uint32_t it = 0;
            int m0p, nb, kw;
            for (int ti = 0; tc2_tile<MODE>(ti, cluster_id, num_clusters, num_tiles, nblk, prm, m0p, nb, kw); ++ti) {
                const int m0 = m0p + (int)rank * BM;
                const int wsel = kw == 4 ? 0 : 1;                              // which set of weight maps (box heights)
                const int n_main = MODE == 0 ? 24 * kw : C::N_MAIN / 2;        // B rows this CTA fetches per plane (half of the MMA's N)
                const int n_in = 8 * kw;                                       // GRU u-chunks: its half of the W_in rows
                for (int c = 0; c < NCHUNK; ++c, ++it) {
                    if (it % kProducers != me) continue;
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);            // local: released by the leader's multicast commit
                    const uint32_t a_dst = tiles + s * STAGE_BYTES, b_dst = a_dst + PARTS * A_BYTES;
                    const bool first = MODE == 0 && c < KCH_U;
                    const int kc = (first ? c : c - KCH_U) * BKK;
                    const int b_row = MODE == 0 ? nb * 48 + (int)rank * n_main : nb * C::N_MAIN + (int)rank * n_main;
                    if (elect_one()) {
                        // diagnostics: 32 = hi planes only (half the bytes), 64 = every cluster fetches tile (0, 0) (hot lines)
                        const int parts = (prm.debug & 32) ? 1 : PARTS;
                        const int m0d = (prm.debug & 64) ? (int)rank * BM : m0, b_rowd = (prm.debug & 64) ? (int)rank * n_main : b_row;
                        if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * parts * (A_BYTES + (n_main + (first ? n_in : 0)) * ROWB));
#pragma unroll
                        for (int part = 0; part < PARTS; ++part) {
                            if (part >= parts) break;
                            if (first) {
                                tma_load_2d_pair(a_dst + part * A_BYTES, &maps.u, lfull0 + 8 * s, kc, m0d + part * prm.rows_pad);
                                tma_load_2d_pair(b_dst + part * B_BYTES, &maps.wum[wsel], lfull0 + 8 * s, kc, b_rowd + part * B_PLANE);
                                tma_load_2d_pair(b_dst + part * B_BYTES + n_main * ROWB, &maps.wui[wsel], lfull0 + 8 * s, kc,
                                                 nb * 16 + (int)rank * n_in + part * H);
                            } else {
                                tma_load_2d_pair(a_dst + part * A_BYTES, mapA, lfull0 + 8 * s, kc, m0d + part * prm.rows_pad);
                                tma_load_2d_pair(b_dst + part * B_BYTES, &maps.b[wsel], lfull0 + 8 * s, kc, b_rowd + part * B_PLANE);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {                    // converged warp, elected lane issues (see the producer)
            uint32_t it = 0, lt = 0;
            int m0p, nb, kw;
            for (int ti = 0; tc2_tile<MODE>(ti, cluster_id, num_clusters, num_tiles, nblk, prm, m0p, nb, kw); ++ti, ++lt) {
                const uint32_t ab = lt & 1u, aph = (lt >> 1) & 1u;
                mbar_wait(tempty0 + 8 * ab, aph ^ 1u);            // both CTAs' epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + ab * C::ACC_COLS;
                const uint32_t idesc = umma_idesc(2 * BM, MODE == 0 ? 48 * kw : C::N_MAIN);
                // GRU u-chunks: the W_in rows (N = 16 kw) sit behind this CTA's Wu_main rows and feed accumulator columns 48 kw ..
                const uint32_t idesc_in = umma_idesc(2 * BM, 16 * kw), acc_in = acc + 48 * kw, in_off = 24 * kw * ROWB;
                for (int c = 0; c < NCHUNK; ++c, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_src = tiles + s * STAGE_BYTES, b_src = a_src + PARTS * A_BYTES;
                    const bool first = MODE == 0 && c < KCH_U;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BKK / UMMA_K; ++k) {
                            const uint32_t ko = k * UMMA_K * 2;
                            const uint32_t fresh = (c > 0 || k > 0) ? 1u : 0u;
                            if (prm.debug & 1) continue;
                            if (NS == 1) {
                                umma_bf16_pair(acc, umma_desc(a_src + ko), umma_desc(b_src + ko), idesc, fresh);
                                if (first) umma_bf16_pair(acc_in, umma_desc(a_src + ko), umma_desc(b_src + in_off + ko), idesc_in, fresh);
                            } else {       // hi.hi + hi.lo + lo.hi
                                const uint64_t ah = umma_desc_sw64(a_src + ko), al = umma_desc_sw64(a_src + A_BYTES + ko);
                                const uint64_t bh = umma_desc_sw64(b_src + ko), bl = umma_desc_sw64(b_src + B_BYTES + ko);
                                umma_bf16_pair(acc, ah, bh, idesc, fresh);
                                umma_bf16_pair(acc, ah, bl, idesc, 1u);
                                umma_bf16_pair(acc, al, bh, idesc, 1u);
                                if (first) {
                                    const uint64_t ih = umma_desc_sw64(b_src + in_off + ko), il = umma_desc_sw64(b_src + B_BYTES + in_off + ko);
                                    umma_bf16_pair(acc_in, ah, ih, idesc_in, fresh);
                                    umma_bf16_pair(acc_in, ah, il, idesc_in, 1u);
                                    umma_bf16_pair(acc_in, al, ih, idesc_in, 1u);
                                }
                            }
                        }
                        umma_commit_pair(empty0 + 8 * s);         // frees the stage in both CTAs
                        if (c == NCHUNK - 1) umma_commit_pair(tfull0 + 8 * ab);            // accumulator complete, both CTAs
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..17, both CTAs): the single-CTA kernel's math, global traffic staged =====================
        const int q = warp & 3;
        const int part = (warp - 2) >> 2;
        const uint32_t ltempty0 = mapa_u32(tempty0, 0);            // the leader's tmem_empty barriers
        uint32_t lt = 0;
        int m0p, nb, kw;
        for (int ti = 0; tc2_tile<MODE>(ti, cluster_id, num_clusters, num_tiles, nblk, prm, m0p, nb, kw); ++ti, ++lt) {
            const int m0 = m0p + (int)rank * BM;
            const uint32_t ab = lt & 1u, aph = (lt >> 1) & 1u;
            const int row = m0 + q * 32 + lane;
            const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + ab * C::ACC_COLS;
            if (MODE == 0) {
                // Global traffic of the epilogue goes through a shared-memory staging tile so that every LDG / STG covers
                // whole 128-byte lines (a warp <-> 2 or 4 contiguous row pieces).  With lane <-> row (the TMEM layout) each
                // 16-byte access of a warp hit 32 different lines and the L1 data pipe -- shared with the tensor core's
                // operand reads and the TMA writes -- ran at 36 % on those wavefronts alone.
                // A tile is kw sub-blocks of 16 units (kw = 4, 2 or 1): epilogue part p owns sub-block p (parts >= kw idle
                // through the math but keep the barrier protocol), the staging tile uses its first 16 kw columns.
                const int e = threadIdx.x - 64, ew = e >> 5;          // epilogue thread / warp index 0..511 / 0..15
                const int rloc = q * 32 + lane;                        // row of the tile this thread owns in TMEM
                const int unit0 = nb * 16, tw = kw * 16;               // first hidden unit and width of the tile
                const bool active = part < kw;
                float (*Hs)[kHsPitch] = reinterpret_cast<float (*)[kHsPitch]>(stage_gen);
                // NS = 1: bf16 images of h' and tanh(h') are staged as well; NS = 3: only h' (fp32) is staged, the store pass takes
                // tanh(h') and splits both into bf16 hi / lo pairs
                __nv_bfloat16 (*Bs)[kBsPitch] = reinterpret_cast<__nv_bfloat16 (*)[kBsPitch]>(stage_gen + BM * kHsPitch * 4);
                __nv_bfloat16 (*Ts)[kBsPitch] = Bs + BM;
                const size_t cur_off = (size_t)cur * prm.rows_pad, nxt_off = (size_t)(1 - cur) * prm.rows_pad;
                // (1) old state of the tile, coalesced: warp ew takes rows 8 ew .. 8 ew + 7, two 256-byte row pieces per instruction
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 8 * ew + 2 * i + (lane >> 4), c4 = lane & 15;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (m0 + r < prm.rows && 4 * c4 < tw && !(prm.debug & 128))
                        v = *reinterpret_cast<const float4 *>(prm.hidden + (cur_off + m0 + r) * H + unit0 + 4 * c4);
                    *reinterpret_cast<float4 *>(&Hs[r][4 * c4]) = v;
                }
                {   // this tile's 4 x 16 kw gate biases -> shared memory (read back as warp-uniform broadcasts)
                    const int gsel = e >> 6, j = unit0 + (e & 63);
                    if (e < 256 && (e & 63) < tw)
                        bias_s[ab][gsel][e & 63] = gsel == 0 ? prm.bih[j] + prm.bhh[j]
                                                 : gsel == 1 ? prm.bih[H + j] + prm.bhh[H + j]
                                                 : gsel == 2 ? prm.bhh[2 * H + j] : prm.bih[2 * H + j];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // staged state + biases visible
                float hv[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(&hv[i]) = *reinterpret_cast<const float4 *>(&Hs[rloc][part * 16 + i]);
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // every thread holds its values: Hs is free again
                mbar_wait(tfull0 + 8 * ab, aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (active && !(prm.debug & 2)) {
                    // accumulator columns: r | z | hn of sub-block `part` at 48 part, its `in` columns behind all main columns
                    float ar[16], az[16], ahn[16], ain[16];
                    tmem_ld16(tbase + 48 * part, ar);
                    tmem_ld16(tbase + 48 * part + 16, az);
                    tmem_ld16(tbase + 48 * part + 32, ahn);
                    tmem_ld16(tbase + 48 * kw + 16 * part, ain);
                    tmem_ld_wait();
                    if (NS == 1) {
                        __align__(16) __nv_bfloat16 ob[16], ot[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int jl = part * 16 + i;           // unit within the tile
                            const float r = sigmoid_fast(ar[i] + bias_s[ab][0][jl]);
                            const float z = sigmoid_fast(az[i] + bias_s[ab][1][jl]);
                            const float n = tanh_fast(fmaf(ahn[i] + bias_s[ab][2][jl], r, ain[i] + bias_s[ab][3][jl]));
                            const float hn = fmaf(hv[i] - n, z, n);
                            hv[i] = hn;
                            ob[i] = __float2bfloat16(hn);
                            ot[i] = __float2bfloat16(tanh_fast(hn));
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(&Hs[rloc][part * 16 + i]) = *reinterpret_cast<float4 *>(&hv[i]);
                        *reinterpret_cast<uint4 *>(&Bs[rloc][part * 16]) = *reinterpret_cast<uint4 *>(&ob[0]);
                        *reinterpret_cast<uint4 *>(&Bs[rloc][part * 16 + 8]) = *reinterpret_cast<uint4 *>(&ob[8]);
                        *reinterpret_cast<uint4 *>(&Ts[rloc][part * 16]) = *reinterpret_cast<uint4 *>(&ot[0]);
                        *reinterpret_cast<uint4 *>(&Ts[rloc][part * 16 + 8]) = *reinterpret_cast<uint4 *>(&ot[8]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int jl = part * 16 + i;
                            const float r = gate_sigmoid(ar[i] + bias_s[ab][0][jl]);
                            const float z = gate_sigmoid(az[i] + bias_s[ab][1][jl]);
                            const float n = gate_tanh(fmaf(ahn[i] + bias_s[ab][2][jl], r, ain[i] + bias_s[ab][3][jl]));
                            hv[i] = fmaf(hv[i] - n, z, n);
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4 *>(&Hs[rloc][part * 16 + i]) = *reinterpret_cast<float4 *>(&hv[i]);
                    }
                }
                // this warp has read its part of the accumulator: hand it back to the MMA issuer before the stores
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(ltempty0 + 8 * ab);
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // the staged tiles are complete
                // (2) coalesced stores: h' fp32 (two 256-byte row pieces per instruction), h' and tanh(h') bf16 (four 128-byte pieces)
                float *hnew = const_cast<float *>(prm.hidden);
                if (NS == 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = 8 * ew + 2 * i + (lane >> 4), c4 = lane & 15;
                        if (m0 + r < prm.rows && 4 * c4 < tw)
                            *reinterpret_cast<float4 *>(hnew + (nxt_off + m0 + r) * H + unit0 + 4 * c4) = *reinterpret_cast<const float4 *>(&Hs[r][4 * c4]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = 8 * ew + 4 * i + (lane >> 3), c8 = lane & 7;
                        if (m0 + r < prm.rows && 8 * c8 < tw) {
                            *reinterpret_cast<uint4 *>(prm.hidden_bf16 + (nxt_off + m0 + r) * H + unit0 + 8 * c8) = *reinterpret_cast<const uint4 *>(&Bs[r][8 * c8]);
                            *reinterpret_cast<uint4 *>(prm.th_bf16 + (size_t)(m0 + r) * H + unit0 + 8 * c8) = *reinterpret_cast<const uint4 *>(&Ts[r][8 * c8]);
                        }
                    }
                } else {
                    // hidden_bf16 [parity][hi|lo][rows_pad][H], th_bf16 [hi|lo][rows_pad][H]: a thread splits four values, a
                    // half-warp covers one 128-byte row piece of each of the four bf16 planes
                    __nv_bfloat16 *hb_hi = prm.hidden_bf16 + (size_t)(1 - cur) * 2 * prm.rows_pad * H;
                    __nv_bfloat16 *hb_lo = hb_hi + (size_t)prm.rows_pad * H;
                    __nv_bfloat16 *tb_lo = prm.th_bf16 + (size_t)prm.rows_pad * H;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = 8 * ew + 2 * i + (lane >> 4), c4 = lane & 15;
                        if (m0 + r < prm.rows && 4 * c4 < tw && !(prm.debug & 128)) {
                            const float4 hv4 = *reinterpret_cast<const float4 *>(&Hs[r][4 * c4]);
                            const float4 tv4 = make_float4(gate_tanh(hv4.x), gate_tanh(hv4.y), gate_tanh(hv4.z), gate_tanh(hv4.w));   // projection input
                            const size_t o = (size_t)(m0 + r) * H + unit0 + 4 * c4;
                            *reinterpret_cast<float4 *>(hnew + nxt_off * H + o) = hv4;
                            uint2 hi, lo;
                            split_bf16x4(hv4, hi, lo);
                            *reinterpret_cast<uint2 *>(hb_hi + o) = hi;
                            *reinterpret_cast<uint2 *>(hb_lo + o) = lo;
                            split_bf16x4(tv4, hi, lo);
                            *reinterpret_cast<uint2 *>(prm.th_bf16 + o) = hi;
                            *reinterpret_cast<uint2 *>(tb_lo + o) = lo;
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // staging tiles are free for the next tile
            } else {
                // PROJ: bias added in registers, the tile staged in shared memory and stored with whole-line STGs (see above)
                constexpr int PC = C::TILE_N / 4;                 // 44 columns per warp: 11 groups of 4
                constexpr int kYsPitch = C::TILE_N + 4;           // 180 floats: 16-byte aligned rows, 4-way STS.128
                const int e = threadIdx.x - 64;
                const int rloc = q * 32 + lane;
                float (*Ys)[kYsPitch] = reinterpret_cast<float (*)[kYsPitch]>(stage_gen);
                const int c00 = nb * C::TILE_N + part * PC;
                mbar_wait(tfull0 + 8 * ab, aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float acc[PC];
#pragma unroll
                for (int jj = 0; jj < PC; jj += 4) tmem_ld4(tbase + part * PC + jj, acc + jj);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < PC; jj += 4) {
                    const int j = c00 + jj;
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < ECORD_DIM_P1) b4 = *reinterpret_cast<const float4 *>(prm.p1_b + j);   // the last block's tail columns do not exist
                    *reinterpret_cast<float4 *>(&Ys[rloc][part * PC + jj]) =
                        make_float4(acc[jj] + b4.x, acc[jj + 1] + b4.y, acc[jj + 2] + b4.z, acc[jj + 3] + b4.w);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(ltempty0 + 8 * ab);
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // the staged tile is complete
                constexpr int kPieces = C::TILE_N / 4;            // float4 pieces per row
                for (int idx = e; idx < BM * kPieces; idx += kEpiThreads) {
                    const int r = idx / kPieces, c4 = idx - r * kPieces;
                    const int col = nb * C::TILE_N + 4 * c4;
                    if (m0 + r < prm.rows && col < ECORD_DIM_P1)
                        *reinterpret_cast<float4 *>(prm.y1 + (size_t)(m0 + r) * ECORD_DIM_P1 + col) = *reinterpret_cast<const float4 *>(&Ys[r][4 * c4]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // free for the next tile
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                    // nothing remote may still target this CTA's barriers / shared memory
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
template <int MODE, int NS>
constexpr int tc2_smem_bytes() {
    return stages2<MODE, NS>() * tc2_stage_bytes<MODE, NS>() + 1024 +
           (MODE == 0 ? (NS == 3 ? BM * kHsPitch * 4 : kStageTileBytes) : MODE == 1 ? BM * (TcCfg<MODE>::TILE_N + 4) * 4 : 0);
}
static_assert(tc2_smem_bytes<0, 3>() <= 227 * 1024 && tc2_smem_bytes<1, 3>() <= 227 * 1024 && tc2_smem_bytes<2, 3>() <= 227 * 1024 &&
              tc2_smem_bytes<2, 3>() - 1024 >= BM * (TcCfg<2>::TILE_N + 4) * 4, "shared-memory budget of the pair kernels");

// ---------------------------------------------------------------------------------------------
// weight packing (device): fp32 PyTorch layout -> bf16 tile-permuted, plus the folded encoder constants
// ---------------------------------------------------------------------------------------------
// bf16 hi part and bf16 lo remainder of x (x ~ hi + lo to 2^-17 relative)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16(x);
    lo = __float2bfloat16(x - __bfloat162float(hi));
}
__global__ void k_pack_weights(ecord_weights w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_wh = (long long)3072 * H, n_wu = (long long)4096 * 64, n_p1 = (long long)512 * H, n_p2 = 32 * 512;
    // layouts: gru_w_bf16 = [Wh hi | Wh lo | Wu_main hi | Wu_main lo | Wu_in hi | Wu_in lo], p1_w_bf16 = [W1 hi | W1 lo | W2 hi | W2 lo]
    __nv_bfloat16 *wh = reinterpret_cast<__nv_bfloat16 *>(w.gru_w_bf16);
    __nv_bfloat16 *wu = wh + 2 * n_wh;
    __nv_bfloat16 *p1 = reinterpret_cast<__nv_bfloat16 *>(w.p1_w_bf16);
    __nv_bfloat16 *p2 = p1 + 2 * n_p1;
    if (i < n_wh) {                               // Wh pack: row = sb*48 + gate*16 + jj  <-  W_hh[gate*1024 + sb*16 + jj]
        const int row = (int)(i / H), k = (int)(i % H);
        const int sb = row / 48, gate = (row % 48) / 16, jj = row % 16;
        split_bf16(w.gru_whh[(size_t)(gate * H + sb * 16 + jj) * H + k], wh[i], wh[n_wh + i]);
    } else if (i < n_wh + n_wu) {                 // Wu_main pack [3072][64]: row = sb*48 + slot*16 + jj, slots r, z, (zero);
        const long long t = i - n_wh;             // then Wu_in [1024][64]: row = unit  <-  W_ih[2*1024 + unit]
        const long long n_main = (long long)3 * H * 64;
        float v = 0.0f;
        __nv_bfloat16 *hi_p, *lo_p;
        if (t < n_main) {
            const int row = (int)(t / 64), k = (int)(t % 64);
            const int sb = row / 48, slot = (row % 48) / 16, jj = row % 16;
            if (slot != 2) v = w.gru_wih[(size_t)(slot * H + sb * 16 + jj) * 64 + k];
            hi_p = wu + t; lo_p = wu + n_main + t;
        } else {
            const long long q = t - n_main;
            v = w.gru_wih[(size_t)(2 * H + (int)(q / 64)) * 64 + (int)(q % 64)];
            hi_p = wu + 2 * n_main + q; lo_p = wu + 2 * n_main + (long long)H * 64 + q;
        }
        split_bf16(v, *hi_p, *lo_p);
    } else if (i < n_wh + n_wu + n_p1) {
        const long long t = i - n_wh - n_wu;
        split_bf16(w.p1_w[t], p1[t], p1[n_p1 + t]);
    } else if (i < n_wh + n_wu + n_p1 + n_p2) {   // proj_rnn_to_gnn.4 weight [32][512], behind W1
        const long long t = i - n_wh - n_wu - n_p1;
        split_bf16(w.p2_w[t], p2[t], p2[n_p2 + t]);
    }
}

// Operands of the tensor-core Q kernel (layout: ecord_weights.q_tc).  One block of 64 threads, thread = channel c.
// Column means are taken in double and subtracted exactly as ecord_fold_q_host() does, so that the host-side
// constants (LC, Wb, C^T C) and these device-side operands describe the same centred matrices.
__global__ void k_pack_qtc(ecord_weights w) {
    constexpr int Q = ECORD_DIM_Q;
    __shared__ float raw[Q][20];      // [c][0..15] Wq[c][32+k] | [16..18] C[c][j] | [19] Wq[:,48:64] b_enc
    __shared__ double mean[20];
    const int c = threadIdx.x;
    for (int k = 0; k < 16; ++k) raw[c][k] = w.q1_w[c * Q + 32 + k];
    for (int j = 0; j < 3; ++j) {
        float a = 0.0f;
        for (int i = 0; i < 16; ++i) a = fmaf(w.q1_w[c * Q + 48 + i], w.enc_w[i * 3 + j], a);
        raw[c][16 + j] = a;
    }
    {
        float a = 0.0f;
        for (int i = 0; i < 16; ++i) a = fmaf(w.q1_w[c * Q + 48 + i], w.enc_b[i], a);
        raw[c][19] = a;
    }
    __syncthreads();
    if (c < 20) {
        double m = 0.0;
        for (int r = 0; r < Q; ++r) m += (double)raw[r][c];
        mean[c] = m / Q;
    }
    __syncthreads();
    const float gam = w.qln_w[c];
    float *out = w.q_tc;
    for (int k = 0; k < 16; ++k) {
        const float wg = (float)((double)raw[c][k] - mean[k]);
        out[1280 + c * 16 + k] = wg;
        out[c * 16 + k] = 0.0f;
    }
    for (int j = 0; j < 4; ++j) {
        const float cc = (float)((double)raw[c][16 + j] - mean[16 + j]);
        out[2304 + c * 4 + j] = cc;
        out[2560 + c * 4 + j] = 0.0f;
    }
    // shared-memory image of the B operand of k_q_tc: three [80][16] fp16 blocks in the K-major SWIZZLE_32B layout
    //   0: hi(gamma_c Wg[c][k])   1: lo(...)   2: W2 = [C0h C1h C2h C1h . . C2h 0 | C0l C1l C2l 0 0 0 0 0], C = gamma_c Cc[c]
    // (hi = fp16(value), lo = fp16(value - hi)).  The per-trajectory entries (row 64 everywhere, W2 halves 4 and 5)
    // stay zero here.
    __half *img = reinterpret_cast<__half *>(out + 2816);
    for (int i = c; i < 3 * 80 * 16; i += ECORD_DIM_Q) img[i] = __float2half_rn(0.0f);
    __syncthreads();
    auto at = [&](int blk, int r, int k) -> __half & {
        return img[blk * 1280 + (r * 32 + ((((k >> 3) ^ (r >> 2)) & 1) << 4) + (k & 7) * 2) / 2];
    };
    for (int k = 0; k < 16; ++k) {
        const float v = gam * out[1280 + c * 16 + k];
        const __half hi = __float2half_rn(v);
        at(0, c, k) = hi;
        at(1, c, k) = __float2half_rn(v - __half2float(hi));
    }
    for (int j = 0; j < 3; ++j) {
        const float v = gam * out[2304 + c * 4 + j];
        const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
        at(2, c, j) = hi;
        at(2, c, 8 + j) = lo;
        if (j == 1) at(2, c, 3) = hi;       // against f1_lo
        if (j == 2) at(2, c, 6) = hi;       // against f2_lo
    }
}

}  // namespace

// Shared-memory image of the policy tail's constants (layout: ecord_weights.q_tc [4736, 10148)); runs after k_pack_qtc.
__global__ void k_pack_tail(ecord_weights w) {
    constexpr int P1 = ECORD_DIM_P1, ND = ECORD_DIM_NODE, QC = ECORD_DIM_Q;
    float *img = w.q_tc + 4736;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P1) { img[i] = w.ln1_w[i]; img[P1 + i] = w.ln1_b[i]; }
    if (i < ND * ND) img[2 * P1 + i] = w.v1_w[i];
    if (i < QC * ND) img[2 * P1 + ND * ND + i] = w.q1_w[(i / ND) * QC + (i % ND)];
    float *wg = img + 2 * P1 + ND * ND + QC * ND;
    if (i < QC * 16) wg[i] = w.q_tc[1280 + i];
    if (i < QC * 4) wg[QC * 16 + i] = w.q_tc[2304 + i];
    float *colsum = wg + QC * 16 + QC * 4;
    if (i <= ND) {                              // serial over the channels: the order the tail used to sum in
        float s = 0.0f;
        for (int c = 0; c < QC; ++c) s += i < ND ? w.q1_w[c * QC + i] : w.q1_b[c];
        colsum[i] = s;
    }
    if (i > ND && i < ND + 4) colsum[i] = 0.0f;
}

extern "C" int ecord_pack_weights(const ecord_weights *w, void *stream) {
    ECORD_RET_IF(!w || !w->gru_w_bf16 || !w->p1_w_bf16 || !w->gru_whh || !w->gru_wih || !w->p1_w || !w->p2_w, ECORD_E_NULL);
    const long long n = (long long)3072 * H + (long long)4096 * 64 + (long long)512 * H + 32 * 512;      // one thread per weight: hi and lo
    k_pack_weights<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(*w);
    ECORD_LAUNCH_CHECK();
    if (w->q_tc) {
        ECORD_RET_IF(!w->q1_w || !w->enc_w || !w->enc_b || !w->qln_w, ECORD_E_NULL);
        k_pack_qtc<<<1, ECORD_DIM_Q, 0, as_stream(stream)>>>(*w);
        ECORD_LAUNCH_CHECK();
        ECORD_RET_IF(!w->ln1_w || !w->ln1_b || !w->v1_w || !w->q1_b, ECORD_E_NULL);
        k_pack_tail<<<ceil_div(ECORD_DIM_Q * ECORD_DIM_NODE, 256), 256, 0, as_stream(stream)>>>(*w);
        ECORD_LAUNCH_CHECK();
    }
    return 0;
}

// Host-side fold of the node-feature encoder into mlp_out[0] (all arguments are HOST arrays); layout of `out`
// is documented at ecord_weights.q_fold:
//   C[c][k] = sum_i Wq[c][48+i] W_enc[i][k]   (64 x 3), stored as planes C0|C1|C2, then gamma|beta|w_o, then the
//   channel-centred C planes, LC_k = sum_c w_o[c] gamma[c] Cc[c][k], and sum_c w_o[c] beta[c].
extern "C" int ecord_fold_q_host(const float *q1_w, const float *enc_w, const float *qln_w, const float *qln_b,
                                 const float *q2_w, float *out) {
    ECORD_RET_IF(!q1_w || !enc_w || !qln_w || !qln_b || !q2_w || !out, ECORD_E_NULL);
    const int Q = ECORD_DIM_Q;
    double mean[3] = {0, 0, 0};
    for (int c = 0; c < Q; ++c) {
        for (int k = 0; k < 3; ++k) {
            float a = 0.0f;
            for (int i = 0; i < 16; ++i) a = fmaf(q1_w[c * Q + 48 + i], enc_w[i * 3 + k], a);
            out[k * Q + c] = a;
            mean[k] += a;
        }
        out[3 * Q + c] = qln_w[c];
        out[4 * Q + c] = qln_b[c];
        out[5 * Q + c] = q2_w[c];
    }
    double wb = 0.0;
    for (int k = 0; k < 3; ++k) {
        double lc = 0.0;
        for (int c = 0; c < Q; ++c) {
            const float cc = (float)((double)out[k * Q + c] - mean[k] / Q);
            out[384 + k * Q + c] = cc;
            lc += (double)q2_w[c] * (double)qln_w[c] * (double)cc;
        }
        out[576 + k] = (float)lc;
    }
    for (int c = 0; c < Q; ++c) wb += (double)q2_w[c] * (double)qln_b[c];
    out[579] = (float)wb;
    int o = 580;
    for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b) {          // C^T C of the centred planes: 00 01 02 11 12 22
            double t = 0.0;
            for (int c = 0; c < Q; ++c) t += (double)out[384 + a * Q + c] * (double)out[384 + b * Q + c];
            out[o++] = (float)t;
        }
    out[586] = out[587] = 0.0f;
    return 0;
}

int tc_init_attrs() {
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<0, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<0, 1>()));
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<1, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<1, 1>()));
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<2, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<2, 1>()));
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<0, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<0, 3>()));
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<1, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<1, 3>()));
    ECORD_CUDA(cudaFuncSetAttribute((k_gemm_tc2<2, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem_bytes<2, 3>()));
    return 0;
}

// The GRU's tile schedule (tc2_tile): `full` rounds of 64-unit tiles over the clusters, then the leftover 64-unit tiles cut
// into kn-sub-block ones.  kn minimises the makespan under relative tile costs of 1 / 0.85 / 0.7: a narrow tile fetches the
// same A bytes as a 64-unit one (22 and 19 KB per k-chunk and SM against 28), and operand delivery, not the MMAs, bounds the
// kernel, so halving the width saves far less than half the time.  ECORD_GRU_NARROW = 4 | 2 | 1 forces the width (A/B tests).
struct GruSched { int clusters, full, kn, narrow; };
static GruSched gru_schedule(int tiles64) {
    static const int forced = [] { const char *e = getenv("ECORD_GRU_NARROW"); return e ? atoi(e) : 0; }();
    const int pairs = kNumSMs / 2;
    GruSched best{};
    double best_cost = 1e30;
    for (int kn = 4; kn >= 1; kn >>= 1) {
        if (forced && kn != forced) continue;
        const int per = 4 / kn;
        const double cost = kn == 4 ? 1.0 : kn == 2 ? 0.85 : 0.7;
        GruSched s;
        s.kn = kn;
        if (tiles64 >= pairs) { s.clusters = pairs; s.full = tiles64 / pairs; s.narrow = (tiles64 % pairs) * per; }
        else { s.full = 0; s.narrow = tiles64 * per; s.clusters = s.narrow < pairs ? s.narrow : pairs; }
        const double make = s.full + ceil_div(s.narrow, s.clusters) * cost;
        if (make < best_cost - 1e-9) { best_cost = make; best = s; }
    }
    return best;
}

// split != 0: ECORD_GEMM_BF16X3 (bf16 hi / lo operands, three products; CTA-pair kernels only)
// Diagnostics / tests: the tile schedule the GRU kernel would use for a batch of `rows_pad` rows (a multiple of 256), as
// (first row, first hidden unit, units) triples in out[3 * i], cluster by cluster.  Returns the number of tiles, or a negative
// error; *num_clusters receives the CTA pairs of the launch.  No GPU needed.
extern "C" int ecord_gru_tile_schedule(int32_t rows_pad, int32_t *out, int32_t max_tiles, int32_t *num_clusters) {
    if (rows_pad <= 0 || rows_pad % (2 * BM) != 0 || !out || max_tiles < 0) return ECORD_E_SIZE;
    const GruSched sc = gru_schedule((H / 64) * (rows_pad / (2 * BM)));
    if (num_clusters) *num_clusters = sc.clusters;
    int n = 0;
    for (int c = 0; c < sc.clusters; ++c) {
        int m0, nb, k;
        for (int i = 0; gru_tile(i, c, sc.clusters, sc.full, sc.kn, sc.narrow, m0, nb, k); ++i, ++n) {
            if (n < max_tiles) { out[3 * n] = m0; out[3 * n + 1] = nb * 16; out[3 * n + 2] = k * 16; }
        }
    }
    return n;
}

int launch_gru_tc_ex(const ecord_weights *w, const ecord_policy *p, const int32_t *step_base, int step_off, int split, cudaStream_t st) {
    PFN_encodeTiled enc;
    int rc = get_encoder(&enc);
    if (rc) return rc;
    const int Mp = p->rows_pad;
    ECORD_RET_IF(Mp % (2 * BM) != 0, ECORD_E_SIZE);                   // CTA pairs: 256-row tiles
    const int np = split ? 2 : 1;                                      // planes (hi [, lo]) the activation maps cover
    const __nv_bfloat16 *wh = reinterpret_cast<const __nv_bfloat16 *>(w->gru_w_bf16);
    const __nv_bfloat16 *wum = wh + (size_t)2 * 3072 * H, *wui = wum + (size_t)2 * 3072 * 64;
    const __nv_bfloat16 *hb = reinterpret_cast<const __nv_bfloat16 *>(p->hidden_bf16);
    const uint32_t bc = split ? 32 : 64;                               // k-chunk width (tc2_bk)
    const GruSched sc = gru_schedule((H / 64) * (Mp / (2 * BM)));
    TcMaps m;
    if ((rc = make_map(enc, &m.u, p->u_bf16, (uint64_t)np * Mp, 64, BM, bc))) return rc;
    if ((rc = make_map(enc, &m.a0, hb, (uint64_t)np * Mp, H, BM, bc))) return rc;
    if ((rc = make_map(enc, &m.a1, hb + (size_t)np * Mp * H, (uint64_t)np * Mp, H, BM, bc))) return rc;
    for (int i = 0; i < 2; ++i) {                                      // box heights: this CTA's half of the tile's rows
        const int kw = i == 0 ? 4 : sc.kn;
        if ((rc = make_map(enc, &m.b[i], wh, 2 * 3072, H, 24 * kw, bc))) return rc;
        if ((rc = make_map(enc, &m.wum[i], wum, 2 * 3072, 64, 24 * kw, bc))) return rc;
        if ((rc = make_map(enc, &m.wui[i], wui, 2 * 1024, 64, 8 * kw, bc))) return rc;
    }
    TcParams prm{};
    prm.hidden = p->hidden; prm.hidden_bf16 = reinterpret_cast<__nv_bfloat16 *>(p->hidden_bf16);
    prm.th_bf16 = reinterpret_cast<__nv_bfloat16 *>(p->th_bf16);
    prm.bih = w->gru_bih; prm.bhh = w->gru_bhh;
    prm.rows = p->rows; prm.rows_pad = Mp; prm.step_base = step_base; prm.step_off = step_off;
    prm.sched_full = sc.full; prm.sched_kn = sc.kn; prm.sched_narrow = sc.narrow;
    { static const int dbg = [] { const char *e = getenv("ECORD_TC_DEBUG"); return e ? atoi(e) : 0; }(); prm.debug = dbg; }
    ECORD_RET_IF(prm.debug & 4, ECORD_E_UNSUPPORTED);      // the no-TMA diagnostic no longer terminates (it hung the GPU when tried)
    int clusters = sc.clusters;
    { const char *e = getenv("ECORD_TC_CLUSTERS"); if (e && atoi(e) > 0 && atoi(e) < clusters && sc.kn == 4 && sc.full == 0) clusters = atoi(e); }   // diagnostics
    if (split)
        ECORD_PDL_LAUNCH((k_gemm_tc2<0, 3>), dim3(2 * clusters), dim3(kThreads2), (size_t)tc2_smem_bytes<0, 3>(), st, m, prm, 0, 0);
    else
        ECORD_PDL_LAUNCH((k_gemm_tc2<0, 1>), dim3(2 * clusters), dim3(kThreads2), (size_t)tc2_smem_bytes<0, 1>(), st, m, prm, 0, 0);
    return 0;
}

int launch_proj1_tc(const ecord_weights *w, const ecord_policy *p, int split, cudaStream_t st) {
    PFN_encodeTiled enc;
    int rc = get_encoder(&enc);
    if (rc) return rc;
    const int Mp = p->rows_pad;
    ECORD_RET_IF(Mp % (2 * BM) != 0, ECORD_E_SIZE);
    TcMaps m;
    if ((rc = make_map(enc, &m.a0, p->th_bf16, (uint64_t)(split ? 2 : 1) * Mp, H, BM, split ? 32 : 64))) return rc;
    if ((rc = make_map(enc, &m.b[0], w->p1_w_bf16, 2 * 512, H, 88, split ? 32 : 64))) return rc;
    m.u = m.a1 = m.a0;                                     // (unused by the projection; valid descriptors all the same)
    m.b[1] = m.wum[0] = m.wum[1] = m.wui[0] = m.wui[1] = m.b[0];
    TcParams prm{};
    prm.p1_b = w->p1_b; prm.y1 = p->y1; prm.rows = p->rows; prm.rows_pad = Mp;
    const int nblk2 = 3, tiles2 = nblk2 * (Mp / (2 * BM));
    const int clusters = tiles2 < kNumSMs / 2 ? tiles2 : kNumSMs / 2;
    const dim3 grid(2 * clusters), block(kThreads2);
#define ECORD_PROJ_LAUNCH(MODE, NS) \
    ECORD_PDL_LAUNCH((k_gemm_tc2<MODE, NS>), grid, block, (size_t)tc2_smem_bytes<MODE, NS>(), st, m, prm, tiles2, nblk2)
    if (tiles2 <= clusters) {        // one tile per CTA pair: deep pipeline, staging on top of the stages
        if (split) ECORD_PROJ_LAUNCH(2, 3); else ECORD_PROJ_LAUNCH(2, 1);
    } else {
        if (split) ECORD_PROJ_LAUNCH(1, 3); else ECORD_PROJ_LAUNCH(1, 1);
    }
#undef ECORD_PROJ_LAUNCH
    return 0;
}
