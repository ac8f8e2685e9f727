// tail_tc.cu -- the per-(graph,trajectory) tail of the policy step with its 512 -> 32 contraction on tcgen05
// (throughput mode; the fp32 parity mode keeps k_policy_tail in policy.cu).
//
//   y1 [M][512] -> LN512 -> LReLU -> W2 (512 -> 32) -> LN32 -> LReLU = P ;  v = val_out(P) ;  A = Wq[:, :32] P + b_q
//   + the per-(g,t) operands of the fast / tensor-core Q kernels          (models/rnn_decoder.py:84-90,108-119,285-311)
//
// One CTA per block of R = 32 or 64 rows (the MMA still has M = 128: the unused A rows are whatever shared memory
// holds, their accumulator lanes are never read), so that a few thousand rows spread over all SMs:
//   phase 1  16 warps, R/16 rows each (both requested up front): LayerNorm(512) + LeakyReLU in registers (16 values
//            per lane), result rounded to bf16 and written to shared memory as the A operand of the MMA -- eight
//            [128][64] K-major SWIZZLE_128B sub-tiles, the layout TMA would produce
//   phase 2  W2 (bf16, eight [32][64] sub-tiles fetched by TMA during phase 1) x A: 32 tcgen05.mma (M=128, N=32, K=16),
//            accumulator [128][32] fp32 in TMEM
//   phase 3  R threads (thread <-> row <-> TMEM lane) read the accumulator, apply LayerNorm(32) and publish P, tanh(P)
//            and the channel mean of A in shared memory; then all 512 threads work, 512/R per row: each takes
//            32/(512/R) value-head units and 64/(512/R) Q channels (weights as warp-uniform broadcast LDS.128) and the
//            22 per-row sums are combined through shared memory
// The warp-per-row CUDA-core version spends its time streaming W2 (64 KB) out of shared memory once per row
// (5000 x 64 KB per step) and in shuffle reductions; here W2 is read by the tensor core once per CTA.
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {
constexpr int P1 = ECORD_DIM_P1;     // 512
constexpr int ND = ECORD_DIM_NODE;   // 32
constexpr int QC = ECORD_DIM_Q;      // 64
constexpr int kTtThreads = 512;
constexpr int A_SUB = BM * 128;      // bytes of one [128][64] bf16 sub-tile
constexpr int W_SUB = ND * 128;      // bytes of one [32][64] bf16 sub-tile
constexpr int kTailImgFloats = 2 * P1 + ND * ND + QC * ND + QC * 16 + QC * 4 + ND + 4;     // ecord_weights.q_tc [4736, ...)

struct TailTcPost {                  // written after the MMAs have completed: shares the memory of the W2 tiles
    float Ps[64][ND + 1], Ts[64][ND + 1];   // P and tanh(P) of the CTA's rows (padded: conflict-free row reads)
    float lo[64][ND + 1];            // NS = 3: accumulator rows 64.. (the lo-part products), handed to the rows' threads
    float meanA[64];
};
struct TailTcSmem {
    uint8_t act[8][A_SUB];           // 128 KB
    union {
        uint8_t w2[2][8][W_SUB];     // 32 KB (hi) + 32 KB (lo, NS = 3)
        TailTcPost post;
    };
    float ln1g[P1], ln1b[P1];
    float v1[ND * ND];               // val_out.1 weight [j][o]
    float q1[QC * ND];               // Wq[:, :32]  [c][o]
    float wg[QC * 16];               // centred Wq[:,32:48] [c][k]
    float cc[QC * 4];                // centred C [c][0..2], bconst[c]
    float colsum[ND + 4];            // sum_c Wq[c][o]; [32] = sum_c b_q[c]; padded: ln1g .. colsum is one 16-byte-granular image
    uint64_t bar_w, bar_mma, bar_c;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t split_h2(float x) {     // fp16 (hi | lo << 16), see qtc.cu
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}

// NS = 1: bf16 A operand and weights (ECORD_GEMM_BF16).  NS = 3 (ECORD_GEMM_BF16X3): the LN512 output is split into a bf16 hi
// part (A-operand rows 0..R-1) and its bf16 lo remainder (rows 64..64+R-1) of the SAME [128][512] operand, so one pass of
// MMAs against W2 hi yields hi.hi in accumulator rows 0.. and lo.hi in rows 64.., and a second pass against W2 lo adds
// hi.lo (and the negligible lo.lo); the two row halves are added through shared memory.  Exact tanh in the value head.
template <int NS>
__global__ void __launch_bounds__(kTtThreads, 1)
k_tail_tc(const __grid_constant__ CUtensorMap mapW2, ecord_weights w, ecord_policy p, int R) {
    extern __shared__ uint8_t smem_raw[];
    TailTcSmem &S = *reinterpret_cast<TailTcSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * R;
    const uint32_t bar_w = smem_u32(&S.bar_w), bar_mma = smem_u32(&S.bar_mma);

    const uint32_t bar_c = smem_u32(&S.bar_c);
    const bool img = w.q_tc != nullptr;          // the constants come as one packed image (ecord_pack_weights) ...
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_c, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tmap_prefetch(&mapW2);
        if (img) {
            mbar_expect_tx(bar_c, kTailImgFloats * 4);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(S.ln1g)), "l"(w.q_tc + 4736), "r"(kTailImgFloats * 4), "r"(bar_c) : "memory");
        }
        mbar_expect_tx(bar_w, (NS == 3 ? 2 : 1) * 8 * W_SUB);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) tma_load_2d(smem_u32(S.w2[0][kb]), &mapW2, bar_w, kb * 64, 0);
        if (NS == 3) {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) tma_load_2d(smem_u32(S.w2[1][kb]), &mapW2, bar_w, kb * 64, ND);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(32u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!img) {                                  // ... or, without q_tc, are gathered by the threads
        for (int i = tid; i < P1; i += kTtThreads) { S.ln1g[i] = w.ln1_w[i]; S.ln1b[i] = w.ln1_b[i]; }
        for (int i = tid; i < ND * ND; i += kTtThreads) S.v1[i] = w.v1_w[i];
        for (int i = tid; i < QC * ND; i += kTtThreads) S.q1[i] = w.q1_w[(i / ND) * QC + (i % ND)];
        if (tid <= ND) {
            float s = 0.0f;
            for (int c = 0; c < QC; ++c) s += tid < ND ? w.q1_w[c * QC + tid] : w.q1_b[c];
            S.colsum[tid] = s;
        }
    }
    __syncthreads();
    pdl_launch_dependents();
    if (img) mbar_wait(bar_c, 0);           // constants landed (async-proxy writes, made visible by the barrier's completion)
    pdl_wait();                             // y1 comes from the projection kernel

    // ---------------- phase 1: LN512 + LReLU -> bf16 A operand ----------------
    // (the next row's 2 KB are requested before the current row is reduced: a warp's 8 rows would otherwise pay
    //  8 serial memory latencies)
    {
        constexpr int NW = kTtThreads / 32;
        auto row_ptr = [&](int r) {
            return reinterpret_cast<const float4 *>(p.y1 + (long long)min(m0 + r, p.rows - 1) * P1);   // rows past the end:
        };                                                                       // the last row again, never stored
        const int nrow = R / NW;                         // 2 or 4 rows per warp, all requested before the first reduction
        float4 nx[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nrow) {
                const float4 *yr = row_ptr(warp + j * NW);
#pragma unroll
                for (int k = 0; k < 4; ++k) nx[j][k] = yr[lane + 32 * k];
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nrow) break;
            const int r = warp + j * NW;
            float x[16];
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                x[4 * k] = nx[j][k].x; x[4 * k + 1] = nx[j][k].y; x[4 * k + 2] = nx[j][k].z; x[4 * k + 3] = nx[j][k].w;
                s += (nx[j][k].x + nx[j][k].y) + (nx[j][k].z + nx[j][k].w);
            }
            const float mean = warp_sum(s) * (1.0f / P1);
            float ss = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] -= mean; ss = fmaf(x[i], x[i], ss); }
            const float rstd = rsqrtf(warp_sum(ss) * (1.0f / P1) + kLnEps);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * (lane + 32 * k);               // first of this lane's 4 consecutive elements
                const float4 g4 = *reinterpret_cast<const float4 *>(&S.ln1g[i]);
                const float4 b4 = *reinterpret_cast<const float4 *>(&S.ln1b[i]);
                const float y0 = lrelu(fmaf(x[4 * k + 0] * rstd, g4.x, b4.x)), y1 = lrelu(fmaf(x[4 * k + 1] * rstd, g4.y, b4.y));
                const float y2 = lrelu(fmaf(x[4 * k + 2] * rstd, g4.z, b4.z)), y3 = lrelu(fmaf(x[4 * k + 3] * rstd, g4.w, b4.w));
                const __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
                // K-major SWIZZLE_128B: sub-tile i/64, row r (128 B), 16-byte chunk ((i%64)/8) ^ (r%8)
                const uint32_t off = (uint32_t)(r * 128 + (((((i & 63) >> 3) ^ (r & 7)) & 7) << 4) + (i & 7) * 2);
                *reinterpret_cast<uint2 *>(&S.act[i >> 6][off]) =
                    make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
                if (NS == 3) {          // the remainders, as operand row r + 64 (same swizzle phase: 64 is a multiple of 8)
                    const float2 f01 = __bfloat1622float2(lo), f23 = __bfloat1622float2(hi);
                    const __nv_bfloat162 r01 = __floats2bfloat162_rn(y0 - f01.x, y1 - f01.y), r23 = __floats2bfloat162_rn(y2 - f23.x, y3 - f23.y);
                    *reinterpret_cast<uint2 *>(&S.act[i >> 6][off + 64 * 128]) =
                        make_uint2(*reinterpret_cast<const uint32_t *>(&r01), *reinterpret_cast<const uint32_t *>(&r23));
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = S.tmem_base;

    // ---------------- phase 2: [128 x 512] x [512 x 32] on the tensor core ----------------
    if (tid == 0) {
        mbar_wait(bar_w, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        constexpr uint32_t idesc = umma_idesc(BM, ND);
#pragma unroll
        for (int part = 0; part < (NS == 3 ? 2 : 1); ++part) {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                const uint32_t a_src = smem_u32(S.act[kb]), b_src = smem_u32(S.w2[part][kb]);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_bf16(tmem_d, umma_desc(a_src + k * UMMA_K * 2), umma_desc(b_src + k * UMMA_K * 2), idesc, (part | kb | k) ? 1u : 0u);
            }
        }
        umma_commit(bar_mma);
    }

    // ---------------- phase 3a: R threads read the accumulator (thread <-> row <-> TMEM lane), LN32, publish ----------------
    TailTcPost &Z = S.post;              // (the W2 tiles are dead once bar_mma has completed)
    if (NS == 3) {                       // accumulator rows 64 .. 64 + R - 1 hold the lo-part products of rows 0 .. R - 1
        if (tid >= 64 && tid < 64 + R) {
            mbar_wait(bar_mma, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float L[ND];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
            tmem_ld16(taddr, L);
            tmem_ld16(taddr + 16, L + 16);
            tmem_ld_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
            for (int o = 0; o < ND; ++o) Z.lo[tid - 64][o] = L[o];
        }
        __syncthreads();
    }
    if (tid < R) {
        const int m = m0 + tid;
        mbar_wait(bar_mma, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float P[ND];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        tmem_ld16(taddr, P);
        tmem_ld16(taddr + 16, P + 16);
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (NS == 3) {
#pragma unroll
            for (int o = 0; o < ND; ++o) P[o] += Z.lo[tid][o];
        }
        float s = 0.0f;
#pragma unroll
        for (int o = 0; o < ND; ++o) { P[o] += w.p2_b[o]; s += P[o]; }
        const float mean = s * (1.0f / ND);
        float ss = 0.0f;
#pragma unroll
        for (int o = 0; o < ND; ++o) { P[o] -= mean; ss = fmaf(P[o], P[o], ss); }
        const float rstd = rsqrtf(ss * (1.0f / ND) + kLnEps);
        float sa = S.colsum[ND];                              // sum_c b_q[c]
#pragma unroll
        for (int o = 0; o < ND; ++o) {
            P[o] = lrelu(fmaf(P[o] * rstd, w.ln2_w[o], w.ln2_b[o]));
            Z.Ps[tid][o] = P[o];
            Z.Ts[tid][o] = NS == 3 ? gate_tanh(P[o]) : tanh_fast(P[o]);   // MUFU.TANH on the bf16-tolerance path only
            sa = fmaf(S.colsum[o], P[o], sa);
        }
        // mean over the 64 channels of A = Wq[:, :32] P + b_q, without forming A
        Z.meanA[tid] = sa * (1.0f / QC);
        if (m < p.rows) {
#pragma unroll
            for (int o = 0; o < ND; o += 4)
                *reinterpret_cast<float4 *>(p.P + (long long)m * ND + o) = make_float4(P[o], P[o + 1], P[o + 2], P[o + 3]);
        }
    }
    __syncthreads();          // (the MMAs have completed: the A-operand tile is free for the reduction below)

    // ---------------- phase 3b: 512 / R threads per row ----------------
    // The 32 value-head units and 64 Q channels of a row are cut into SIXTEEN groups (2 units + 4 channels each) whatever R is:
    // a thread takes one group (R = 32) or two (R = 64), every group's 22 partial sums go to shared memory and the row's first
    // thread adds the sixteen partials in group order.  The association of every sum is therefore the same for both R, i.e. a
    // row's result does not depend on how many rows the launch has -- which is what lets a trajectory or graph shard of a
    // rollout reproduce the single-GPU run bit for bit (R follows the batch size, see launch_tail_tc).
    {
        const int nparts = kTtThreads / R;                    // 16 or 8
        const int rloc = tid % R, part = tid / R;             // part is warp-uniform (R is a multiple of 32)
        const int m = m0 + rloc;
        const bool live = m < p.rows;
        float P[ND], Tp[ND];
#pragma unroll
        for (int o = 0; o < ND; ++o) { P[o] = Z.Ps[rloc][o]; Tp[o] = Z.Ts[rloc][o]; }
        const float meanA = Z.meanA[rloc];
        float *Ao = p.A + (long long)m * QC;
        float *Aco = p.Ac ? p.Ac + (long long)m * QC : nullptr;
        // the qa rows of the CTA are assembled in shared memory (behind the partial sums in the free A-operand tile) and leave as
        // one contiguous block: word-by-word stores with lane <-> row cost 32 L1 wavefronts per instruction
        constexpr int kQsPitch = ECORD_QA_STRIDE + 1;
        uint32_t *qs_base = reinterpret_cast<uint32_t *>(S.act) + 24576;              // 96 KB into the 128 KB tile
        uint32_t *qw = p.qa ? qs_base + rloc * kQsPitch : nullptr;
        float *red = reinterpret_cast<float *>(S.act);                                // [22][16 groups][64 rows]: 88 KB
        float acc[22];                                        // Wg^T a' (16) | Cc^T a' (3) | |a'|^2 | LA | v   (partial sums)
        const int ngrp = 16 / nparts;                         // groups per thread: 1 or 2
        for (int gq = 0; gq < ngrp; ++gq) {
            const int grp = part * ngrp + gq;
#pragma unroll
            for (int i = 0; i < 22; ++i) acc[i] = 0.0f;
            // state value: v = W_v2 LReLU(W_v1 tanh(P) + b_v1) + b_v2 -- this group's value-head units
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp * 2 + jj;
                const float4 *vr = reinterpret_cast<const float4 *>(&S.v1[j * ND]);
                float h0 = w.v1_b[j], h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                for (int o4 = 0; o4 < ND / 4; ++o4) {
                    const float4 v4 = vr[o4];
                    h0 = fmaf(v4.x, Tp[4 * o4], h0); h1 = fmaf(v4.y, Tp[4 * o4 + 1], h1);
                    h2 = fmaf(v4.z, Tp[4 * o4 + 2], h2); h3 = fmaf(v4.w, Tp[4 * o4 + 3], h3);
                }
                acc[21] = fmaf(w.v2_w[j], lrelu((h0 + h1) + (h2 + h3)), acc[21]);
            }
            // this group's Q channels: a_c, then everything that is linear in it
            for (int cc_ = 0; cc_ < 4; ++cc_) {
                const int c = grp * 4 + cc_;
                const float4 *qr = reinterpret_cast<const float4 *>(&S.q1[c * ND]);
                float a0 = w.q1_b[c], a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                for (int o4 = 0; o4 < ND / 4; ++o4) {
                    const float4 q4 = qr[o4];
                    a0 = fmaf(q4.x, P[4 * o4], a0); a1 = fmaf(q4.y, P[4 * o4 + 1], a1);
                    a2 = fmaf(q4.z, P[4 * o4 + 2], a2); a3 = fmaf(q4.w, P[4 * o4 + 3], a3);
                }
                const float a = (a0 + a1) + (a2 + a3);
                const float ac = a - meanA;
                acc[20] = fmaf(w.q2_w[c] * w.qln_w[c], ac, acc[20]);
                if (live) {
                    if (!qw) Ao[c] = a;               // A itself is read by the exact Q kernel only
                    if (Aco) Aco[c] = ac;
                }
                if (qw) {
                    const float4 cc4 = *reinterpret_cast<const float4 *>(&S.cc[c * 4]);
                    const float ap = ac + cc4.w;
                    qw[c] = split_h2(w.qln_w[c] * ap);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 g4 = *reinterpret_cast<const float4 *>(&S.wg[c * 16 + 4 * k4]);
                        acc[4 * k4] = fmaf(g4.x, ap, acc[4 * k4]); acc[4 * k4 + 1] = fmaf(g4.y, ap, acc[4 * k4 + 1]);
                        acc[4 * k4 + 2] = fmaf(g4.z, ap, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(g4.w, ap, acc[4 * k4 + 3]);
                    }
                    acc[16] = fmaf(cc4.x, ap, acc[16]); acc[17] = fmaf(cc4.y, ap, acc[17]); acc[18] = fmaf(cc4.z, ap, acc[18]);
                    acc[19] = fmaf(ap, ap, acc[19]);
                }
            }
            // publish the group's partial sums: red[i][group][row] (conflict-free)
#pragma unroll
            for (int i = 0; i < 22; ++i) red[(i * 16 + grp) * 64 + rloc] = acc[i];
        }
        __syncthreads();
        if (part == 0 && live) {
#pragma unroll
            for (int i = 0; i < 22; ++i) acc[i] = red[(i * 16) * 64 + rloc];
            for (int q = 1; q < 16; ++q) {
#pragma unroll
                for (int i = 0; i < 22; ++i) acc[i] += red[(i * 16 + q) * 64 + rloc];
            }
            const float vv = acc[21] + w.v2_b[0];
            p.v[m] = vv;
            if (p.LA) p.LA[m] = acc[20];
            if (qw) {
                uint32_t e[19];
#pragma unroll
                for (int i = 0; i < 19; ++i) e[i] = split_h2(acc[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {                 // Wg^T a': hi words, lo words
                    qw[64 + j] = (e[2 * j] & 0xffffu) | (e[2 * j + 1] << 16);
                    qw[72 + j] = (e[2 * j] >> 16) | (e[2 * j + 1] & 0xffff0000u);
                }
                // row 64 of W2: fp16 [ca0h ca1h ca2h ca1h 0 0 ca2h 0 | ca0l ca1l ca2l 0 0 0 0 0]
                qw[80] = (e[16] & 0xffffu) | (e[17] << 16);
                qw[81] = (e[18] & 0xffffu) | (e[17] << 16);
                qw[82] = 0u;
                qw[83] = e[18] & 0xffffu;
                qw[84] = (e[16] >> 16) | (e[17] & 0xffff0000u);
                qw[85] = e[18] >> 16;
                qw[86] = 0u;
                qw[87] = 0u;
                float *qo = reinterpret_cast<float *>(qw);
                qo[88] = acc[19];
                qo[89] = acc[20];
                qo[90] = vv;
                qo[91] = 0.0f;
            }
        }
    }
    __syncthreads();
    if (p.qa) {
        const int nrows = min(R, p.rows - m0);
        const uint32_t *qs = reinterpret_cast<const uint32_t *>(S.act) + 24576;
        uint32_t *gq = reinterpret_cast<uint32_t *>(p.qa + (long long)m0 * ECORD_QA_STRIDE);
        for (int idx = tid; idx < nrows * ECORD_QA_STRIDE; idx += kTtThreads) {
            const int r = idx / ECORD_QA_STRIDE, k = idx - r * ECORD_QA_STRIDE;
            gq[idx] = qs[r * (ECORD_QA_STRIDE + 1) + k];
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u) : "memory");
    }
}
}  // namespace

int tail_tc_init_attrs() {
    ECORD_CUDA(cudaFuncSetAttribute(k_tail_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailTcSmem) + 1024));
    ECORD_CUDA(cudaFuncSetAttribute(k_tail_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailTcSmem) + 1024));
    return 0;
}

int launch_tail_tc(const ecord_weights *w, const ecord_policy *p, int split, cudaStream_t st) {
    static_assert(sizeof(TailTcSmem) + 1024 <= 227 * 1024, "TailTcSmem too large");
    PFN_encodeTiled enc;
    int rc = get_encoder(&enc);
    if (rc) return rc;
    CUtensorMap mW2;
    // p1_w_bf16 = [W1 hi | W1 lo | W2 hi | W2 lo]: the map covers W2 hi (rows 0..31) and W2 lo (rows 32..63)
    const __nv_bfloat16 *w2 = reinterpret_cast<const __nv_bfloat16 *>(w->p1_w_bf16) + (size_t)2 * ECORD_DIM_P1 * H;
    if ((rc = make_map(enc, &mW2, w2, 2 * ND, P1, ND))) return rc;
    // 32 rows per CTA while that stays within one wave of CTAs (one per SM: the A-operand tile is 128 KB), else 64
    const int R = ceil_div(p->rows, 32) <= kNumSMs ? 32 : 64;
    if (split)
        ECORD_PDL_LAUNCH(k_tail_tc<3>, dim3(ceil_div(p->rows, R)), dim3(kTtThreads), sizeof(TailTcSmem) + 1024, st, mW2, *w, *p, R);
    else
        ECORD_PDL_LAUNCH(k_tail_tc<1>, dim3(ceil_div(p->rows, R)), dim3(kTtThreads), sizeof(TailTcSmem) + 1024, st, mW2, *w, *p, R);
    return 0;
}
