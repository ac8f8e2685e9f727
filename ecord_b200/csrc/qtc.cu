// qtc.cu -- ECORD_Q_TC: the per-vertex Q head with its contraction on 5th-gen tensor cores (sm_100a).
//
// Same function as qselect.cu (SURVEY.md 8a rows a7 + a8 per-vertex half + a9; reference
// models/rnn_decoder.py:285-311, actors/ecord.py:78-91,123-193), evaluated in the centred / split form of
// ECORD_Q_FAST:
//     x_c   = a'_c[g,t] + b'_c[n] + Cc_c . f[n,t]            a' = Ac + bconst,  b' = Wg node_emb[n]  (channel-centred)
//     sigma = sqrt(sum_c x_c^2 / 64 + eps)
//     q     = 0.505 (lin / sigma + Wb) + 0.495 (sum_c w_c |gamma_c x_c + beta_c sigma|) / sigma + b_o + v[g,t]
//
// What the tensor core does.  For one tile of 128 vertices of a graph and one trajectory t, the 128 x 64 block
// gamma_c x_c is a small GEMM  [128 rows] x [K] x [N = 80]  with fp32 accumulation in TMEM.  Every operand is split
// into an fp16 hi part and an fp16 lo remainder and the products hi.hi + lo.hi + hi.lo are accumulated (4 MMAs of
// K = 16, kind::f16, per trajectory), which makes the result fp32-grade (~1e-6 relative; operands must stay below
// the fp16 range, 65504).  A single low-precision product (tf32 or bf16) puts a fixed per-vertex bias of ~3e-4 on Q,
// enough to change which vertex wins the arg-max and to degrade the rollout.  Logical operands:
//     A operand rows (one per vertex, written by the vertex's own thread)
//         R1 = node_emb[n] (16), hi block and lo block                TENSOR memory (tcgen05.st), written once per CTA
//         R2 = [ f0 f1h f2h f1l 1 1 f2l 0 | f0 f1h f2h 0 0 0 0 0 ]    TENSOR memory (tcgen05.st), rewritten per trajectory
//     B operand rows (one per output column), blocks W1 hi, W1 lo, W2
//         c < 64 : W1 = gamma_c Wg[c][0..15];   W2 = [ C0h C1h C2h C1h ga'_hi ga'_lo C2h 0 | C0l C1l C2l 0 ... ]  (C = gamma_c Cc[c])
//         c = 64 : W1 = (Wg^T a')[0..15];       W2 = [ ca0h ca1h ca2h ca1h 0 0 ca2h 0 | ca0l ca1l ca2l 0 ... ]  (ca = Cc^T a')
//                                                                              ->  d = a' . (b' + Cc f)
//     MMAs: R1h.W1h + R1l.W1h + R1h.W1l + R2.W2
// Column 64 turns the LayerNorm sum of squares into a quadratic form,
//     sum_c x_c^2 = |a'|^2 + 2 d + |b' + Cc f|^2,     |b' + Cc f|^2 = bb[n] + 2 cb[n].f + f^T (Cc^T Cc) f,
// so the CUDA cores are left with the |.| sum: 64 packed FFMA2 per (vertex, trajectory) row instead of ~500 FFMA.
// Only the per-trajectory entries of the B operand (ga', Wg^T a', Cc^T a': 88 words from the policy tail) and the
// 128 R2 rows change between MMAs.
//
// CTA = 4 row warps (thread i <-> vertex i of the tile <-> TMEM lane i) + 1 control warp (copies the per-trajectory
// operand entries, issues tcgen05.mma / commit).  No CTA-wide barrier in the loop: the roles meet on mbarriers
// (`ready`: R2 written + accumulator drained, 4 warp arrivals; `full`: tcgen05.commit).  A CTA owns 128 TMEM columns
// (accumulator [0,80), R2 double-buffered [80,96), R1 hi / lo [96,112)) so that FOUR CTAs share an SM: a row warp hands trajectory t+1 to the tensor core right
// after reading the accumulator of t out of TMEM, its MMAs overlap the epilogue math of t, and the other CTAs'
// warps cover what latency remains (the kernel is latency-, not issue-bound: see profiles/).
// Shared-memory operand blocks are [rows][16 fp16] with 32-byte rows in the canonical K-major SWIZZLE_32B layout
// (16-byte chunk index XOR bit 2 of the row; 8-row groups 256 B apart); only the B operand lives there.  R2 lives in tensor
// memory because a generic-proxy write to a shared-memory operand needs fence.proxy.async (a MEMBAR) before the MMA may
// read it, and that fence would drain the row warps' prefetched global loads and arg-max atomics every trajectory; R1
// lives there because an A operand in TMEM costs no shared-memory read per MMA (12 of the 22 KB a trajectory used to pull
// through the L1 data pipe).
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>

namespace {
constexpr int Q = ECORD_DIM_Q;        // 64
constexpr int QT = ECORD_QTILE;       // 128 vertices per tile
constexpr int QN = 80;                // MMA N (multiple of 16 for M = 128)
constexpr int QA = ECORD_QA_STRIDE;   // 88
constexpr int kTcMaxT = 32;           // trajectories per CTA (rows of qa staged in shared memory)
constexpr int kCtasPerSM = 4;         // 4 x 128 TMEM columns; shared memory and registers are sized for it
constexpr int A_BLK = QT * 32;        // bytes of one [128][16] fp16 operand block
constexpr int B_BLK = QN * 32;        // bytes of one [80][16] block
constexpr int kQtcThreads = 160;
constexpr int kTmemA = QN;            // TMEM columns [0,80): accumulator; [80,88) and [88,96): the per-trajectory A operand R2
                                      // (16 fp16 = 8 columns), double-buffered by trajectory parity
constexpr int kTmemR1 = QN + 16;      // TMEM columns [96,104): R1 hi, [104,112): R1 lo (node_emb of the tile's vertices, written once per CTA)
constexpr int kTmemCols = 128;

constexpr int kRing = 8;               // cp.async prefetch depth of the per-(vertex, trajectory) state, in trajectories
constexpr int kImgBytes = 3 * B_BLK;   // packed image of the B operand: W1 hi | W1 lo | W2

struct QTcSmem {
    uint8_t W[3][B_BLK];               // B operand: W1 hi | W1 lo | W2
    float qa[kTcMaxT][QA];             // per-(g,t) rows of ecord_policy.qa
    int sp[kRing + 1][QT];             // state ring (cp.async), one slot per trajectory in flight
    float pk[kRing + 1][QT];
    uint64_t full;                     // MMA -> row warps: accumulator complete (one phase per trajectory)
    uint64_t ready;                    // row warps -> MMA: R2 written, accumulator drained (4 warp arrivals per trajectory)
    uint64_t setup;                    // bulk copies of the set-up landed
    uint32_t tmem_base;
};

// epilogue constants, passed by value (constant bank operands of the FFMAs)
struct QTcConst { float bet[Q], wo[Q], CC[6], LC[3], Wb; };

__device__ __forceinline__ uint32_t sw32(int r, int k) {      // byte offset of fp16 element (row r, k) in a SWIZZLE_32B block
    return (uint32_t)(r * 32 + ((((k >> 3) ^ (r >> 2)) & 1) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {   // D = f32, A = B = fp16, K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// A operand from tensor memory (lane = row, two fp16 elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// fp16 hi / lo split of an fp32 value: x ~ hi + lo with |x - hi - lo| <= 2^-22 |x| (or 3e-8 absolute below 2^-3)
__device__ __forceinline__ void split_h(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                         uint32_t a5, uint32_t a6, uint32_t a7) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// SAMPLE: tau > 0.  The reference's exponential race argmin_n -log(U_n) / exp((q_n - max q) / tau) (actors/ecord.py:12-19)
// is the Gumbel-max trick in disguise: argmax_n q_n / tau - log(-log U_n), which needs no maximum first and therefore fits
// this single pass.  U_n comes from the same Philox counter (vertex, unit, step; seed) as in the exact kernel, so both
// kernels draw the same vertex wherever their Q-values agree to rounding.
template <bool DUMP_Q, bool SAMPLE>
__global__ void __launch_bounds__(kQtcThreads, kCtasPerSM)
k_q_tc(ecord_graph g, ecord_env e, ecord_policy p, const float *__restrict__ q_tc, const __grid_constant__ QTcConst K,
       const float *__restrict__ q2_b, const int32_t *step_base, int step_off, int t_chunk, float inv_tau, uint64_t seed) {
    extern __shared__ uint8_t smem_raw[];
    QTcSmem &S = *reinterpret_cast<QTcSmem *>(smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_row = warp < 4;
    const int tile = blockIdx.x;
    const int gi = g.qtile_graph[tile], v0 = g.qtile_v0[tile];
    const int T = e.num_tradj;
    const int t_lo = blockIdx.y * t_chunk, nt = min(T, t_lo + t_chunk) - t_lo;
    const int lo = g.graph_ptr[gi], n = g.graph_ptr[gi + 1] - lo;
    pdl_launch_dependents();
    const uint32_t full0 = smem_u32(&S.full), ready0 = smem_u32(&S.ready), setup_bar = smem_u32(&S.setup);
    constexpr uint32_t idesc = umma_idesc_f16(QT, QN);

    // ---------------- set-up ----------------
    if (warp == 4) {
        if (lane == 0) {
            mbar_init(full0, 1);
            mbar_init(ready0, 4);
            mbar_init(setup_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // B-operand image (constant) and this CTA's qa rows (written by the policy tail: after the PDL wait):
            // bulk copies through the async proxy
            const uint32_t qa_bytes = (uint32_t)nt * QA * 4u;
            mbar_expect_tx(setup_bar, (uint32_t)kImgBytes + qa_bytes);
            bulk_g2s(smem_u32(S.W), q_tc + 2816, kImgBytes, setup_bar);
            pdl_wait();
            bulk_g2s(smem_u32(S.qa), p.qa + (long long)(gi * T + t_lo) * QA, qa_bytes, setup_bar);
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // per-vertex: the R1 row (node_emb as fp16 hi + lo) and the quadratic-form scalars
    const bool valid = is_row && v0 + tid < n;
    const int v = is_row ? min(v0 + tid, n - 1) : 0;
    float bb = 0.f, cb0 = 0.f, cb1 = 0.f, cb2 = 0.f, lb = 0.f;
    uint32_t r1h[8], r1l[8];
    // state row entries: element (t_lo, v); consecutive trajectories are n elements apart
    const long long base0 = (long long)lo * T + (long long)t_lo * n + v;
    const int *__restrict__ sp_ptr = e.sp + base0;          // next trajectory to fetch
    const float *__restrict__ pk_ptr = e.peek + base0;
    const uint32_t ring_sp = smem_u32(&S.sp[0][tid & (QT - 1)]), ring_pk = smem_u32(&S.pk[0][tid & (QT - 1)]);
    if (is_row) {
        const float4 *ep = reinterpret_cast<const float4 *>(p.node_emb + (long long)(lo + v) * 16);
        const float4 *qv = reinterpret_cast<const float4 *>(p.node_qv + (long long)(lo + v) * 8);
        float4 x[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) x[c4] = ep[c4];
        const float4 q0 = qv[0];
        lb = qv[1].x;
        // the vertex's R1 row: 16 fp16 hi values and the 16 lo remainders, as eight 32-bit words each; they go to tensor memory
        // (below, once the allocation is known): an A operand in TMEM costs no shared-memory read per MMA, and the L1 data
        // pipe -- tensor-core operand reads + LSU wavefronts -- is this kernel's busiest unit (78 % before this change)
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
            const float e[8] = {x[2 * c8].x, x[2 * c8].y, x[2 * c8].z, x[2 * c8].w,
                                x[2 * c8 + 1].x, x[2 * c8 + 1].y, x[2 * c8 + 1].z, x[2 * c8 + 1].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __half h0, l0, h1, l1;
                split_h(e[2 * i], h0, l0);
                split_h(e[2 * i + 1], h1, l1);
                r1h[4 * c8 + i] = pack_h2(h0, h1);
                r1l[4 * c8 + i] = pack_h2(l0, l1);
            }
        }
        bb = q0.x; cb0 = 2.0f * q0.y; cb1 = 2.0f * q0.z; cb2 = 2.0f * q0.w;
        pdl_wait();                     // the state planes and step counter belong to the step: predecessor kernels first
#pragma unroll
        for (int j = 0; j < kRing; ++j) {
            if (j < nt) {
                cp_async4(ring_sp + (uint32_t)(j * QT * 4), sp_ptr);
                cp_async4(ring_pk + (uint32_t)(j * QT * 4), pk_ptr);
                sp_ptr += n; pk_ptr += n;
            }
            cp_async_commit();
        }
    }
    const int steps_done = (step_base ? *step_base : 0) + step_off;      // (row threads: after their pdl_wait)
    const float bo = q2_b[0];
    const float inv_n = 1.0f / (float)n;
    constexpr float inv_clip = 1.0f / (float)ECORD_STATIC_CLIP;
    float *q_ptr = DUMP_Q ? p.q + base0 : nullptr;
    unsigned long long *key_ptr = p.keys + (gi * T + t_lo);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                 // barrier inits, TMEM base, R1
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    if (is_row) {        // R1 hi / lo rows -> tensor memory; published to the control warp by the first `ready` arrival
        const uint32_t trow0 = tmem_base + ((uint32_t)(warp * 32) << 16);
        tmem_st8(trow0 + (uint32_t)kTmemR1, r1h[0], r1h[1], r1h[2], r1h[3], r1h[4], r1h[5], r1h[6], r1h[7]);
        tmem_st8(trow0 + (uint32_t)(kTmemR1 + 8), r1l[0], r1l[1], r1l[2], r1l[3], r1l[4], r1l[5], r1l[6], r1l[7]);
    }

    if (warp == 4) {
        // =============== control warp: per-trajectory B-operand entries, MMA issue ===============
        mbar_wait(setup_bar, 0);
        // Per-trajectory B-operand entries: the qa row holds them as ready-made 32-bit words (fp16 pairs, see
        // ecord_policy.qa); this lane copies two channel words (ga' hi|lo -> W2 halves 4|5 of rows c, c+32) and, for
        // lane < 24, one word of row 64 (column d): W1 hi (lanes 0..7), W1 lo (8..15), W2 (16..23).
        uint8_t *Wb = S.W[0];
        const uint32_t dst_c0 = 2u * B_BLK + sw32(lane, 4), dst_c1 = 2u * B_BLK + sw32(lane + 32, 4);
        const uint32_t dst_r = (uint32_t)(lane >> 3) * B_BLK + sw32(Q, 2 * (lane & 7));
        const uint64_t dB0 = umma_desc_sw32(smem_u32(S.W[0])), dB1 = umma_desc_sw32(smem_u32(S.W[1])),
                       dB2 = umma_desc_sw32(smem_u32(S.W[2]));
        for (int tt = 0; tt < nt; ++tt) {
            // the B operand was last read by the MMAs of trajectory tt-1
            if (tt >= 1) mbar_wait(full0, (uint32_t)(tt - 1) & 1u);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(S.qa[tt]);
            *reinterpret_cast<uint32_t *>(Wb + dst_c0) = src[lane];
            *reinterpret_cast<uint32_t *>(Wb + dst_c1) = src[lane + 32];
            if (lane < 24) *reinterpret_cast<uint32_t *>(Wb + dst_r) = src[64 + lane];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_wait(ready0, (uint32_t)tt & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // x = (e_hi + e_lo)(W_hi + W_lo) ~ e_hi W_hi + e_lo W_hi + e_hi W_lo  (the lo.lo term is 2^-22 relative)
                umma_f16_ts(tmem_base, tmem_base + (uint32_t)kTmemR1, dB0, idesc, 0u);
                umma_f16_ts(tmem_base, tmem_base + (uint32_t)(kTmemR1 + 8), dB0, idesc, 1u);
                umma_f16_ts(tmem_base, tmem_base + (uint32_t)kTmemR1, dB1, idesc, 1u);
                umma_f16_ts(tmem_base, tmem_base + (uint32_t)(kTmemA + 8 * (tt & 1)), dB2, idesc, 1u);
                umma_commit(full0);
            }
            __syncwarp();
        }
        // the row warps wait for every accumulator, so all MMAs have completed when they reach the final barrier
    } else {
        // =============== row warps: thread <-> vertex <-> TMEM lane ===============
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        // prepare(tt): features of trajectory tt -> R2 row in tensor memory (buffer tt&1); refill the state ring
        auto prepare = [&](int tt, float &o0, float &o1, float &o2) {
            cp_async_wait<kRing - 1>();
            const int slot = tt % (kRing + 1);
            const int sp_c = S.sp[slot][tid];
            const float pk_c = S.pk[slot][tid];
            if (tt + kRing < nt) {                 // trajectory tt + kRing goes into the slot trajectory tt - 1 vacated
                const uint32_t off = (uint32_t)(((tt + kRing) % (kRing + 1)) * QT * 4);
                cp_async4(ring_sp + off, sp_ptr);
                cp_async4(ring_pk + off, pk_ptr);
                sp_ptr += n; pk_ptr += n;
            }
            cp_async_commit();
            o0 = (float)(sp_c & 1);
            o1 = pk_c * inv_n;
            o2 = fminf((float)(steps_done - (sp_c >> 1)), (float)ECORD_STATIC_CLIP) * inv_clip;
            __half f1h, f1l, f2h, f2l;
            split_h(o1, f1h, f1l);
            split_h(o2, f2h, f2l);
            const __half hz = __ushort_as_half((unsigned short)0), h1 = __ushort_as_half((unsigned short)0x3c00);   // 0, 1
            const __half f0h = (sp_c & 1) ? h1 : hz;
            const uint32_t w0 = pack_h2(f0h, f1h);
            tmem_st8(trow + (uint32_t)(kTmemA + 8 * (tt & 1)), w0, pack_h2(f2h, f1l), pack_h2(h1, h1), pack_h2(f2l, hz),
                     w0, pack_h2(f2h, hz), 0u, 0u);
            tmem_st_wait();
        };
        // hand a trajectory to the tensor core: its R2 row is in tensor memory and the accumulator has been drained
        auto hand_over = [&]() {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(ready0);
        };
        float f0, f1, f2, g0 = 0.f, g1 = 0.f, g2 = 0.f;
        prepare(0, f0, f1, f2);
        hand_over();
        mbar_wait(setup_bar, 0);          // qa rows (bulk copy) are read by the epilogue
        for (int tt = 0; tt < nt; ++tt) {
            // The R2 operand is double-buffered, so the next trajectory's row is written while the MMAs of this one run:
            // the hand-off cycle of a CTA (accumulator read -> arrive -> control warp -> MMAs -> commit -> accumulator read)
            // then contains only the accumulator read, not the feature preparation and its tcgen05.st round trip.
            if (tt + 1 < nt) prepare(tt + 1, g0, g1, g2);
            mbar_wait(full0, (uint32_t)tt & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float gx[Q];
            tmem_ld16(trow, gx);
            tmem_ld16(trow + 16, gx + 16);
            tmem_ld16(trow + 32, gx + 32);
            tmem_ld16(trow + 48, gx + 48);
            const float d = tmem_ld1(trow + 64);
            tmem_ld_wait();
            // accumulator drained, MMAs of trajectory tt complete: trajectory tt+1 goes to the tensor core before the
            // epilogue math, which its MMAs overlap
            if (tt + 1 < nt) hand_over();
            const float *sc = S.qa[tt];
            // |b' + Cc f|^2 = bb + 2 cb.f + f^T CC f
            float xx = fmaf(f0, fmaf(K.CC[0], f0, fmaf(2.0f * K.CC[1], f1, fmaf(2.0f * K.CC[2], f2, cb0))), bb);
            xx = fmaf(f1, fmaf(K.CC[3], f1, fmaf(2.0f * K.CC[4], f2, cb1)), xx);
            xx = fmaf(f2, fmaf(K.CC[5], f2, cb2), xx);
            const float ss = fmaxf(fmaf(2.0f, d, sc[88]) + xx, 0.0f);
            const float var = fmaf(ss, 1.0f / Q, kLnEps);
            const float rstd = rsqrtf(var);
            const float sigma = var * rstd;
            // sum_c w_c |gamma_c x_c + beta_c sigma| on packed fp32 pairs (FFMA2 with |.| and scalar-broadcast operands):
            // 64 instructions for the 64 channels
            const unsigned long long sig2 = pack2(sigma, sigma);
            unsigned long long acc01 = 0ull, acc23 = 0ull;
#pragma unroll
            for (int c = 0; c < Q; c += 4) {
                const unsigned long long y01 = fma2(pack2(K.bet[c + 0], K.bet[c + 1]), sig2, pack2(gx[c + 0], gx[c + 1]));
                const unsigned long long y23 = fma2(pack2(K.bet[c + 2], K.bet[c + 3]), sig2, pack2(gx[c + 2], gx[c + 3]));
                acc01 = fma2(pack2(K.wo[c + 0], K.wo[c + 1]), abs2(y01), acc01);
                acc23 = fma2(pack2(K.wo[c + 2], K.wo[c + 3]), abs2(y23), acc23);
            }
            float a0, a1, a2, a3;
            unpack2(acc01, a0, a1);
            unpack2(acc23, a2, a3);
            const float acc = (a0 + a1) + (a2 + a3);
            const float lin = sc[89] + lb + fmaf(K.LC[2], f2, fmaf(K.LC[1], f1, K.LC[0] * f0));
            const float dot = fmaf(0.505f, fmaf(lin, rstd, K.Wb), 0.495f * acc * rstd);
            const float q = (dot + bo) + sc[90];
            if (DUMP_Q) { if (valid) *q_ptr = q; q_ptr += n; }
            float sel = q;
            if (SAMPLE) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)(v0 + tid), (uint32_t)(gi * T + t_lo + tt), (uint32_t)steps_done,
                                                         0x45434F52u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                sel = fmaf(q, inv_tau, -__logf(-__logf(uniform01_open(r.x))));
            }
            // arg-max: ordered-int max across the warp, lowest vertex among equals, one atomic per warp
            const uint32_t u = __float_as_uint(sel);
            const uint32_t ord = valid ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0u;
            const uint32_t mx = __reduce_max_sync(0xffffffffu, ord);
            const uint32_t bal = __ballot_sync(0xffffffffu, ord == mx);
            if (lane == 0 && mx != 0u) {
                const int vbest = v0 + warp * 32 + (__ffs(bal) - 1);
                atomicMax(key_ptr + tt, ((unsigned long long)mx << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)vbest));
            }
            f0 = g0; f1 = g1; f2 = g2;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
}
}  // namespace

constexpr int kQtcSmemBytes = 56 * 1024;

int qtc_init_attrs() {
    // the dynamic request is padded so that at most four CTAs (4 x 128 TMEM columns) share an SM
    ECORD_CUDA(cudaFuncSetAttribute(k_q_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQtcSmemBytes));
    ECORD_CUDA(cudaFuncSetAttribute(k_q_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQtcSmemBytes));
    ECORD_CUDA(cudaFuncSetAttribute(k_q_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQtcSmemBytes));
    ECORD_CUDA(cudaFuncSetAttribute(k_q_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQtcSmemBytes));
    return 0;
}

// trajectory split that minimises (waves of CTAs) x (set-up + trajectories per CTA)
static int qtc_t_chunk(int tiles, int T) {
    static const int forced = [] { const char *e = getenv("ECORD_QTC_CHUNK"); return e ? atoi(e) : 0; }();      // A/B tests
    if (forced > 0) return forced < T ? (forced < kTcMaxT ? forced : kTcMaxT) : (T < kTcMaxT ? T : kTcMaxT);
    const int slots = kCtasPerSM * kNumSMs, setup = 6;
    long long best_cost = -1;
    int best = 1;
    for (int s = 1; s <= T; ++s) {
        const int tc = ceil_div(T, s);
        if (tc > kTcMaxT) continue;
        const long long waves = ceil_div((long long)tiles * ceil_div(T, tc), slots);
        const long long cost = waves * (setup + tc);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tc; }
    }
    return best;
}

int launch_q_tc(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                const int32_t *step_base, int step_off, float tau, uint64_t seed, cudaStream_t st) {
    static_assert(sizeof(QTcSmem) + 256 <= kQtcSmemBytes, "QTcSmem too large");
    QTcConst K;
    memcpy(K.bet, w->q_fold + 4 * Q, Q * sizeof(float));
    memcpy(K.wo, w->q_fold + 5 * Q, Q * sizeof(float));
    memcpy(K.CC, w->q_fold + 580, 6 * sizeof(float));
    memcpy(K.LC, w->q_fold + 576, 4 * sizeof(float));     // LC[3], Wb
    const int T = env->num_tradj;
    const int t_chunk = qtc_t_chunk(g->num_qtiles, T);
    const dim3 grid(g->num_qtiles, ceil_div(T, t_chunk));
    const float inv_tau = tau > 0.0f ? 1.0f / tau : 0.0f;
#define ECORD_QTC_LAUNCH(DQ, SM)                                                                                             \
    ECORD_PDL_LAUNCH((k_q_tc<DQ, SM>), grid, dim3(kQtcThreads), (size_t)kQtcSmemBytes, st, *g, *env, *p, (const float *)w->q_tc, K, \
                     w->q2_b, step_base, step_off, t_chunk, inv_tau, seed)
    if (tau > 0.0f) { if (p->q) ECORD_QTC_LAUNCH(true, true); else ECORD_QTC_LAUNCH(false, true); }
    else { if (p->q) ECORD_QTC_LAUNCH(true, false); else ECORD_QTC_LAUNCH(false, false); }
#undef ECORD_QTC_LAUNCH
    return 0;
}
