// policy.cu -- per-(graph,trajectory) recurrent unit and state head (SURVEY.md 8a rows a6 and the
// per-(g,t) half of a8; "Subsystem 2a").
//
// Replaces, per step (reference paths relative to the upstream tree):
//   _ActorBase._update_internal_state            solvers/actors/_base.py:247-269
//   GRUDecoder.update_embeddings                 models/rnn_decoder.py:369-381
//       u  = LReLU(W_m [e_{a_{t-1}} (32); glob_t (2)] + b_m)            (34 -> 64)
//       h' = GRUCell(u, h)                                              (64 -> 1024; gate order r,z,n)
//   proj_rnn_to_gnn + val_out of _forward_action_value                   models/rnn_decoder.py:84-90,108-113,285-311
//       P  = LReLU(LN32(W2 LReLU(LN512(W1 tanh(h') + b1)) + b2))
//       v  = W_v2 LReLU(W_v1 tanh(P) + b_v1) + b_v2
// and pre-computes the per-(g,t) part of mlp_out[0]:  A = Wq[:, :32] P + b_q   (see qselect.cu).
//
// Rows are m = g*T + t (M = G*T of them), padded to a multiple of 128.
// Two arithmetic paths for the two dense contractions ([M,1088]x[1088,3072] and [M,1024]x[1024,512]):
//   ECORD_GEMM_FP32: CUDA-core FMA SGEMM below (parity mode),
//   ECORD_GEMM_BF16: tcgen05/TMA kernels in gemm_tc.cu (throughput mode).
#include "common.cuh"

// gemm_tc.cu
int launch_gru_tc_ex(const ecord_weights *w, const ecord_policy *p, const int32_t *step_base, int step_off, cudaStream_t st);
int launch_proj1_tc(const ecord_weights *w, const ecord_policy *p, cudaStream_t st);
int tc_init_attrs();

namespace {
constexpr int H = ECORD_DIM_H;      // 1024
constexpr int MSG = ECORD_DIM_MSG;  // 64
constexpr int P1 = ECORD_DIM_P1;    // 512
constexpr int ND = ECORD_DIM_NODE;  // 32

// ------------------------------------------------------------------------------------------------
// u = LReLU(msg_in([last_sel ; glob]))   one warp per row
// glob = ((best - score) / n_g, 0)  (SCORE_FROM_BEST_NORM; 2nd feature dead: quirk C)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_policy_pre(ecord_graph g, ecord_env e, ecord_weights w, ecord_policy p) {
    __shared__ float wt[34 * MSG];   // transposed msg_in weight: wt[i][o]
    __shared__ float wb[MSG];
    msg_in_load(w.msg_w, w.msg_b, wt, wb);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= p.rows) return;
    const int T = e.num_tradj, gi = m / T;
    const float x = p.last_sel[(long long)m * ND + lane];
    const float glob0 = (e.best_score[m] - e.score[m]) / (float)(g.graph_ptr[gi + 1] - g.graph_ptr[gi]);
    float u0, u1;
    msg_in_row(wt, wb, x, glob0, lane, u0, u1);
    msg_in_store(p, m, lane, u0, u1);
}

// ------------------------------------------------------------------------------------------------
// fp32 SGEMM, C[M][ldc] (+col offset) = A[M][K] . B[N][K]^T + bias[N]    (both operands K-contiguous)
// 64x64x16 tiles, 256 threads, 4x4 register blocking.  M % 64 == 0, N % 64 == 0, K % 16 == 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sgemm_nt_bias(const float *__restrict__ A0, const float *__restrict__ A1, const int32_t *step_base, int step_off,
                const float *__restrict__ B, const float *__restrict__ bias, float *__restrict__ C, int K, int ldc) {
    // A = A0 when the step parity is even, A1 when odd (double-buffered hidden state)
    const float *__restrict__ A = (((step_base ? *step_base : 0) + step_off) & 1) ? A1 : A0;
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int lr = tid >> 2, lk = (tid & 3) * 4;        // loader: row 0..63, k-quad
    const int ty = tid >> 4, tx = tid & 15;             // compute: 16x16 threads, 4x4 each
    float acc[4][4] = {};
    const float *Ap = A + (long long)(m0 + lr) * K + lk;
    const float *Bp = B + (long long)(n0 + lr) * K + lk;
    for (int k0 = 0; k0 < K; k0 += 16) {
        const float4 a = *reinterpret_cast<const float4 *>(Ap + k0);
        const float4 b = *reinterpret_cast<const float4 *>(Bp + k0);
        As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float4 bb = *reinterpret_cast<const float4 *>(bias + n0 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
        *reinterpret_cast<float4 *>(C + (long long)(m0 + ty * 4 + i) * ldc + n0 + tx * 4) = o;
    }
}

// ------------------------------------------------------------------------------------------------
// GRU gate math (fp32 mode), element-wise over [M][1024].  gates = [gi (3072) | gh (3072)] incl. biases.
// Follows ATen's cell: r = s(gh_r+gi_r), z = s(gh_z+gi_z), n = tanh(gi_n + gh_n*r), h' = (h-n)*z + n.
// ------------------------------------------------------------------------------------------------
__global__ void k_gru_gates(ecord_policy p, const int32_t *step_base, int step_off) {
    const int cur = ((step_base ? *step_base : 0) + step_off) & 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.rows * H) return;
    const long long m = i / H;
    const int j = (int)(i - m * H);
    const float *gi = p.gates + m * 6144, *gh = gi + 3072;
    const float *hold = p.hidden + (long long)cur * p.rows_pad * H;
    float *hnew = p.hidden + (long long)(1 - cur) * p.rows_pad * H;
    const float r = sigmoidf_acc(gh[j] + gi[j]);
    const float z = sigmoidf_acc(gh[H + j] + gi[H + j]);
    const float n = tanhf(__fadd_rn(gi[2 * H + j], __fmul_rn(gh[2 * H + j], r)));
    const float h = hold[i];
    const float hn = __fadd_rn(__fmul_rn(__fsub_rn(h, n), z), n);
    hnew[i] = hn;
    p.th[i] = tanhf(hn);
}

// ------------------------------------------------------------------------------------------------
// tail: y1 [M][512] -> LN512 -> LReLU -> W2 (512->32) -> LN32 -> LReLU = P ; v = val_out(P) ; A = Wq[:, :32] P + b_q.
//
// One warp owns FOUR rows at once so that every W2 element fetched from shared memory feeds four FMAs
// (LDS : FFMA = 1 : 4, the ratio the SM can sustain), and the 4 x 32 dot products are finished with a
// reduce-scatter butterfly over the warp (62 shuffles for 64 sums) instead of one 5-step all-reduce per output.
// 32 rows per block (8 warps); W2 (64 KB) is staged in shared memory once per block.
// ------------------------------------------------------------------------------------------------
constexpr int kTailR = 4;
constexpr int kTailWarps = 8;
constexpr int kTailRows = kTailR * kTailWarps;
struct TailSmem {
    float w2[ND * P1];          // [o][i]
    float v1t[ND * ND];         // transposed val_out.1: [i][o]
    float qt[ND * ECORD_DIM_Q]; // transposed Wq[:, :32]: [i][c]
    float wg[ECORD_DIM_Q * 16]; // tensor-core Q kernel: centred Wq[:,32:48]  [c][k]
    float cc[ECORD_DIM_Q * 4];  //                        centred C [c][0..2], bconst[c]
};

__device__ __forceinline__ float tf32_round(float x) {     // round-to-nearest onto the tf32 grid (see tc_common.cuh)
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// one butterfly level: lanes with (lane & OFF) == 0 keep v[0..N/2), the others keep v[N/2..N)
template <int N, int OFF>
__device__ __forceinline__ void reduce_scatter_step(float (&v)[64], bool upper) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = upper ? v[i] : v[i + N / 2];
        const float keep = upper ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}
// 64 per-lane partial sums -> lane L ends with the warp totals of entries 2L and 2L+1 in v[0], v[1]
__device__ __forceinline__ void reduce_scatter64(float (&v)[64], int lane) {
    reduce_scatter_step<64, 16>(v, lane & 16);
    reduce_scatter_step<32, 8>(v, lane & 8);
    reduce_scatter_step<16, 4>(v, lane & 4);
    reduce_scatter_step<8, 2>(v, lane & 2);
    reduce_scatter_step<4, 1>(v, lane & 1);
}
__device__ __forceinline__ float group8_sum(float v) {   // all-reduce over the 8 lanes that share a row
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__global__ void __launch_bounds__(kTailWarps * 32, 2)
k_policy_tail(ecord_weights w, ecord_policy p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TailSmem &S = *reinterpret_cast<TailSmem *>(smem_raw);
    for (int i = threadIdx.x; i < ND * P1 / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(S.w2)[i] = reinterpret_cast<const float4 *>(w.p2_w)[i];
    for (int i = threadIdx.x; i < ND * ND; i += blockDim.x) S.v1t[i] = w.v1_w[(i % ND) * ND + (i / ND)];
    for (int i = threadIdx.x; i < ND * ECORD_DIM_Q; i += blockDim.x)
        S.qt[i] = w.q1_w[(i % ECORD_DIM_Q) * ECORD_DIM_Q + (i / ECORD_DIM_Q)];
    if (p.qa) {
        for (int i = threadIdx.x; i < ECORD_DIM_Q * 16; i += blockDim.x) S.wg[i] = w.q_tc[1280 + i];
        for (int i = threadIdx.x; i < ECORD_DIM_Q * 4; i += blockDim.x) S.cc[i] = w.q_tc[2304 + i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = blockIdx.x * kTailRows + warp * kTailR;
    if (m0 >= p.rows) return;

    // ---- LN512 + LReLU on 4 rows (lane holds elements lane + 32 j of each) ----
    float x[kTailR][16];
#pragma unroll
    for (int r = 0; r < kTailR; ++r) {
        const int m = min(m0 + r, p.rows - 1);          // rows past the end are computed on a copy and not stored
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[r][j] = p.y1[(long long)m * P1 + lane + 32 * j]; s += x[r][j]; }
        const float mean = warp_sum(s) * (1.0f / P1);
        float ss = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[r][j] -= mean; ss = fmaf(x[r][j], x[r][j], ss); }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / P1) + kLnEps);
#pragma unroll
        for (int j = 0; j < 16; ++j) x[r][j] *= rstd;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float gj = w.ln1_w[lane + 32 * j], bj = w.ln1_b[lane + 32 * j];
#pragma unroll
        for (int r = 0; r < kTailR; ++r) x[r][j] = lrelu(fmaf(x[r][j], gj, bj));
    }
    // ---- W2: 512 -> 32 for 4 rows, two passes of 16 outputs.  After each pass lane L holds, for row L>>3,
    //      outputs oh*16 + 2*(L&7) + {0,1} ----
    const int grp = lane >> 3, k8 = lane & 7;           // row within the warp, position within the row's 8 lanes
    float pre[4];
#pragma unroll
    for (int oh = 0; oh < 2; ++oh) {
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int oo = 0; oo < 16; ++oo) {
                const float wv = S.w2[(oh * 16 + oo) * P1 + lane + 32 * j];
#pragma unroll
                for (int r = 0; r < kTailR; ++r) acc[r * 16 + oo] = fmaf(wv, x[r][j], acc[r * 16 + oo]);
            }
        }
        reduce_scatter64(acc, lane);
        const int o = oh * 16 + 2 * k8;
        pre[oh * 2 + 0] = acc[0] + w.p2_b[o];
        pre[oh * 2 + 1] = acc[1] + w.p2_b[o + 1];
    }
    // ---- LN32 + LReLU per row (the row's 32 values live in its 8 lanes x 4 slots) ----
    const float mean2 = group8_sum((pre[0] + pre[1]) + (pre[2] + pre[3])) * (1.0f / ND);
    float d[4], ssq = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { d[q] = pre[q] - mean2; ssq = fmaf(d[q], d[q], ssq); }
    const float rstd2 = rsqrtf(group8_sum(ssq) * (1.0f / ND) + kLnEps);
    float Pv[4], Tv[4];
    const int m = m0 + grp;
    const bool live = m < p.rows;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int o = (q >> 1) * 16 + 2 * k8 + (q & 1);
        Pv[q] = lrelu(fmaf(d[q] * rstd2, w.ln2_w[o], w.ln2_b[o]));
        Tv[q] = tanhf(Pv[q]);
        if (live) p.P[(long long)m * ND + o] = Pv[q];
    }
    // ---- every lane of the row group gathers the row's 32 P / tanh(P) values ----
    float Pa[ND], Ta[ND];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o = (q >> 1) * 16 + 2 * kk + (q & 1);
            Pa[o] = __shfl_sync(0xffffffffu, Pv[q], grp * 8 + kk);
            Ta[o] = __shfl_sync(0xffffffffu, Tv[q], grp * 8 + kk);
        }
    }
    // ---- state value: v = W_v2 LReLU(W_v1 tanh(P) + b_v1) + b_v2 ; lane computes 4 of the 32 hidden units ----
    float vpart = 0.0f;
    {
        float4 hv = *reinterpret_cast<const float4 *>(w.v1_b + k8 * 4);
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            const float4 wv = *reinterpret_cast<const float4 *>(&S.v1t[i * ND + k8 * 4]);
            hv.x = fmaf(wv.x, Ta[i], hv.x); hv.y = fmaf(wv.y, Ta[i], hv.y);
            hv.z = fmaf(wv.z, Ta[i], hv.z); hv.w = fmaf(wv.w, Ta[i], hv.w);
        }
        const float4 w2v = *reinterpret_cast<const float4 *>(w.v2_w + k8 * 4);
        vpart = w2v.x * lrelu(hv.x) + w2v.y * lrelu(hv.y) + w2v.z * lrelu(hv.z) + w2v.w * lrelu(hv.w);
    }
    const float vv = group8_sum(vpart) + w.v2_b[0];
    // ---- A = Wq[:, :32] P + b_q ; lane computes 8 of the 64 channels ----
    float a[8];
    {
        const float4 b0 = *reinterpret_cast<const float4 *>(w.q1_b + k8 * 8);
        const float4 b1 = *reinterpret_cast<const float4 *>(w.q1_b + k8 * 8 + 4);
        a[0] = b0.x; a[1] = b0.y; a[2] = b0.z; a[3] = b0.w; a[4] = b1.x; a[5] = b1.y; a[6] = b1.z; a[7] = b1.w;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            const float4 w0 = *reinterpret_cast<const float4 *>(&S.qt[i * ECORD_DIM_Q + k8 * 8]);
            const float4 w1 = *reinterpret_cast<const float4 *>(&S.qt[i * ECORD_DIM_Q + k8 * 8 + 4]);
            a[0] = fmaf(w0.x, Pa[i], a[0]); a[1] = fmaf(w0.y, Pa[i], a[1]); a[2] = fmaf(w0.z, Pa[i], a[2]);
            a[3] = fmaf(w0.w, Pa[i], a[3]); a[4] = fmaf(w1.x, Pa[i], a[4]); a[5] = fmaf(w1.y, Pa[i], a[5]);
            a[6] = fmaf(w1.z, Pa[i], a[6]); a[7] = fmaf(w1.w, Pa[i], a[7]);
        }
    }
    float asum = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) asum += a[q];
    const float meanA = group8_sum(asum) * (1.0f / ECORD_DIM_Q);
    float la = 0.0f, ac[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        ac[q] = a[q] - meanA;
        la = fmaf(w.q2_w[k8 * 8 + q] * w.qln_w[k8 * 8 + q], ac[q], la);
    }
    la = group8_sum(la);
    if (live) {
        float *Ao = p.A + (long long)m * ECORD_DIM_Q + k8 * 8;
        *reinterpret_cast<float4 *>(Ao) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4 *>(Ao + 4) = make_float4(a[4], a[5], a[6], a[7]);
        if (p.Ac) {   // fast Q kernel operands: A - mean_c(A), and its projection on w_o * gamma
            float *Aco = p.Ac + (long long)m * ECORD_DIM_Q + k8 * 8;
            *reinterpret_cast<float4 *>(Aco) = make_float4(ac[0], ac[1], ac[2], ac[3]);
            *reinterpret_cast<float4 *>(Aco + 4) = make_float4(ac[4], ac[5], ac[6], ac[7]);
            if (k8 == 0) p.LA[m] = la;
        }
        if (k8 == 0) p.v[m] = vv;
    }
    if (p.qa) {
        // operands of the tensor-core Q kernel (layout: ecord_policy.qa), a' = Ac + bconst.  Each lane holds 8 channels
        // of its row; the 20 channel sums (Wg^T a', Cc^T a', |a'|^2) are reduced over the row's 8 lanes.
        float part[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) part[i] = 0.0f;
        float ga[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = k8 * 8 + q;
            const float ap = ac[q] + S.cc[c * 4 + 3];
            ga[q] = w.qln_w[c] * ap;
#pragma unroll
            for (int k = 0; k < 16; ++k) part[k] = fmaf(S.wg[c * 16 + k], ap, part[k]);
#pragma unroll
            for (int j = 0; j < 3; ++j) part[16 + j] = fmaf(S.cc[c * 4 + j], ap, part[16 + j]);
            part[19] = fmaf(ap, ap, part[19]);
        }
#pragma unroll
        for (int i = 0; i < 20; ++i) part[i] = group8_sum(part[i]);
        if (live) {
            // (hi, lo) pairs: hi on the tf32 grid, lo the exact remainder -- the kernel stores them as they are
            float *qo = p.qa + (long long)m * ECORD_QA_STRIDE;
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                const float h0 = tf32_round(ga[q]), h1 = tf32_round(ga[q + 1]);
                *reinterpret_cast<float4 *>(qo + 2 * (k8 * 8 + q)) = make_float4(h0, ga[q] - h0, h1, ga[q + 1] - h1);
            }
#pragma unroll
            for (int i = 0; i < 19; ++i)
                if ((i & 7) == k8) qo[128 + i] = tf32_round(part[i]);
            if (k8 == 3) qo[147] = part[19];
            if (k8 == 4) qo[148] = la;
            if (k8 == 5) qo[149] = vv;
            if (k8 == 6) qo[150] = 0.0f;
            if (k8 == 7) qo[151] = 0.0f;
        }
    }
}
}  // namespace

// the hidden buffer holding the current state is ((*step_base) + step_off) & 1, resolved on the device so that a
// captured CUDA graph stays valid when replayed at a later step count
int launch_policy_pre(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                      cudaStream_t st) {
    k_policy_pre<<<ceil_div((long long)p->rows * 32, 256), 256, 0, st>>>(*g, *env, *w, *p);
    ECORD_LAUNCH_CHECK();
    return 0;
}

int launch_policy_step(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                       const int32_t *step_base, int step_off, int gemm_mode, int skip_pre, cudaStream_t st) {
    const int M = p->rows, Mp = p->rows_pad;
    if (!skip_pre) {
        int rc = launch_policy_pre(g, env, w, p, st);
        if (rc) return rc;
    }
    if (gemm_mode == ECORD_GEMM_FP32) {
        const float *h0 = p->hidden, *h1 = p->hidden + (long long)Mp * H;
        k_sgemm_nt_bias<<<dim3(3072 / 64, Mp / 64), 256, 0, st>>>(p->u, p->u, nullptr, 0, w->gru_wih, w->gru_bih, p->gates,
                                                                   MSG, 6144);
        ECORD_LAUNCH_CHECK();
        k_sgemm_nt_bias<<<dim3(3072 / 64, Mp / 64), 256, 0, st>>>(h0, h1, step_base, step_off, w->gru_whh, w->gru_bhh,
                                                                   p->gates + 3072, H, 6144);
        ECORD_LAUNCH_CHECK();
        k_gru_gates<<<ceil_div((long long)M * H, 256), 256, 0, st>>>(*p, step_base, step_off);
        ECORD_LAUNCH_CHECK();
        k_sgemm_nt_bias<<<dim3(P1 / 64, Mp / 64), 256, 0, st>>>(p->th, p->th, nullptr, 0, w->p1_w, w->p1_b, p->y1, H, P1);
        ECORD_LAUNCH_CHECK();
    } else {
        int rc = launch_gru_tc_ex(w, p, step_base, step_off, st);
        if (rc) return rc;
        rc = launch_proj1_tc(w, p, st);
        if (rc) return rc;
    }
    k_policy_tail<<<ceil_div(M, kTailRows), kTailWarps * 32, sizeof(TailSmem), st>>>(*w, *p);
    ECORD_LAUNCH_CHECK();
    return 0;
}

int policy_kernels_per_step(int gemm_mode, int skip_pre) { return (gemm_mode == ECORD_GEMM_FP32 ? 6 : 4) - (skip_pre ? 1 : 0); }

int policy_init_attrs() {
    ECORD_CUDA(cudaFuncSetAttribute(k_policy_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
    return tc_init_attrs();
}

int check_policy(const ecord_policy *p, int gemm_mode) {
    ECORD_RET_IF(!p, ECORD_E_NULL);
    ECORD_RET_IF(p->rows <= 0 || p->rows_pad < p->rows || (p->rows_pad % 128) != 0, ECORD_E_SIZE);
    ECORD_RET_IF(!p->hidden || !p->u || !p->y1 || !p->P || !p->A || !p->v || !p->last_sel, ECORD_E_NULL);
    if (gemm_mode == ECORD_GEMM_FP32) ECORD_RET_IF(!p->gates || !p->th, ECORD_E_NULL);
    else ECORD_RET_IF(!p->hidden_bf16 || !p->th_bf16 || !p->u_bf16, ECORD_E_NULL);
    return 0;
}

extern "C" int ecord_policy_pre(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                                void *stream) {
    ECORD_RET_IF(!g || !env || !w || !p, ECORD_E_NULL);
    ECORD_RET_IF(!p->u || !p->last_sel || !w->msg_w || !w->msg_b, ECORD_E_NULL);
    ECORD_RET_IF(p->rows != g->num_graphs * env->num_tradj, ECORD_E_SIZE);
    return launch_policy_pre(g, env, w, p, as_stream(stream));
}

extern "C" int ecord_policy_step(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                                 int32_t cur, int32_t gemm_mode, int32_t skip_pre, void *stream) {
    ECORD_RET_IF(!g || !env || !w, ECORD_E_NULL);
    ECORD_RET_IF(gemm_mode != ECORD_GEMM_FP32 && gemm_mode != ECORD_GEMM_BF16, ECORD_E_UNSUPPORTED);
    ECORD_RET_IF(cur != 0 && cur != 1, ECORD_E_SIZE);
    int rc = check_policy(p, gemm_mode);
    if (rc) return rc;
    ECORD_RET_IF(p->rows != g->num_graphs * env->num_tradj, ECORD_E_SIZE);
    if (gemm_mode == ECORD_GEMM_BF16) ECORD_RET_IF(!w->gru_w_bf16 || !w->p1_w_bf16, ECORD_E_NULL);
    if ((rc = policy_init_attrs())) return rc;
    ECORD_RET_IF((p->Ac == nullptr) != (p->LA == nullptr), ECORD_E_NULL);
    ECORD_RET_IF(p->qa && !w->q_tc, ECORD_E_NULL);
    return launch_policy_step(g, env, w, p, nullptr, cur, gemm_mode, skip_pre, as_stream(stream));
}
