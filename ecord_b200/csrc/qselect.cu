// qselect.cu -- per-vertex Q-values, action selection and the chosen-embedding cache
// (SURVEY.md 8a rows a7, a8 (per-vertex half), a9; "Subsystem 2b/2c").
//
// Replaces, per step:
//   _ActorBase._get_node_embeddings / Actor_Q._get_node_features   actors/_base.py:279-294, actors/ecord.py:78-91
//   _RNNDecoderBase._forward_action_value (mlp_out over [P[g,t]; emb[n,t]])   models/rnn_decoder.py:285-311,115-119
//   Actor_Q.action_select (scatter_max / scatter_softmax_sample, offset removal, embedding cache)
//                                                                  actors/ecord.py:123-193,12-19; _base.py:296-308
//
// The reference materialises emb[N,T,32] and inp[N,T,64] and runs a 64x64 Linear per (vertex, trajectory)
// row.  mlp_out[0] is linear in its concatenated input, so it folds exactly (up to fp32 re-association) into
//     pre[n,t,:] = A[g,t,:] + B[n,:] + C f[n,t,:]          A = Wq[:, :32] P[g,t] + b_q   (policy.cu, per step)
//                                                          B = Wq[:,32:48] node_emb[n] + Wq[:,48:64] b_enc (gnn.cu, per rollout)
//                                                          C = Wq[:,48:64] W_enc (64x3)  (ecord_pack_weights)
// with f = (state, peek/n_g, min(static,40)/40) the canonical observation (tradjectories.py:336-353).
// What remains per row is LayerNorm(64) -> LeakyReLU -> dot(64): ~10 fp32 ops per channel on CUDA cores,
// with 8 bytes of state read per row; nothing [N,T,*]-sized is ever written.
//
// One CTA owns one (graph, trajectory) unit: its threads stride over the unit's contiguous state row
// (coalesced), each thread keeps the running (q, index) maximum, a packed 64-bit key makes the block
// reduction a plain max with "lowest index wins on ties" (torch_scatter CPU semantics).
#include "common.cuh"
#include <string.h>
#include <math.h>

namespace {
constexpr int Q = ECORD_DIM_Q;   // 64

// epsilon-greedy override (actors/ecord.py:171-180): with probability epsilon the unit's action becomes a uniformly random
// vertex of its graph, floor(U * n_g).  One Philox draw per (unit, step); the stream differs from torch.rand, the
// distribution is the reference's.
__device__ __forceinline__ int eps_greedy(int a, float epsilon, uint64_t seed, int m, int steps_done, int n) {
    if (epsilon <= 0.0f) return a;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)m, (uint32_t)steps_done, 0x45505347u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    if (uniform01_open(r.x) >= epsilon) return a;
    return min((int)(((float)(r.y >> 8) * (1.0f / 16777216.0f)) * (float)n), n - 1);
}

// monotone float -> uint32 map (larger float <=> larger uint)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key(float val, int idx) {
    return ((unsigned long long)f2ord(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ unsigned long long block_max_key(unsigned long long k, unsigned long long *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();            // protect `red` against the previous use
    if (l == 0) red[w] = k;
    __syncthreads();
    k = red[0];
    for (int i = 1; i < nw; ++i) k = red[i] > k ? red[i] : k;
    return k;
}

// Per-channel constants of the folded head.  Passed BY VALUE as a kernel parameter: parameters live in the
// constant bank, and with the channel loops fully unrolled every use becomes a constant-bank operand of the
// FFMA itself (no load instruction, no register).  Layout = ecord_weights.q_fold (host array).
struct QConst { float C0[Q], C1[Q], C2[Q], gam[Q], bet[Q], wo[Q]; };

struct QSmem {
    float A[Q];
    unsigned long long red[32];
};

// canonical observation of one vertex from the SoA state
__device__ __forceinline__ void features(const ecord_env &e, long long o, int n, int steps_done, float &f0, float &f1,
                                         float &f2) {
    const int sp = e.sp[o];
    f0 = (float)(sp & 1);
    f1 = e.peek[o] / (float)n;                                                           // PEEK_NORM
    f2 = fminf((float)(steps_done - (sp >> 1)), (float)ECORD_STATIC_CLIP) / (float)ECORD_STATIC_CLIP;  // STEPS_STATIC_CLIP
}

__device__ __forceinline__ float q_row(const QSmem &S, const QConst &K, const float *__restrict__ Brow, float f0, float f1,
                                       float f2) {
    float x[Q];
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < Q; c += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(Brow + c);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = S.A[c + i] + bb[i];
            v = fmaf(K.C0[c + i], f0, v);
            v = fmaf(K.C1[c + i], f1, v);
            v = fmaf(K.C2[c + i], f2, v);
            x[c + i] = v;
            s += v;
        }
    }
    const float mean = s * (1.0f / Q);
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < Q; ++c) { x[c] -= mean; ss = fmaf(x[c], x[c], ss); }
    const float rstd = rsqrtf(ss * (1.0f / Q) + kLnEps);
    float dot = 0.0f;
#pragma unroll
    for (int c = 0; c < Q; ++c) dot = fmaf(K.wo[c], lrelu(fmaf(x[c] * rstd, K.gam[c], K.bet[c])), dot);
    return dot;
}

__global__ void __launch_bounds__(256)
k_q_select(ecord_graph g, ecord_env e, ecord_weights w, ecord_policy p, const __grid_constant__ QConst K,
           const int32_t *step_base, int step_off, float tau, uint64_t seed, int compat,
           const int32_t *__restrict__ forced, float epsilon) {
    __shared__ QSmem S;
    const int m = blockIdx.x;
    const int T = e.num_tradj;
    const int gi = m / T, t = m - gi * T;
    const int steps_done = (step_base ? *step_base : 0) + step_off;
    for (int c = threadIdx.x; c < Q; c += blockDim.x) S.A[c] = p.A[(long long)m * Q + c];
    __syncthreads();
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    const int lo = g.graph_ptr[gi];
    const float bo = w.q2_b[0], vm = p.v[m];

    unsigned long long key = 0ull;    // below every real key (f2ord(-inf) > 0)
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        float f0, f1, f2;
        features(e, base + v, n, steps_done, f0, f1, f2);
        const float dot = q_row(S, K, p.node_B + (long long)(lo + v) * Q, f0, f1, f2);
        const float q = (dot + bo) + vm;              // action_val (+bias) then + state_val, as the reference rounds it
        if (p.q) p.q[base + v] = q;
        const unsigned long long k = pack_key(q, v);
        key = k > key ? k : key;
    }
    key = block_max_key(key, S.red);
    int a = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));

    if (tau > 0.0f) {
        // exponential race (actors/ecord.py:12-19): argmin_n  -log(U_n) / exp((q_n - max q) / tau)
        const uint32_t ord = (uint32_t)(key >> 32);
        const float qmax = __uint_as_float((ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord);
        unsigned long long k2 = 0ull;
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            const float q = p.q[base + v];
            const uint4 r = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)m, (uint32_t)steps_done, 0x45434F52u),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const float z = -logf(uniform01_open(r.x)) / expf(q / tau - qmax / tau);
            const unsigned long long k = pack_key(-z, v);       // min z == max -z
            k2 = k > k2 ? k : k2;
        }
        k2 = block_max_key(k2, S.red);
        a = (int)(0xFFFFFFFFu - (uint32_t)(k2 & 0xFFFFFFFFull));
    }
    a = eps_greedy(a, epsilon, seed, m, steps_done, n);
    if (forced) a = forced[m];

    // cache the embedding of the chosen vertex for the next GRU update (actors/_base.py:296-308).
    // compat: the reference gathers with the LOCAL index from the FLAT [N,T,32] array (quirk A).
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int j = compat ? a : lo + a;                       // flat vertex whose embedding is cached
        const int gj = g.node_graph[j];
        int nj;
        const long long bj = unit_base(g.graph_ptr, gj, t, T, nj);
        float f0, f1, f2;
        features(e, bj + (j - g.graph_ptr[gj]), nj, steps_done, f0, f1, f2);
        float val;
        if (lane < 16) val = p.node_emb[(long long)j * 16 + lane];
        else {
            const int c = lane - 16;
            val = w.enc_b[c];
            val = fmaf(w.enc_w[c * 3 + 0], f0, val);
            val = fmaf(w.enc_w[c * 3 + 1], f1, val);
            val = fmaf(w.enc_w[c * 3 + 2], f2, val);
        }
        p.last_sel[(long long)m * ECORD_DIM_NODE + lane] = val;
        if (lane == 0) p.actions[m] = a;
    }
}

// =================================================================================================
// ECORD_Q_FAST: the throughput form of the same head.
//
// (1) Work decomposition.  The exact kernel streams B[n] (256 B) from L2 for every (vertex, trajectory) row:
//     N*T*256 B per step.  Here one thread owns ONE vertex, keeps its 64-float operand in registers and loops over
//     a range of trajectories of its graph, so B is read once per (vertex, step); the per-(g,t) operand A comes
//     from shared memory as warp-uniform LDS.128 broadcasts.  Grid = (vertex tiles of 128) x (trajectory splits).
// (2) Arithmetic.  With every operand centred over the 64 channels (Ac, Bc, Cc: rows sum to 0) the LayerNorm mean
//     vanishes identically, and LeakyReLU(y) = a1 y + a2 |y| (a1 = 0.505, a2 = 0.495) splits the head into a part
//     LINEAR in the pre-activation -- which folds into three scalars LA[g,t] + LB[n] + LC.f -- and an |.| part:
//         x_c   = Ac_c + Bc_c + Cc_c . f                         (1 FADD + 3 FFMA)
//         ss   += x_c^2                                          (1 FFMA)      sigma = sqrt(ss/64 + eps)
//         acc  += w_c |gamma_c x_c + beta_c sigma|               (1 FMUL + 2 FFMA; |.| is a free operand modifier)
//         q     = a1 (LA + LB + LC.f)/sigma + a1 sum(w beta) + a2 acc/sigma + b_o + v[g,t]
//     8 fp32 ops per channel instead of ~12: the same function re-associated (differences ~1e-6 relative).
// (3) Arg-max across the CTAs that share a (g,t): warp REDUX.MAX on the order-preserving integer image of q,
//     a ballot picks the lowest vertex among equals, one 64-bit atomicMax per warp on keys[g*T+t].
// =================================================================================================
struct QConstFast { float C0[Q], C1[Q], C2[Q], gam[Q], bet[Q], wo[Q]; float LC[3], Wb; };
constexpr int QT = ECORD_QTILE;      // vertices per CTA = threads per CTA
constexpr int kMaxTChunk = 32;       // trajectories per CTA (shared-memory A rows)

struct QFastSmem { float A[kMaxTChunk][Q]; float LA[kMaxTChunk], V[kMaxTChunk]; };

template <bool DUMP_Q>
__global__ void __launch_bounds__(QT, 3)
k_q_fast(ecord_graph g, ecord_env e, ecord_policy p, const __grid_constant__ QConstFast K, const float *__restrict__ q2_b,
         const int32_t *step_base, int step_off, int t_chunk) {
    __shared__ __align__(16) QFastSmem S;
    const int tile = blockIdx.x;
    const int gi = g.qtile_graph[tile], v0 = g.qtile_v0[tile];
    const int T = e.num_tradj;
    const int t_lo = blockIdx.y * t_chunk, t_hi = min(T, t_lo + t_chunk);
    const int steps_done = (step_base ? *step_base : 0) + step_off;
    const int lo = g.graph_ptr[gi], n = g.graph_ptr[gi + 1] - lo;
    const int lane = threadIdx.x & 31;
    // per-(g,t) operands of this CTA's trajectory range -> shared memory
    for (int i = threadIdx.x; i < (t_hi - t_lo) * (Q / 4); i += QT) {
        const int tt = i / (Q / 4), c4 = i % (Q / 4);
        reinterpret_cast<float4 *>(S.A[tt])[c4] =
            reinterpret_cast<const float4 *>(p.Ac + (long long)(gi * T + t_lo + tt) * Q)[c4];
    }
    for (int i = threadIdx.x; i < t_hi - t_lo; i += QT) {
        S.LA[i] = p.LA[gi * T + t_lo + i];
        S.V[i] = p.v[gi * T + t_lo + i];
    }
    // per-vertex operand -> registers (kept across the trajectory loop)
    const bool valid = v0 + (int)threadIdx.x < n;
    const int v = valid ? v0 + (int)threadIdx.x : n - 1;
    float b[Q];
    {
        const float4 *bp = reinterpret_cast<const float4 *>(p.node_Bc + (long long)(lo + v) * Q);
#pragma unroll
        for (int c = 0; c < Q / 4; ++c) {
            const float4 t4 = bp[c];
            b[4 * c] = t4.x; b[4 * c + 1] = t4.y; b[4 * c + 2] = t4.z; b[4 * c + 3] = t4.w;
        }
    }
    const float lb = p.node_LB[lo + v];
    const float bo = q2_b[0];
    // reciprocal multiplies instead of the IEEE divisions of the exact kernel (<= 1 ulp apart): a division per
    // row costs ~30 instructions and its slow path is taken whenever peek == 0
    const float inv_n = 1.0f / (float)n;
    constexpr float inv_clip = 1.0f / (float)ECORD_STATIC_CLIP;
    __syncthreads();

    // element (t_lo, v) of this graph's planes; consecutive trajectories are n elements apart
    const long long base0 = (long long)lo * T + (long long)t_lo * n + v;
    const int *__restrict__ sp_ptr = e.sp + base0;
    const float *__restrict__ pk_ptr = e.peek + base0;
    float *q_ptr = DUMP_Q ? p.q + base0 : nullptr;
    unsigned long long *key_ptr = p.keys + (gi * T + t_lo);
    const int nt = t_hi - t_lo;
    int sp = *sp_ptr;
    float pk = *pk_ptr;
    for (int tt = 0; tt < nt; ++tt) {
        const int sp_c = sp;
        const float pk_c = pk;
        if (tt + 1 < nt) {                        // prefetch the next trajectory's state row entry
            sp_ptr += n; pk_ptr += n;
            sp = *sp_ptr;
            pk = *pk_ptr;
        }
        const float f0 = (float)(sp_c & 1);
        const float f1 = pk_c * inv_n;
        const float f2 = fminf((float)(steps_done - (sp_c >> 1)), (float)ECORD_STATIC_CLIP) * inv_clip;
        const float *As = S.A[tt];
        float x[Q];
        float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
#pragma unroll
        for (int c = 0; c < Q; c += 4) {
            const float4 a4 = *reinterpret_cast<const float4 *>(As + c);
            x[c + 0] = fmaf(K.C2[c + 0], f2, fmaf(K.C1[c + 0], f1, fmaf(K.C0[c + 0], f0, a4.x + b[c + 0])));
            x[c + 1] = fmaf(K.C2[c + 1], f2, fmaf(K.C1[c + 1], f1, fmaf(K.C0[c + 1], f0, a4.y + b[c + 1])));
            x[c + 2] = fmaf(K.C2[c + 2], f2, fmaf(K.C1[c + 2], f1, fmaf(K.C0[c + 2], f0, a4.z + b[c + 2])));
            x[c + 3] = fmaf(K.C2[c + 3], f2, fmaf(K.C1[c + 3], f1, fmaf(K.C0[c + 3], f0, a4.w + b[c + 3])));
            ss0 = fmaf(x[c + 0], x[c + 0], ss0); ss1 = fmaf(x[c + 1], x[c + 1], ss1);
            ss2 = fmaf(x[c + 2], x[c + 2], ss2); ss3 = fmaf(x[c + 3], x[c + 3], ss3);
        }
        const float var = ((ss0 + ss1) + (ss2 + ss3)) * (1.0f / Q);
        const float sigma = sqrtf(var + kLnEps);
        const float rstd = 1.0f / sigma;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int c = 0; c < Q; c += 4) {
            a0 = fmaf(K.wo[c + 0], fabsf(fmaf(x[c + 0], K.gam[c + 0], K.bet[c + 0] * sigma)), a0);
            a1 = fmaf(K.wo[c + 1], fabsf(fmaf(x[c + 1], K.gam[c + 1], K.bet[c + 1] * sigma)), a1);
            a2 = fmaf(K.wo[c + 2], fabsf(fmaf(x[c + 2], K.gam[c + 2], K.bet[c + 2] * sigma)), a2);
            a3 = fmaf(K.wo[c + 3], fabsf(fmaf(x[c + 3], K.gam[c + 3], K.bet[c + 3] * sigma)), a3);
        }
        const float acc = (a0 + a1) + (a2 + a3);
        const float lin = S.LA[tt] + lb + fmaf(K.LC[2], f2, fmaf(K.LC[1], f1, K.LC[0] * f0));
        const float dot = fmaf(0.505f, fmaf(lin, rstd, K.Wb), 0.495f * acc * rstd);
        const float q = (dot + bo) + S.V[tt];
        if (DUMP_Q) { if (valid) *q_ptr = q; q_ptr += n; }
        // arg-max: ordered-int max across the warp, lowest vertex among equals, one atomic per warp
        const uint32_t ord = valid ? f2ord(q) : 0u;
        const uint32_t mx = __reduce_max_sync(0xffffffffu, ord);
        const uint32_t bal = __ballot_sync(0xffffffffu, ord == mx);
        if (lane == 0 && mx != 0u) {
            const int vbest = v0 + (int)(threadIdx.x & ~31u) + (__ffs(bal) - 1);
            atomicMax(key_ptr + tt, ((unsigned long long)mx << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)vbest));
        }
    }
}

// one warp per (g,t): decode the arg-max key, cache the chosen vertex's embedding, reset the key
__global__ void __launch_bounds__(256)
k_select_finish(ecord_graph g, ecord_env e, ecord_weights w, ecord_policy p, const int32_t *step_base, int step_off,
                int compat, const int32_t *__restrict__ forced, float epsilon, uint64_t seed) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    if (m >= p.rows) return;
    const int T = e.num_tradj;
    const int gi = m / T, t = m - gi * T;
    const int steps_done = (step_base ? *step_base : 0) + step_off;
    const unsigned long long key = p.keys[m];
    int a = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const int lo = g.graph_ptr[gi];
    a = eps_greedy(a, epsilon, seed, m, steps_done, g.graph_ptr[gi + 1] - lo);
    if (forced) a = forced[m];
    const int j = compat ? a : lo + a;
    const int gj = g.node_graph[j];
    int nj;
    const long long bj = unit_base(g.graph_ptr, gj, t, T, nj);
    float f0, f1, f2;
    features(e, bj + (j - g.graph_ptr[gj]), nj, steps_done, f0, f1, f2);
    float val;
    if (lane < 16) val = p.node_emb[(long long)j * 16 + lane];
    else {
        const int c = lane - 16;
        val = w.enc_b[c];
        val = fmaf(w.enc_w[c * 3 + 0], f0, val);
        val = fmaf(w.enc_w[c * 3 + 1], f1, val);
        val = fmaf(w.enc_w[c * 3 + 2], f2, val);
    }
    p.last_sel[(long long)m * ECORD_DIM_NODE + lane] = val;
    __syncwarp();
    if (lane == 0) { p.actions[m] = a; p.keys[m] = 0ull; }
}

// ------------------------------------------------------------------------------------------------
// Greedy local search on the peeks (reference solvers/solver.py:38-95 greedy_solve).
// One CTA per (g,t): the action is drawn by the reference's exponential race on peek / max(tau, 1e-9)
// (solver.py:64 + scatter_softmax_sample): at tau = 0 every vertex below the maximum gets exp(-inf) = 0, so the draw is
// uniform over the vertices of maximal peek.  at_local_min[m] = (max peek <= 0) for the state BEFORE the flip.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_greedy_select(ecord_graph g, ecord_env e, float inv_tau, uint64_t seed, int step_no, int32_t *__restrict__ actions,
                int32_t *__restrict__ at_local_min) {
    __shared__ unsigned long long red[32];
    const int m = blockIdx.x;
    const int T = e.num_tradj;
    const int gi = m / T, t = m - gi * T;
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    unsigned long long key = 0ull;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const unsigned long long k = pack_key(e.peek[base + v], v);
        key = k > key ? k : key;
    }
    key = block_max_key(key, red);
    const uint32_t ord = (uint32_t)(key >> 32);
    const float pmax = __uint_as_float((ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord);
    unsigned long long k2 = 0ull;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const float x = (e.peek[base + v] - pmax) * inv_tau;                    // <= 0; 0 exactly for the maxima
        const uint4 r = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)m, (uint32_t)step_no, 0x47524459u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float z = -logf(uniform01_open(r.x)) / expf(x);                   // inf where exp underflows
        const unsigned long long k = pack_key(-z, v);
        k2 = k > k2 ? k : k2;
    }
    k2 = block_max_key(k2, red);
    if (threadIdx.x == 0) {
        actions[m] = (int)(0xFFFFFFFFu - (uint32_t)(k2 & 0xFFFFFFFFull));
        if (at_local_min) at_local_min[m] = pmax <= 0.0f ? 1 : 0;
    }
}

// plane [N*T] (trajectory-major) -> reference layout [N][T]
__global__ void k_plane_to_nt(ecord_graph g, int T, const float *__restrict__ plane, float *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.num_nodes * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    const int gi = g.node_graph[v];
    int n;
    out[idx] = plane[unit_base(g.graph_ptr, gi, t, T, n) + (v - g.graph_ptr[gi])];
}
}  // namespace

static int q_block_size(int n_max) { return n_max <= 32 ? 32 : n_max <= 64 ? 64 : n_max <= 128 ? 128 : 256; }

int q_kernels_per_step(int q_mode) { return q_mode == ECORD_Q_EXACT ? 1 : 2; }

// qtc.cu
int launch_q_tc(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                const int32_t *step_base, int step_off, float tau, uint64_t seed, cudaStream_t st);
int qtc_init_attrs();

static int launch_select_finish(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                                const int32_t *step_base, int step_off, int compat, const int32_t *forced, float epsilon,
                                uint64_t seed, cudaStream_t st) {
    ECORD_PDL_LAUNCH(k_select_finish, dim3(ceil_div((long long)p->rows * 32, 256)), dim3(256), (size_t)0, st, *g, *env, *w, *p,
                     step_base, step_off, compat, forced, epsilon, seed);
    return 0;
}

static int launch_q_fast(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                         const int32_t *step_base, int step_off, int compat, const int32_t *forced, float epsilon,
                         uint64_t seed, cudaStream_t st) {
    QConstFast K;
    memcpy(K.C0, w->q_fold + 384, 3 * Q * sizeof(float));          // centred C planes
    memcpy(K.gam, w->q_fold + 3 * Q, 3 * Q * sizeof(float));        // gamma | beta | w_o
    memcpy(K.LC, w->q_fold + 576, 4 * sizeof(float));               // LC[3], Wb
    const int T = env->num_tradj;
    // trajectory splits: enough CTAs for >= ~3 waves of (148 SMs x 3 CTAs), at most kMaxTChunk rows of A in smem
    int splits = ceil_div(3 * 3 * kNumSMs, g->num_qtiles);
    if (splits > T) splits = T;
    int t_chunk = ceil_div(T, splits);
    if (t_chunk > kMaxTChunk) t_chunk = kMaxTChunk;
    splits = ceil_div(T, t_chunk);
    if (p->q) k_q_fast<true><<<dim3(g->num_qtiles, splits), QT, 0, st>>>(*g, *env, *p, K, w->q2_b, step_base, step_off, t_chunk);
    else k_q_fast<false><<<dim3(g->num_qtiles, splits), QT, 0, st>>>(*g, *env, *p, K, w->q2_b, step_base, step_off, t_chunk);
    ECORD_LAUNCH_CHECK();
    return launch_select_finish(g, env, w, p, step_base, step_off, compat, forced, epsilon, seed, st);
}

int launch_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                    const int32_t *step_base, int step_off, float tau, uint64_t seed, int compat, const int32_t *forced,
                    int q_mode, float epsilon, cudaStream_t st) {
    if (q_mode == ECORD_Q_FAST) return launch_q_fast(g, env, w, p, step_base, step_off, compat, forced, epsilon, seed, st);
    if (q_mode == ECORD_Q_TC) {
        int rc = launch_q_tc(g, env, w, p, step_base, step_off, tau, seed, st);
        if (rc) return rc;
        return launch_select_finish(g, env, w, p, step_base, step_off, compat, forced, epsilon, seed, st);
    }
    QConst K;
    static_assert(sizeof(QConst) == 6 * Q * sizeof(float), "QConst must be densely packed");
    memcpy(&K, w->q_fold, sizeof(QConst));      // q_fold is a HOST array (see ecord_weights)
    k_q_select<<<p->rows, q_block_size(g->n_max), 0, st>>>(*g, *env, *w, *p, K, step_base, step_off, tau, seed, compat,
                                                           forced, epsilon);
    ECORD_LAUNCH_CHECK();
    return 0;
}

// argument checks shared by ecord_q_select and ecord_rollout_plan_create
int check_q_mode(const ecord_graph *g, const ecord_weights *w, const ecord_policy *p, int q_mode, float tau) {
    ECORD_RET_IF(q_mode != ECORD_Q_EXACT && q_mode != ECORD_Q_FAST && q_mode != ECORD_Q_TC, ECORD_E_UNSUPPORTED);
    if (q_mode == ECORD_Q_EXACT) return 0;
    ECORD_RET_IF(tau > 0.0f && q_mode != ECORD_Q_TC, ECORD_E_UNSUPPORTED);     // sampling: exact (exponential race) or tensor-core (Gumbel-max) kernel
    ECORD_RET_IF(!p->keys, ECORD_E_NULL);
    ECORD_RET_IF(g->num_qtiles <= 0 || !g->qtile_graph || !g->qtile_v0, ECORD_E_NULL);
    if (q_mode == ECORD_Q_FAST) ECORD_RET_IF(!p->Ac || !p->LA || !p->node_Bc || !p->node_LB, ECORD_E_NULL);
    if (q_mode == ECORD_Q_TC) {
        ECORD_RET_IF(!p->qa || !p->node_qv || !w->q_tc, ECORD_E_NULL);
        return qtc_init_attrs();
    }
    return 0;
}

extern "C" int ecord_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                              int32_t steps_done, float tau, uint64_t seed, int32_t compat, const int32_t *forced_actions,
                              int32_t q_mode, float epsilon, void *stream) {
    ECORD_RET_IF(!g || !env || !w || !p, ECORD_E_NULL);
    ECORD_RET_IF(!p->A || !p->v || !p->node_emb || !p->last_sel || !p->actions || !w->q_fold, ECORD_E_NULL);
    ECORD_RET_IF(!p->node_B && q_mode != ECORD_Q_TC, ECORD_E_NULL);
    ECORD_RET_IF(p->rows != g->num_graphs * env->num_tradj || steps_done < 0, ECORD_E_SIZE);
    ECORD_RET_IF(tau < 0.0f || !(epsilon >= 0.0f && epsilon <= 1.0f), ECORD_E_SIZE);
    ECORD_RET_IF(tau > 0.0f && q_mode == ECORD_Q_EXACT && !p->q, ECORD_E_NULL);     // the exact kernel's second pass re-reads q
    int rc = check_q_mode(g, w, p, q_mode, tau);
    if (rc) return rc;
    return launch_q_select(g, env, w, p, nullptr, steps_done, tau, seed, compat, forced_actions, q_mode, epsilon,
                           as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// Network initialisation (solvers/solver.py:256-277, actors/_base.py:238-245, models/node_classifier.py): the initial
// spins are drawn from Bernoulli(logits = heads[h](gnn_node_embeddings)), one Linear(16 -> 1) head per probability
// map.  With more than one head the reference additionally re-draws each spin uniformly with probability epsilon
// (the perturbation sits inside the multi-head branch, solver.py:268-273).  One thread per (vertex, trajectory);
// output in the reference's [N][T] layout, the form ecord_env_init takes.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_node_init_sample(int N, int T, int t_off, const float *__restrict__ gnn_out, const float *__restrict__ head_w,
                                   const float *__restrict__ head_b, int head, float perturb_eps, uint64_t seed,
                                   int8_t *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    float logit = head_b[head];
#pragma unroll
    for (int k = 0; k < ECORD_DIM_GNN; ++k) logit = fmaf(gnn_out[(long long)v * ECORD_DIM_GNN + k], head_w[head * ECORD_DIM_GNN + k], logit);
    const float prob = 1.0f / (1.0f + expf(-logit));
    const uint4 r = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)(t + t_off), 0x494E4954u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    int s = uniform01_open(r.x) < prob ? 1 : 0;
    if (perturb_eps > 0.0f && uniform01_open(r.y) < perturb_eps) s = (int)(r.z & 1u);
    out[idx] = (int8_t)s;
}
}  // namespace

extern "C" int ecord_node_init_sample(const ecord_graph *g, int32_t num_tradj, int32_t tradj_offset, const float *gnn_out, const float *head_w,
                                      const float *head_b, int32_t num_heads, int32_t head, float perturb_epsilon,
                                      uint64_t seed, int8_t *init_state_out, void *stream) {
    ECORD_RET_IF(!g || !gnn_out || !head_w || !head_b || !init_state_out, ECORD_E_NULL);
    ECORD_RET_IF(g->num_nodes <= 0 || num_tradj <= 0 || tradj_offset < 0 || num_heads <= 0 || head < 0 || head >= num_heads,
                 ECORD_E_SIZE);
    ECORD_RET_IF(!(perturb_epsilon >= 0.0f && perturb_epsilon <= 1.0f), ECORD_E_SIZE);
    const long long NT = (long long)g->num_nodes * num_tradj;
    k_node_init_sample<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(g->num_nodes, num_tradj, tradj_offset, gnn_out, head_w, head_b, head,
                                                                        num_heads > 1 ? perturb_epsilon : 0.0f, seed,
                                                                        init_state_out);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_greedy_select(const ecord_graph *g, const ecord_env *env, float tau, uint64_t seed, int32_t step_no,
                                   int32_t *actions, int32_t *at_local_min, void *stream) {
    ECORD_RET_IF(!g || !env || !actions, ECORD_E_NULL);
    ECORD_RET_IF(tau < 0.0f || step_no < 0 || env->num_tradj <= 0, ECORD_E_SIZE);
    const float inv_tau = 1.0f / fmaxf(tau, 1e-9f);                 // solver.py:64  peeks / max(tau, 1e-9)
    k_greedy_select<<<g->num_graphs * env->num_tradj, q_block_size(g->n_max), 0, as_stream(stream)>>>(
        *g, *env, inv_tau, seed, step_no, actions, at_local_min);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_plane_to_nt(const ecord_graph *g, int32_t num_tradj, const float *plane, float *out, void *stream) {
    ECORD_RET_IF(!g || !plane || !out, ECORD_E_NULL);
    ECORD_RET_IF(num_tradj <= 0, ECORD_E_SIZE);
    const long long NT = (long long)g->num_nodes * num_tradj;
    k_plane_to_nt<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, num_tradj, plane, out);
    ECORD_LAUNCH_CHECK();
    return 0;
}
