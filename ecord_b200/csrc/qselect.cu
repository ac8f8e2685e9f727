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

namespace {
constexpr int Q = ECORD_DIM_Q;   // 64

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
           const int32_t *__restrict__ forced) {
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

int launch_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                    const int32_t *step_base, int step_off, float tau, uint64_t seed, int compat, const int32_t *forced,
                    cudaStream_t st) {
    QConst K;
    static_assert(sizeof(QConst) == 6 * Q * sizeof(float), "QConst must be densely packed");
    memcpy(&K, w->q_fold, sizeof(QConst));      // q_fold is a HOST array (see ecord_weights)
    k_q_select<<<p->rows, q_block_size(g->n_max), 0, st>>>(*g, *env, *w, *p, K, step_base, step_off, tau, seed, compat,
                                                           forced);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                              int32_t steps_done, float tau, uint64_t seed, int32_t compat, const int32_t *forced_actions,
                              void *stream) {
    ECORD_RET_IF(!g || !env || !w || !p, ECORD_E_NULL);
    ECORD_RET_IF(!p->A || !p->v || !p->node_B || !p->node_emb || !p->last_sel || !p->actions || !w->q_fold, ECORD_E_NULL);
    ECORD_RET_IF(p->rows != g->num_graphs * env->num_tradj || steps_done < 0, ECORD_E_SIZE);
    ECORD_RET_IF(tau < 0.0f, ECORD_E_SIZE);
    ECORD_RET_IF(tau > 0.0f && !p->q, ECORD_E_NULL);
    return launch_q_select(g, env, w, p, nullptr, steps_done, tau, seed, compat, forced_actions, as_stream(stream));
}

extern "C" int ecord_plane_to_nt(const ecord_graph *g, int32_t num_tradj, const float *plane, float *out, void *stream) {
    ECORD_RET_IF(!g || !plane || !out, ECORD_E_NULL);
    ECORD_RET_IF(num_tradj <= 0, ECORD_E_SIZE);
    const long long NT = (long long)g->num_nodes * num_tradj;
    k_plane_to_nt<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, num_tradj, plane, out);
    ECORD_LAUNCH_CHECK();
    return 0;
}
