// env.cu -- Max-Cut environment kernels (SURVEY.md 8a rows a2, a3, a4 and seam B3).
//
// Replaces the reference's only native component, _tradjectory_step_cy.pyx:9-171, and the NumPy/
// torch_scatter set-up around it (tradjectories.py:62-148,336-376,387-405).
// All arithmetic is fp32 in the reference's order of operations, so results are bit-identical
// (spins are +-1, so every product below is exact and each update is a single rounding).
//
// Work decomposition: the reference loops serially over (graph, trajectory) and then over the CSR
// row of the flipped vertex.  Here ONE WARP owns one (graph, trajectory) unit; its lanes stride over
// the CSR row (coalesced col/w loads), the neighbour read-modify-writes scatter inside the unit's
// contiguous row of the trajectory-major plane.  HBM/latency-bound: ~20 B per neighbour.
#include "common.cuh"
#include <stdlib.h>
#include <new>

// ------------------------------------------------------------------------------------------------
// a2: initial peeks / scores (tradjectories.py:130-148), sequential per segment like segment_csr.
// ------------------------------------------------------------------------------------------------
__global__ void k_env_init_peek(ecord_graph g, ecord_env e, const int8_t *__restrict__ s0) {
    const int T = e.num_tradj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.num_nodes * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    const float si = (float)(2 * s0[idx] - 1);
    float acc = 0.0f;
    for (int k = g.row_ptr[v]; k < g.row_ptr[v + 1]; ++k) {
        const float sj = (float)(2 * s0[(long long)g.col[k] * T + t] - 1);
        acc += g.w[k] * si * sj;
    }
    const int gi = g.node_graph[v];
    int n;
    const long long o = unit_base(g.graph_ptr, gi, t, T, n) + (v - g.graph_ptr[gi]);
    e.peek[o] = acc;
    e.sp[o] = (int32_t)s0[idx];          // last_flip_step = 0
    e.best_state[o] = (uint8_t)s0[idx];  // tradjectories.py:116
}

// One warp per (g,t): lanes stride over the graph's vertices, each sums its vertices' edge terms, then a warp sum.
// (fp32 sums of +-1 / small-integer weights are exact in any order; the reference's own summation order for
// non-integer weights is a torch segment reduction that nothing pins.)
__global__ void __launch_bounds__(256)
k_env_init_score(ecord_graph g, ecord_env e, const int8_t *__restrict__ s0) {
    const int T = e.num_tradj;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= g.num_graphs * T) return;
    const int gi = m / T, t = m % T;
    float acc = 0.0f;
    for (int v = g.graph_ptr[gi] + lane; v < g.graph_ptr[gi + 1]; v += 32) {
        const int sv = s0[(long long)v * T + t];
        for (int k = g.row_ptr[v]; k < g.row_ptr[v + 1]; ++k) {
            const int su = s0[(long long)g.col[k] * T + t];
            acc += g.w[k] * (float)((sv + su) & 1);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float sc = 0.5f * acc;
        e.score[m] = sc;
        e.best_score[m] = sc;   // tradjectories.py:81
        e.best_step[m] = 0;
    }
}

// The same score from the peek plane k_env_init_peek has just written: with s = +-1 spins,
//     cut = 1/2 sum_{(i,j) directed} w_ij [s_i != s_j] = 1/4 (sum_{directed} w_ij - sum_v peek_v),   peek_v = sum_j w_ij s_i s_j.
// For integer-valued weights every partial sum is an integer below 2^24, so the result is exact in any order -- bit-equal
// to the edge walk above -- and a unit reads its n_g contiguous peeks instead of walking all the edges of its graph
// (ER500: 500 coalesced floats instead of 37 K scattered edge terms per unit; 848 us -> a few us per rollout).
// graph_wsum[g] = sum of the graph's directed edge weights (host, exact); the edge walk stays for non-integer weights.
__global__ void __launch_bounds__(256)
k_env_init_score_peek(ecord_graph g, ecord_env e) {
    const int T = e.num_tradj;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= g.num_graphs * T) return;
    const int gi = m / T, t = m % T;
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    float acc = 0.0f;
    for (int v = lane; v < n; v += 32) acc += e.peek[base + v];
    acc = warp_sum(acc);
    if (lane == 0) {
        const float sc = 0.25f * (g.graph_wsum[gi] - acc);
        e.score[m] = sc;
        e.best_score[m] = sc;
        e.best_step[m] = 0;
    }
}

// Re-initialise spins, peeks and scores from given spins (init_state i8 [N][T]) or, with init_state == NULL, from the
// best-state plane -- greedy_solve(init_from_best=True), solvers/solver.py:39-43: scores and node features are rebuilt
// (static counters restart: last_flip_step = steps_done), best_score / best_state / best_step are NOT touched.
__global__ void k_env_reset_peek(ecord_graph g, ecord_env e, const int8_t *__restrict__ s0, int steps_done) {
    const int T = e.num_tradj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.num_nodes * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    const int gi = g.node_graph[v];
    int n;
    const long long ub = unit_base(g.graph_ptr, gi, t, T, n);
    const int lo = g.graph_ptr[gi];
    auto spin = [&](int u) -> int { return s0 ? (int)s0[(long long)u * T + t] : (int)e.best_state[ub + (u - lo)]; };
    const int sv = spin(v);
    const float si = (float)(2 * sv - 1);
    float acc = 0.0f;
    for (int k = g.row_ptr[v]; k < g.row_ptr[v + 1]; ++k) acc += g.w[k] * si * (float)(2 * spin(g.col[k]) - 1);
    e.peek[ub + (v - lo)] = acc;
    e.sp[ub + (v - lo)] = steps_done * 2 + sv;
}
__global__ void __launch_bounds__(256)
k_env_reset_score(ecord_graph g, ecord_env e) {          // after k_env_reset_peek: the cut of the spins now in sp
    const int T = e.num_tradj;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= g.num_graphs * T) return;
    const int gi = m / T, t = m % T;
    int n;
    const long long ub = unit_base(g.graph_ptr, gi, t, T, n);
    const int lo = g.graph_ptr[gi];
    float acc = 0.0f;
    for (int v = lo + lane; v < lo + n; v += 32) {
        const int sv = e.sp[ub + (v - lo)] & 1;
        for (int k = g.row_ptr[v]; k < g.row_ptr[v + 1]; ++k)
            acc += g.w[k] * (float)((sv + (e.sp[ub + (g.col[k] - lo)] & 1)) & 1);
    }
    acc = warp_sum(acc);
    if (lane == 0) e.score[m] = 0.5f * acc;
}
// shift the environment clock back by `by` steps (after a greedy pre-solve the rollout starts at step 0 again, but the
// "steps since last flip" features must keep counting from the greedy flips): last_flip_step -= by, best_step -= by
__global__ void k_env_rebase(ecord_graph g, ecord_env e, int by) {
    const long long NT = (long long)g.num_nodes * e.num_tradj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < NT) {
        const int sp = e.sp[idx];
        e.sp[idx] = ((sp >> 1) - by) * 2 + (sp & 1);
    }
    if (idx < (long long)g.num_graphs * e.num_tradj) e.best_step[idx] -= by;
}

// ------------------------------------------------------------------------------------------------
// a3: one step on the SoA state.  One warp per (g,t).
// ------------------------------------------------------------------------------------------------
// ecord_env.true_snapshot: on a new best score the unit's whole spin row is copied into the best-state plane, so that
// plane is the configuration whose cut IS best_score.  (The reference records only the flipped vertex, quirk D --
// _tradjectory_step_cy.pyx:142-146 -- which is what the flag's default, 0, reproduces bit for bit.)  Improvements become
// rare after the first few hundred steps of a rollout, so the n_g-byte copy is paid seldom.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void snapshot_on_new_best(const ecord_env &e, long long base, int n, int new_best, int lane) {
    new_best = __shfl_sync(0xffffffffu, new_best, 0);          // (also orders lane 0's flip before the row copy)
    if (!new_best) return;
    for (int v = lane; v < n; v += 32) e.best_state[base + v] = (uint8_t)(e.sp[base + v] & 1);
}
__global__ void __launch_bounds__(256)
k_env_step(ecord_graph g, ecord_env e, const int32_t *__restrict__ actions, const int32_t *step_base, int step_off,
           float *score_log, int32_t *action_log, int log_stride, unsigned long long *time_log) {
    const int T = e.num_tradj;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= g.num_graphs * T) return;
    const int step_no = (step_base ? *step_base : 0) + step_off;   // steps completed before this one
    const int gi = m / T, t = m - gi * T;
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    const int lo = g.graph_ptr[gi];
    const int a = actions[m];
    const float pa = e.peek[base + a];
    const int spa = e.sp[base + a];
    const int sa = spa & 1;
    // 5(ii): neighbours gain/lose the connecting edge weight (pyx:119-137), PRE-flip state of a
    const float fa = (float)(2 * sa - 1);
    const int k1 = g.row_ptr[lo + a + 1];
    for (int k = g.row_ptr[lo + a] + lane; k < k1; k += 32) {
        const int nb = g.col[k] - lo;
        const float fn = (float)(2 * (e.sp[base + nb] & 1) - 1);
        const float sw = fa * fn * g.w[k];
        e.peek[base + nb] = e.peek[base + nb] - 2.0f * sw;
    }
    __syncwarp();                       // the lanes' reads of a's entries precede lane 0's rewrite of them
    int new_best = 0;
    if (lane == 0) {
        const float sc = e.score[m] + pa;                      // 1. pyx:87
        e.score[m] = sc;
        new_best = sc > e.best_score[m];                       // 2. pyx:90-93
        if (new_best) { e.best_score[m] = sc; e.best_step[m] = step_no + 1; }
        e.peek[base + a] = -1.0f * pa;                         // 5(i). pyx:109
        e.sp[base + a] = ((step_no + 1) << 1) | (sa ^ 1);      // 7 + 9: flip, static = 0
        if (new_best) e.best_state[base + a] = (uint8_t)(sa ^ 1);   // 8. pyx:142-146 (quirk D)
        if (score_log) score_log[(long long)step_no * log_stride + m] = sc;
        if (action_log) action_log[(long long)step_no * log_stride + m] = a;
        if (time_log && m == 0) time_log[step_no] = globaltimer_ns();
    }
    if (e.true_snapshot) snapshot_on_new_best(e, base, n, new_best, lane);
}

// The same step fused with the NEXT step's msg_in stage (policy.cu k_policy_pre): the warp that just applied the
// flip holds everything that stage reads -- the new (best - score) and the cached embedding of the chosen vertex.
__global__ void __launch_bounds__(256)
k_env_step_fused(ecord_graph g, ecord_env e, ecord_weights w, ecord_policy p, const int32_t *step_base, int step_off,
                 float *score_log, int32_t *action_log, int log_stride, unsigned long long *time_log) {
    __shared__ float wt[kMsgWtFloats];
    __shared__ float wb[64];
    msg_in_load(w.msg_w, w.msg_b, wt, wb);
    __syncthreads();
    pdl_wait();                         // the chosen actions come from the selection kernel (a no-op in a plain launch; no
                                        // early trigger: the next kernel is the 148-CTA GRU, whose early residency only
                                        // gets in the way)
    const int T = e.num_tradj;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= g.num_graphs * T) return;
    const int step_no = (step_base ? *step_base : 0) + step_off;
    const int gi = m / T, t = m - gi * T;
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    const int lo = g.graph_ptr[gi];
    const int a = p.actions[m];
    const float pa = e.peek[base + a];
    const int spa = e.sp[base + a];
    const int sa = spa & 1;
    const float fa = (float)(2 * sa - 1);
    const int k1 = g.row_ptr[lo + a + 1];
    for (int k = g.row_ptr[lo + a] + lane; k < k1; k += 32) {
        const int nb = g.col[k] - lo;
        const float fn = (float)(2 * (e.sp[base + nb] & 1) - 1);
        const float sw = fa * fn * g.w[k];
        e.peek[base + nb] = e.peek[base + nb] - 2.0f * sw;
    }
    __syncwarp();                       // the lanes' reads of a's entries precede lane 0's rewrite of them
    float glob0 = 0.0f;
    int new_best = 0;
    if (lane == 0) {
        const float sc = e.score[m] + pa;
        e.score[m] = sc;
        float best = e.best_score[m];
        new_best = sc > best;
        if (new_best) { best = sc; e.best_score[m] = sc; e.best_step[m] = step_no + 1; }
        e.peek[base + a] = -1.0f * pa;
        e.sp[base + a] = ((step_no + 1) << 1) | (sa ^ 1);
        if (new_best) e.best_state[base + a] = (uint8_t)(sa ^ 1);
        if (score_log) score_log[(long long)step_no * log_stride + m] = sc;
        if (action_log) action_log[(long long)step_no * log_stride + m] = a;
        if (time_log && m == 0) time_log[step_no] = globaltimer_ns();     // device clock at the end of step `step_no` (t_step, t_opt)
        glob0 = (best - sc) / (float)n;                      // SCORE_FROM_BEST_NORM of the new observation
    }
    if (e.true_snapshot) snapshot_on_new_best(e, base, n, new_best, lane);
    glob0 = __shfl_sync(0xffffffffu, glob0, 0);
    float u0, u1;
    msg_in_row(wt, wb, p.last_sel[(long long)m * ECORD_DIM_NODE + lane], glob0, lane, u0, u1);
    msg_in_store(p, m, lane, u0, u1);
}

// ------------------------------------------------------------------------------------------------
// a4 / parity view: materialise the reference's arrays.
// ------------------------------------------------------------------------------------------------
__global__ void k_env_export_nodes(ecord_graph g, ecord_env e, int steps_done, int normalise,
                                   float *__restrict__ nf, float *__restrict__ bs) {
    const int T = e.num_tradj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.num_nodes * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    const int gi = g.node_graph[v];
    int n;
    const long long o = unit_base(g.graph_ptr, gi, t, T, n) + (v - g.graph_ptr[gi]);
    const int sp = e.sp[o];
    float peek = e.peek[o];
    float stat = (float)(steps_done - (sp >> 1));
    if (normalise) {
        peek = peek / (float)n;                                            // PEEK_NORM
        stat = fminf(stat, (float)ECORD_STATIC_CLIP) / (float)ECORD_STATIC_CLIP;   // STEPS_STATIC_CLIP
    }
    if (nf) { nf[idx * 3 + 0] = (float)(sp & 1); nf[idx * 3 + 1] = peek; nf[idx * 3 + 2] = stat; }
    if (bs) bs[idx] = (float)e.best_state[o];
}

__global__ void k_env_export_glob(ecord_graph g, ecord_env e, int normalise, float *__restrict__ gf) {
    const int T = e.num_tradj;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.num_graphs * T) return;
    const int gi = m / T;
    float d = e.best_score[m] - e.score[m];                               // pyx:96-97
    if (normalise) d = d / (float)(g.graph_ptr[gi + 1] - g.graph_ptr[gi]); // SCORE_FROM_BEST_NORM
    gf[m * 2 + 0] = d;
    gf[m * 2 + 1] = 0.0f;   // quirk C: PEEK_MAX* is never configured (tradjectories.py:294)
}

// log_stats view (solvers/solver.py:322-327): spins and peeks of every unit in the reference's padded layout
// [G][n_max][T] (tradjectories.py:378-385 _flat2batch; padding -1 for the spins, -1e10 for the peeks).  One thread per
// output element, T innermost: stores are coalesced, the plane reads (stride n_g) come out of L2.
__global__ void k_env_snapshot_padded(ecord_graph g, ecord_env e, float *__restrict__ state_out, float *__restrict__ peek_out) {
    const int T = e.num_tradj;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.num_graphs * g.n_max * T;
    if (idx >= total) return;
    const int t = (int)(idx % T);
    const long long r = idx / T;
    const int v = (int)(r % g.n_max), gi = (int)(r / g.n_max);
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    float s = -1.0f, pk = -1e10f;
    if (v < n) { s = (float)(e.sp[base + v] & 1); pk = e.peek[base + v]; }
    if (state_out) state_out[idx] = s;
    if (peek_out) peek_out[idx] = pk;
}
__global__ void k_stamp_time(unsigned long long *out) { *out = globaltimer_ns(); }

// ------------------------------------------------------------------------------------------------
// B3: step_cy() on the reference's own AoS layout with its full feature-index tables.
// ------------------------------------------------------------------------------------------------
struct StepCyIdx { int state, peek, best, stat, degree; int sfb, ssb, ngr, steps, maxp, meanp; };

__global__ void k_stepcy_prologue(float *nf, float *gf, long long NT, int GT, int Fn, int Fg, int d_static, int g_steps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d_static >= 0 && i < NT) nf[i * Fn + d_static] += 1.0f;            // pyx:160-161
    if (g_steps >= 0 && i < GT) gf[i * Fg + g_steps] += 1.0f;              // pyx:163-164
}

__global__ void __launch_bounds__(256)
k_stepcy(const int64_t *__restrict__ actions, float *scores, float *best_scores, float *best_states,
         float *nf, float *gf, StepCyIdx ix, const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_dst,
         const float *__restrict__ ew, const int32_t *__restrict__ eptr, int G, int T, int Fn, int Fg) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= G * T) return;
    const int t = m % T;
    const long long a = actions[m];
#define NFA(v, d) nf[((long long)(v) * T + t) * Fn + (d)]
    const float pa = NFA(a, ix.peek);
    // neighbours first (they never touch a's own entries; a simple graph has no self loops)
    int dgreedy = 0;
    const int k1 = eptr[a + 1];
    for (int k = eptr[a] + lane; k < k1; k += 32) {
        const int u = e_src[k], v = e_dst[k];
        const float sw = (2.0f * NFA(u, ix.state) - 1.0f) * (2.0f * NFA(v, ix.state) - 1.0f) * ew[k];
        const float old_peek = NFA(v, ix.peek);
        const float new_peek = old_peek - 2.0f * sw;
        if (new_peek > 0.0f && old_peek <= 0.0f) dgreedy += 1;
        else if (new_peek <= 0.0f && old_peek > 0.0f) dgreedy -= 1;
        NFA(v, ix.peek) = new_peek;
    }
    if (ix.ngr >= 0) dgreedy = warp_sum_int(dgreedy);
    __syncwarp();                       // the lanes' reads of a's state precede lane 0's rewrite of it
    if (lane == 0) {
        const float sc = scores[m] + pa;
        scores[m] = sc;
        int new_best = 0;
        if (sc > best_scores[m]) { best_scores[m] = sc; new_best = 1; }
        float *gfm = gf + (long long)m * Fg;
        if (ix.sfb >= 0) gfm[ix.sfb] = best_scores[m] - sc;
        if (ix.ssb >= 0) gfm[ix.ssb] = new_best ? 0.0f : gfm[ix.ssb] + 1.0f;
        const float npa = -1.0f * pa;
        NFA(a, ix.peek) = npa;
        if (ix.ngr >= 0) {
            // integer-valued float counter: +-1 updates are exact, so their order is immaterial
            float c = gfm[ix.ngr];
            if (npa > 0.0f) c += 1.0f;
            if (npa < 0.0f) c -= 1.0f;
            gfm[ix.ngr] = c + (float)dgreedy;
        }
        const float ns = fmodf(NFA(a, ix.state) + 1.0f, 2.0f);
        NFA(a, ix.state) = ns;
        if (new_best) {
            best_states[a * T + t] = ns;
            if (ix.best > 0) NFA(a, ix.best) = ns;      // sic: "> 0", pyx:145
        }
        if (ix.stat >= 0) NFA(a, ix.stat) = 0.0f;
    }
#undef NFA
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int check_graph(const ecord_graph *g) {
    ECORD_RET_IF(!g, ECORD_E_NULL);
    ECORD_RET_IF(g->num_graphs <= 0 || g->num_nodes <= 0 || g->num_edges < 0 || g->n_max <= 0, ECORD_E_SIZE);
    ECORD_RET_IF(!g->graph_ptr || !g->node_graph || !g->row_ptr || (g->num_edges > 0 && (!g->col || !g->w)), ECORD_E_NULL);
    return 0;
}
static int check_env(const ecord_env *e) {
    ECORD_RET_IF(!e, ECORD_E_NULL);
    ECORD_RET_IF(e->num_tradj <= 0, ECORD_E_SIZE);
    ECORD_RET_IF(!e->peek || !e->sp || !e->best_state || !e->score || !e->best_score || !e->best_step, ECORD_E_NULL);
    return 0;
}

extern "C" int ecord_env_init(const ecord_graph *g, const ecord_env *env, const int8_t *init_state, void *stream) {
    int rc;
    if ((rc = check_graph(g)) || (rc = check_env(env))) return rc;
    ECORD_RET_IF(!init_state, ECORD_E_NULL);
    const long long NT = (long long)g->num_nodes * env->num_tradj;
    const int GT = g->num_graphs * env->num_tradj;
    // (a variant transposed through shared memory -- spins read with trajectories innermost, planes written with vertices
    // innermost -- was measured SLOWER, 228 against 154 us at ER500: the kernel is bound by the E x T edge loop, which wants
    // every (vertex, trajectory) pair on its own thread, not by its scattered 4-byte plane stores)
    k_env_init_peek<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, *env, init_state);
    ECORD_LAUNCH_CHECK();
    if (g->int_weights && g->graph_wsum)
        k_env_init_score_peek<<<ceil_div((long long)GT * 32, 256), 256, 0, as_stream(stream)>>>(*g, *env);
    else
        k_env_init_score<<<ceil_div((long long)GT * 32, 256), 256, 0, as_stream(stream)>>>(*g, *env, init_state);
    ECORD_LAUNCH_CHECK();
    return 0;
}

int launch_env_step(const ecord_graph *g, const ecord_env *env, const int32_t *actions, const int32_t *step_base,
                    int step_off, float *score_log, int32_t *action_log, unsigned long long *time_log, cudaStream_t st) {
    const int GT = g->num_graphs * env->num_tradj;
    k_env_step<<<ceil_div((long long)GT * 32, 256), 256, 0, st>>>(*g, *env, actions, step_base, step_off, score_log,
                                                                  action_log, GT, time_log);
    ECORD_LAUNCH_CHECK();
    return 0;
}

int launch_env_step_fused(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                          const int32_t *step_base, int step_off, float *score_log, int32_t *action_log,
                          unsigned long long *time_log, cudaStream_t st) {
    const int GT = g->num_graphs * env->num_tradj;
    // Large grids are launched plainly, small ones as programmatic dependents: letting 600+ small CTAs become resident
    // while the Q kernel still runs costs more than their msg_in weight staging saves (ER500, 625 CTAs: 133.3 -> 131.7 us
    // per step without the early launch), while a grid of about one CTA per SM gains (ER10000, 160 CTAs: 225.7 -> 223.5).
    // Every other kernel of the chain gains from the early launch, 1.8 - 3.6 us each.  (1024-thread CTAs -- a quarter of the CTAs and
    // of the weight staging, launched early -- were measured too: ER500 step 170.1 -> 172.8 us; 512 threads: no change.)
    const int ctas = ceil_div((long long)GT * 32, 256);
    static const int force = [] { const char *e = getenv("ECORD_ENV_PDL"); return e ? atoi(e) : -1; }();   // A/B: 0 plain, 1 PDL
    if (force == 0 || (force < 0 && ctas > 2 * kNumSMs))
        k_env_step_fused<<<dim3(ctas), dim3(256), 0, st>>>(*g, *env, *w, *p, step_base, step_off, score_log, action_log, GT, time_log);
    else
        ECORD_PDL_LAUNCH(k_env_step_fused, dim3(ctas), dim3(256), (size_t)0, st, *g, *env, *w, *p, step_base, step_off, score_log,
                         action_log, GT, time_log);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_env_step_fused(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                                    int32_t step_no, float *score_log, void *stream) {
    int rc;
    if ((rc = check_graph(g)) || (rc = check_env(env))) return rc;
    ECORD_RET_IF(!w || !p || !p->actions || !p->last_sel || !p->u || !w->msg_w || !w->msg_b, ECORD_E_NULL);
    ECORD_RET_IF(step_no < 0 || p->rows != g->num_graphs * env->num_tradj, ECORD_E_SIZE);
    const int GT = g->num_graphs * env->num_tradj;
    float *log0 = score_log ? score_log - (long long)step_no * GT : nullptr;
    return launch_env_step_fused(g, env, w, p, nullptr, step_no, log0, nullptr, nullptr, as_stream(stream));
}

extern "C" int ecord_env_reset_state(const ecord_graph *g, const ecord_env *env, const int8_t *init_state, int32_t steps_done,
                                     void *stream) {
    ECORD_RET_IF(!g || !env, ECORD_E_NULL);
    ECORD_RET_IF(!env->peek || !env->sp || !env->score || (!init_state && !env->best_state), ECORD_E_NULL);
    ECORD_RET_IF(env->num_tradj <= 0 || steps_done < 0, ECORD_E_SIZE);
    const long long NT = (long long)g->num_nodes * env->num_tradj;
    const int GT = g->num_graphs * env->num_tradj;
    k_env_reset_peek<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, *env, init_state, steps_done);
    ECORD_LAUNCH_CHECK();
    k_env_reset_score<<<ceil_div((long long)GT * 32, 256), 256, 0, as_stream(stream)>>>(*g, *env);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_env_rebase_clock(const ecord_graph *g, const ecord_env *env, int32_t by, void *stream) {
    ECORD_RET_IF(!g || !env || !env->sp || !env->best_step, ECORD_E_NULL);
    ECORD_RET_IF(by < 0, ECORD_E_SIZE);
    const long long NT = (long long)g->num_nodes * env->num_tradj;
    k_env_rebase<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, *env, by);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_env_step(const ecord_graph *g, const ecord_env *env, const int32_t *actions, int32_t step_no,
                              float *score_log, void *stream) {
    int rc;
    if ((rc = check_graph(g)) || (rc = check_env(env))) return rc;
    ECORD_RET_IF(!actions, ECORD_E_NULL);
    ECORD_RET_IF(step_no < 0, ECORD_E_SIZE);
    // score_log here is a single [G*T] row: bias the pointer so row `step_no` lands on it
    const int GT = g->num_graphs * env->num_tradj;
    float *log0 = score_log ? score_log - (long long)step_no * GT : nullptr;
    return launch_env_step(g, env, actions, nullptr, step_no, log0, nullptr, nullptr, as_stream(stream));
}

extern "C" int ecord_env_snapshot_padded(const ecord_graph *g, const ecord_env *env, float *state_out, float *peek_out, void *stream) {
    int rc;
    if ((rc = check_graph(g)) || (rc = check_env(env))) return rc;
    ECORD_RET_IF(!state_out && !peek_out, ECORD_E_NULL);
    const long long total = (long long)g->num_graphs * g->n_max * env->num_tradj;
    k_env_snapshot_padded<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(*g, *env, state_out, peek_out);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_stamp_time(uint64_t *out_ns, void *stream) {
    ECORD_RET_IF(!out_ns, ECORD_E_NULL);
    k_stamp_time<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long *>(out_ns));
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_env_export(const ecord_graph *g, const ecord_env *env, int32_t steps_done, int32_t normalise,
                                float *node_features, float *glob_features, float *best_states, void *stream) {
    int rc;
    if ((rc = check_graph(g)) || (rc = check_env(env))) return rc;
    const long long NT = (long long)g->num_nodes * env->num_tradj;
    const int GT = g->num_graphs * env->num_tradj;
    if (node_features || best_states) {
        k_env_export_nodes<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, *env, steps_done, normalise,
                                                                             node_features, best_states);
        ECORD_LAUNCH_CHECK();
    }
    if (glob_features) {
        k_env_export_glob<<<ceil_div(GT, 128), 128, 0, as_stream(stream)>>>(*g, *env, normalise, glob_features);
        ECORD_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int ecord_step_cy_device(const int64_t *actions, float *scores, float *best_scores, float *best_states,
                                    float *node_features, float *global_features, const int32_t *nidx,
                                    const int32_t *gidx, const int32_t *edges_by_node, const float *edge_attrs,
                                    const int32_t *edge_ptr, int32_t N, int32_t G, int32_t T, int32_t Fn, int32_t Fg,
                                    int32_t E, void *stream) {
    ECORD_RET_IF(!actions || !scores || !best_scores || !best_states || !node_features || !nidx || !gidx || !edge_ptr,
                 ECORD_E_NULL);
    ECORD_RET_IF(E > 0 && (!edges_by_node || !edge_attrs), ECORD_E_NULL);
    ECORD_RET_IF(N <= 0 || G <= 0 || T <= 0 || Fn <= 0 || Fg < 0 || E < 0, ECORD_E_SIZE);
    StepCyIdx ix{nidx[0], nidx[1], nidx[2], nidx[3], nidx[4], gidx[0], gidx[1], gidx[2], gidx[3], gidx[4], gidx[5]};
    // the reference asserts STATE and one PEEK variant are configured (tradjectories.py:226-227)
    ECORD_RET_IF(ix.state < 0 || ix.peek < 0 || ix.state >= Fn || ix.peek >= Fn, ECORD_E_UNSUPPORTED);
    ECORD_RET_IF(ix.best >= Fn || ix.stat >= Fn || ix.sfb >= Fg || ix.ssb >= Fg || ix.ngr >= Fg || ix.steps >= Fg,
                 ECORD_E_SIZE);
    const bool any_glob = ix.sfb >= 0 || ix.ssb >= 0 || ix.ngr >= 0 || ix.steps >= 0;
    ECORD_RET_IF(any_glob && !global_features, ECORD_E_NULL);
    cudaStream_t st = as_stream(stream);
    const long long NT = (long long)N * T;
    if (ix.stat >= 0 || ix.steps >= 0) {
        k_stepcy_prologue<<<ceil_div(NT, 256), 256, 0, st>>>(node_features, global_features, NT, G * T, Fn, Fg, ix.stat,
                                                             ix.steps);
        ECORD_LAUNCH_CHECK();
    }
    k_stepcy<<<ceil_div((long long)G * T * 32, 256), 256, 0, st>>>(actions, scores, best_scores, best_states,
                                                                   node_features, global_features, ix, edges_by_node,
                                                                   edges_by_node + E, edge_attrs, edge_ptr, G, T, Fn, Fg);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_step_cy(const int64_t *actions, float *scores, float *best_scores, float *best_states,
                             float *node_features, float *global_features, const int32_t *nidx, const int32_t *gidx,
                             const int32_t *edges_by_node, const float *edge_attrs, const int32_t *edge_ptr, int32_t N,
                             int32_t G, int32_t T, int32_t Fn, int32_t Fg) {
    ECORD_RET_IF(!actions || !scores || !best_scores || !best_states || !node_features || !nidx || !gidx || !edge_ptr,
                 ECORD_E_NULL);
    ECORD_RET_IF(N <= 0 || G <= 0 || T <= 0 || Fn <= 0 || Fg < 0, ECORD_E_SIZE);
    const int E = edge_ptr[N];   // host buffer: the CSR pointer's last entry is the edge count
    ECORD_RET_IF(E < 0, ECORD_E_SIZE);
    const size_t GT = (size_t)G * T, NT = (size_t)N * T;
    // one device arena, carved 256-byte aligned
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_act = carve(GT * 8), o_sc = carve(GT * 4), o_bs = carve(GT * 4), o_bst = carve(NT * 4),
                 o_nf = carve(NT * Fn * 4), o_gf = carve(GT * (Fg > 0 ? Fg : 1) * 4), o_e = carve((size_t)E * 8 + 8),
                 o_w = carve((size_t)E * 4 + 4), o_p = carve((size_t)(N + 1) * 4);
    char *d = nullptr;
    ECORD_CUDA(cudaMalloc(&d, off));
    cudaStream_t st = nullptr;
    int rc = 0;
#define UP(o, src, bytes) if (!rc && (bytes) && (src)) rc = (int)cudaMemcpyAsync(d + (o), (src), (bytes), cudaMemcpyHostToDevice, st)
    UP(o_act, actions, GT * 8); UP(o_sc, scores, GT * 4); UP(o_bs, best_scores, GT * 4); UP(o_bst, best_states, NT * 4);
    UP(o_nf, node_features, NT * Fn * 4); UP(o_gf, global_features, GT * Fg * 4);
    UP(o_e, edges_by_node, (size_t)E * 8); UP(o_w, edge_attrs, (size_t)E * 4); UP(o_p, edge_ptr, (size_t)(N + 1) * 4);
#undef UP
    if (!rc)
        rc = ecord_step_cy_device((const int64_t *)(d + o_act), (float *)(d + o_sc), (float *)(d + o_bs),
                                  (float *)(d + o_bst), (float *)(d + o_nf),
                                  global_features ? (float *)(d + o_gf) : nullptr, nidx, gidx,
                                  (const int32_t *)(d + o_e), (const float *)(d + o_w), (const int32_t *)(d + o_p), N, G, T,
                                  Fn, Fg, E, st);
#define DOWN(dst, o, bytes) if (!rc && (bytes) && (dst)) rc = (int)cudaMemcpyAsync((dst), d + (o), (bytes), cudaMemcpyDeviceToHost, st)
    DOWN(scores, o_sc, GT * 4); DOWN(best_scores, o_bs, GT * 4); DOWN(best_states, o_bst, NT * 4);
    DOWN(node_features, o_nf, NT * Fn * 4); DOWN(global_features, o_gf, GT * Fg * 4);
#undef DOWN
    if (!rc) rc = (int)cudaStreamSynchronize(st);
    cudaFree(d);
    return rc;
}


// ------------------------------------------------------------------------------------------------
// B3 with a cached arena: ecord_step_cy() pays a cudaMalloc / cudaFree and a full H2D + D2H of every array per call
// (about 2 x the state size over PCIe per step: the price of keeping the reference's host-side Tradjectories untouched).
// A caller that can keep the state on the device between steps creates a context once per Tradjectories object (graph
// uploaded once, arena allocated once), uploads the arrays after __init__, then per step ships only the actions
// (G*T*8 bytes) and optionally reads the scores back; download() materialises the host arrays when they are needed.
// ------------------------------------------------------------------------------------------------
struct ecord_step_cy_ctx {
    int N, G, T, Fn, Fg, E;
    char *arena;
    size_t o_act, o_sc, o_bs, o_bst, o_nf, o_gf, o_e, o_w, o_p, bytes;
    cudaStream_t stream;
};

extern "C" int ecord_step_cy_ctx_create(int32_t N, int32_t G, int32_t T, int32_t Fn, int32_t Fg, const int32_t *edges_by_node,
                                        const float *edge_attrs, const int32_t *edge_ptr, ecord_step_cy_ctx **out) {
    ECORD_RET_IF(!edge_ptr || !out, ECORD_E_NULL);
    ECORD_RET_IF(N <= 0 || G <= 0 || T <= 0 || Fn <= 0 || Fg < 0, ECORD_E_SIZE);
    const int E = edge_ptr[N];
    ECORD_RET_IF(E < 0, ECORD_E_SIZE);
    ECORD_RET_IF(E > 0 && (!edges_by_node || !edge_attrs), ECORD_E_NULL);
    ecord_step_cy_ctx *c = new (std::nothrow) ecord_step_cy_ctx();
    ECORD_RET_IF(!c, (int)cudaErrorMemoryAllocation);
    c->N = N; c->G = G; c->T = T; c->Fn = Fn; c->Fg = Fg; c->E = E; c->arena = nullptr; c->stream = nullptr;
    const size_t GT = (size_t)G * T, NT = (size_t)N * T;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    c->o_act = carve(GT * 8); c->o_sc = carve(GT * 4); c->o_bs = carve(GT * 4); c->o_bst = carve(NT * 4);
    c->o_nf = carve(NT * Fn * 4); c->o_gf = carve(GT * (Fg > 0 ? Fg : 1) * 4); c->o_e = carve((size_t)E * 8 + 8);
    c->o_w = carve((size_t)E * 4 + 4); c->o_p = carve((size_t)(N + 1) * 4);
    c->bytes = off;
    cudaError_t e = cudaMalloc(&c->arena, off);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && E) e = cudaMemcpyAsync(c->arena + c->o_e, edges_by_node, (size_t)E * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && E) e = cudaMemcpyAsync(c->arena + c->o_w, edge_attrs, (size_t)E * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->arena + c->o_p, edge_ptr, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { if (c->arena) cudaFree(c->arena); if (c->stream) cudaStreamDestroy(c->stream); delete c; return (int)e; }
    *out = c;
    return 0;
}

static int ctx_copy(ecord_step_cy_ctx *c, float *scores, float *best_scores, float *best_states, float *node_features,
                    float *global_features, bool up) {
    ECORD_RET_IF(!c, ECORD_E_PLAN);
    const size_t GT = (size_t)c->G * c->T, NT = (size_t)c->N * c->T;
    struct { float *host; size_t off, bytes; } parts[5] = {{scores, c->o_sc, GT * 4}, {best_scores, c->o_bs, GT * 4},
        {best_states, c->o_bst, NT * 4}, {node_features, c->o_nf, NT * c->Fn * 4}, {global_features, c->o_gf, GT * c->Fg * 4}};
    for (auto &p : parts) {
        if (!p.host || !p.bytes) continue;
        if (up) ECORD_CUDA(cudaMemcpyAsync(c->arena + p.off, p.host, p.bytes, cudaMemcpyHostToDevice, c->stream));
        else ECORD_CUDA(cudaMemcpyAsync(p.host, c->arena + p.off, p.bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    ECORD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int ecord_step_cy_ctx_upload(ecord_step_cy_ctx *c, const float *scores, const float *best_scores, const float *best_states,
                                        const float *node_features, const float *global_features) {
    return ctx_copy(c, const_cast<float *>(scores), const_cast<float *>(best_scores), const_cast<float *>(best_states),
                    const_cast<float *>(node_features), const_cast<float *>(global_features), true);
}
extern "C" int ecord_step_cy_ctx_download(ecord_step_cy_ctx *c, float *scores, float *best_scores, float *best_states,
                                          float *node_features, float *global_features) {
    return ctx_copy(c, scores, best_scores, best_states, node_features, global_features, false);
}
extern "C" int ecord_step_cy_ctx_step(ecord_step_cy_ctx *c, const int64_t *actions, const int32_t *nidx, const int32_t *gidx,
                                      float *scores_out) {
    ECORD_RET_IF(!c, ECORD_E_PLAN);
    ECORD_RET_IF(!actions || !nidx || !gidx, ECORD_E_NULL);
    const size_t GT = (size_t)c->G * c->T;
    ECORD_CUDA(cudaMemcpyAsync(c->arena + c->o_act, actions, GT * 8, cudaMemcpyHostToDevice, c->stream));
    char *d = c->arena;
    int rc = ecord_step_cy_device((const int64_t *)(d + c->o_act), (float *)(d + c->o_sc), (float *)(d + c->o_bs), (float *)(d + c->o_bst),
                                  (float *)(d + c->o_nf), c->Fg > 0 ? (float *)(d + c->o_gf) : nullptr, nidx, gidx,
                                  (const int32_t *)(d + c->o_e), (const float *)(d + c->o_w), (const int32_t *)(d + c->o_p), c->N, c->G,
                                  c->T, c->Fn, c->Fg, c->E, c->stream);
    if (rc) return rc;
    if (scores_out) ECORD_CUDA(cudaMemcpyAsync(scores_out, d + c->o_sc, GT * 4, cudaMemcpyDeviceToHost, c->stream));
    ECORD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int ecord_step_cy_ctx_destroy(ecord_step_cy_ctx *c) {
    if (!c) return 0;
    if (c->arena) cudaFree(c->arena);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}
