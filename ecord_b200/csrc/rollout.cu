// rollout.cu -- the S-step rollout loop as CUDA graphs, plus the small C-ABI utilities.
//
// Replaces the Python loop of DQNSolver.rollout (solvers/solver.py:316-362): per step the reference does
// observe() on the host, H2D of [N,T,F], the torch model, D2H of the actions, the Cython step and a list
// append of the scores.  Here one step is {policy_step, q_select, env_step} = 6 kernels (throughput mode: GRU, PROJ, tail, Q, finish, env) enqueued
// back to back; `steps_per_graph` steps are captured once into a CUDA graph and replayed, so the host issues
// one cudaGraphLaunch per k steps and never synchronises inside the rollout.  The step index lives in device
// memory (step_counter): kernels read *step_counter + their unrolled offset, and the last node of each graph
// advances the counter, which keeps the captured graph valid for every replay.
#include "common.cuh"
#include <new>

int launch_policy_step(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                       const int32_t *step_base, int step_off, int gemm_mode, int skip_pre, cudaStream_t st);
int policy_kernels_per_step(int gemm_mode, int skip_pre);
int q_kernels_per_step(int q_mode);
int launch_env_step_fused(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                          const int32_t *step_base, int step_off, float *score_log, int32_t *action_log,
                          unsigned long long *time_log, cudaStream_t st);
int policy_init_attrs();
int check_policy(const ecord_policy *p, int gemm_mode);
int check_q_mode(const ecord_graph *g, const ecord_weights *w, const ecord_policy *p, int q_mode, float tau);
int launch_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                    const int32_t *step_base, int step_off, float tau, uint64_t seed, int compat, const int32_t *forced,
                    int q_mode, float epsilon, cudaStream_t st);
int launch_env_step(const ecord_graph *g, const ecord_env *env, const int32_t *actions, const int32_t *step_base,
                    int step_off, float *score_log, int32_t *action_log, unsigned long long *time_log, cudaStream_t st);

struct ecord_rollout_plan {
    ecord_graph graph;
    ecord_env env;
    ecord_weights weights;
    ecord_policy policy;
    ecord_rollout_desc desc;
    cudaStream_t stream;
    cudaGraphExec_t exec_k, exec_rem;      // k steps; the most recent remainder length (rem_k steps), captured on demand
    int k, rem_k;
    int kernels_per_step;
    int64_t steps_enqueued;
};

namespace {
__global__ void k_advance(int32_t *counter, int by) { *counter += by; }

__global__ void k_score_log_max(const float *__restrict__ log, int S, int rows, float *__restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= rows) return;
    float best = -INFINITY;
    for (int s = 0; s < S; ++s) best = fmaxf(best, log[(long long)s * rows + m]);
    out[m] = best;
}

// log['t_opt'] of the reference (solvers/solver.py:241-242): the time of the LAST step at which a unit's score rose AND equals
// its running best (initial score included) -- a best value that is lost and reached again counts again.  out[m] = that
// step (1-based), 0 if the initial score was never improved on.
__global__ void k_t_opt_steps(const float *__restrict__ log, const float *__restrict__ init_score, int S, int rows,
                              int32_t *__restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= rows) return;
    float prev = init_score[m], best = prev;
    int last = 0;
    for (int s = 0; s < S; ++s) {
        const float sc = log[(long long)s * rows + m];
        best = fmaxf(best, sc);
        if (sc > prev && sc == best) last = s + 1;
        prev = sc;
    }
    out[m] = last;
}

__global__ void k_snapshot_init(ecord_graph g, int T, const int8_t *__restrict__ s0, uint8_t *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.num_nodes * T) return;
    const int v = (int)(idx / T), t = (int)(idx % T);
    const int gi = g.node_graph[v];
    int n;
    out[unit_base(g.graph_ptr, gi, t, T, n) + (v - g.graph_ptr[gi])] = (uint8_t)s0[idx];
}
__global__ void k_snapshot_replay(ecord_graph g, ecord_env e, const int32_t *__restrict__ alog, int S, uint8_t *out) {
    const int T = e.num_tradj;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = g.num_graphs * T;
    if (m >= rows) return;
    const int gi = m / T, t = m - gi * T;
    int n;
    const long long base = unit_base(g.graph_ptr, gi, t, T, n);
    const int upto = min(e.best_step[m], S);
    for (int s = 0; s < upto; ++s) out[base + alog[(long long)s * rows + m]] ^= 1;
}

int capture(ecord_rollout_plan *pl, int nsteps, cudaGraphExec_t *out) {
    cudaGraph_t graph = nullptr;
    ECORD_CUDA(cudaStreamBeginCapture(pl->stream, cudaStreamCaptureModeRelaxed));
    int rc = 0;
    const ecord_rollout_desc &d = pl->desc;
    // ECORD_Q_FAST / ECORD_Q_TC also fuse the next step's msg_in stage into the env-step kernel (u must be primed once with
    // ecord_policy_pre before the first step); ECORD_Q_EXACT keeps the literal kernel sequence.
    const int fused = d.q_mode != ECORD_Q_EXACT;
    for (int i = 0; i < nsteps && !rc; ++i) {
        rc = launch_policy_step(&pl->graph, &pl->env, &pl->weights, &pl->policy, d.step_counter, i, d.gemm_mode, fused,
                                pl->stream);
        if (!rc) rc = launch_q_select(&pl->graph, &pl->env, &pl->weights, &pl->policy, d.step_counter, i, d.tau, d.seed,
                                      d.compat, nullptr, d.q_mode, d.epsilon, pl->stream);
        if (!rc && fused) rc = launch_env_step_fused(&pl->graph, &pl->env, &pl->weights, &pl->policy, d.step_counter, i,
                                                     d.score_log, d.action_log, (unsigned long long *)d.step_time_ns, pl->stream);
        if (!rc && !fused) rc = launch_env_step(&pl->graph, &pl->env, pl->policy.actions, d.step_counter, i, d.score_log,
                                                d.action_log, (unsigned long long *)d.step_time_ns, pl->stream);
    }
    if (!rc) {
        k_advance<<<1, 1, 0, pl->stream>>>(d.step_counter, nsteps);
        rc = (int)cudaPeekAtLastError();
    }
    cudaError_t e = cudaStreamEndCapture(pl->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return (int)e;
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    return (int)e;
}
}  // namespace

extern "C" int ecord_rollout_plan_create(const ecord_rollout_desc *desc, void *stream, ecord_rollout_plan **out) {
    ECORD_RET_IF(!desc || !out, ECORD_E_NULL);
    ECORD_RET_IF(!desc->graph || !desc->env || !desc->weights || !desc->policy || !desc->step_counter, ECORD_E_NULL);
    ECORD_RET_IF(desc->gemm_mode != ECORD_GEMM_FP32 && desc->gemm_mode != ECORD_GEMM_BF16 && desc->gemm_mode != ECORD_GEMM_BF16X3,
                 ECORD_E_UNSUPPORTED);
    ECORD_RET_IF(desc->steps_per_graph <= 0 || desc->steps_per_graph > 256, ECORD_E_SIZE);
    ECORD_RET_IF(desc->tau < 0.0f || !(desc->epsilon >= 0.0f && desc->epsilon <= 1.0f), ECORD_E_SIZE);
    ECORD_RET_IF((desc->score_log || desc->action_log) && desc->max_steps <= 0, ECORD_E_SIZE);
    int rc = check_policy(desc->policy, desc->gemm_mode);
    if (rc) return rc;
    ECORD_RET_IF(desc->policy->rows != desc->graph->num_graphs * desc->env->num_tradj, ECORD_E_SIZE);
    ECORD_RET_IF(desc->tau > 0.0f && desc->q_mode == ECORD_Q_EXACT && !desc->policy->q, ECORD_E_NULL);
    if ((rc = check_q_mode(desc->graph, desc->weights, desc->policy, desc->q_mode, desc->tau))) return rc;
    ECORD_RET_IF(!desc->weights->q_fold, ECORD_E_NULL);
    if (desc->gemm_mode != ECORD_GEMM_FP32) ECORD_RET_IF(!desc->weights->gru_w_bf16 || !desc->weights->p1_w_bf16, ECORD_E_NULL);
    if ((rc = policy_init_attrs())) return rc;
    ecord_rollout_plan *pl = new (std::nothrow) ecord_rollout_plan();
    ECORD_RET_IF(!pl, (int)cudaErrorMemoryAllocation);
    pl->graph = *desc->graph; pl->env = *desc->env; pl->weights = *desc->weights; pl->policy = *desc->policy;
    pl->desc = *desc;
    pl->desc.graph = &pl->graph; pl->desc.env = &pl->env; pl->desc.weights = &pl->weights; pl->desc.policy = &pl->policy;
    pl->stream = as_stream(stream);
    pl->k = desc->steps_per_graph;
    pl->kernels_per_step = policy_kernels_per_step(desc->gemm_mode, desc->q_mode != ECORD_Q_EXACT) +
                           q_kernels_per_step(desc->q_mode) + 1;
    pl->exec_k = pl->exec_rem = nullptr;
    pl->rem_k = 0;
    pl->steps_enqueued = 0;
    rc = capture(pl, pl->k, &pl->exec_k);
    if (rc) { ecord_rollout_plan_destroy(pl); return rc; }
    *out = pl;
    return 0;
}

extern "C" int ecord_rollout_plan_run(ecord_rollout_plan *pl, int32_t num_steps, int64_t *launches_out) {
    ECORD_RET_IF(!pl || !pl->exec_k, ECORD_E_PLAN);
    ECORD_RET_IF(num_steps < 0, ECORD_E_SIZE);
    if ((pl->desc.score_log || pl->desc.action_log) && pl->steps_enqueued + num_steps > pl->desc.max_steps)
        return ECORD_E_SIZE;   // would overrun the logs
    const int nk = num_steps / pl->k, rem = num_steps % pl->k;
    for (int i = 0; i < nk; ++i) ECORD_CUDA(cudaGraphLaunch(pl->exec_k, pl->stream));
    if (rem) {
        // the remainder runs as ONE graph of `rem` steps (captured on first use and kept while the length repeats): `rem`
        // single-step graphs would pay a graph launch and lose the programmatic dependent launches across every step
        if (pl->rem_k != rem) {
            if (pl->exec_rem) { cudaGraphExecDestroy(pl->exec_rem); pl->exec_rem = nullptr; pl->rem_k = 0; }
            int rc = capture(pl, rem, &pl->exec_rem);
            if (rc) return rc;
            pl->rem_k = rem;
        }
        ECORD_CUDA(cudaGraphLaunch(pl->exec_rem, pl->stream));
    }
    pl->steps_enqueued += num_steps;
    if (launches_out) *launches_out = (int64_t)num_steps * pl->kernels_per_step + nk + (rem ? 1 : 0);
    return 0;
}

extern "C" int ecord_rollout_plan_reset(ecord_rollout_plan *pl) {
    ECORD_RET_IF(!pl || !pl->exec_k, ECORD_E_PLAN);
    pl->steps_enqueued = 0;
    return 0;
}

extern "C" int ecord_rollout_plan_destroy(ecord_rollout_plan *pl) {
    if (!pl) return 0;
    if (pl->exec_k) cudaGraphExecDestroy(pl->exec_k);
    if (pl->exec_rem) cudaGraphExecDestroy(pl->exec_rem);
    delete pl;
    return 0;
}

extern "C" int ecord_score_log_max(const float *score_log, int32_t num_steps, int32_t rows, float *out_best, void *stream) {
    ECORD_RET_IF(!score_log || !out_best, ECORD_E_NULL);
    ECORD_RET_IF(num_steps <= 0 || rows <= 0, ECORD_E_SIZE);
    k_score_log_max<<<ceil_div(rows, 128), 128, 0, as_stream(stream)>>>(score_log, num_steps, rows, out_best);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_t_opt_steps(const float *score_log, const float *init_score, int32_t num_steps, int32_t rows,
                                 int32_t *out_steps, void *stream) {
    ECORD_RET_IF(!score_log || !init_score || !out_steps, ECORD_E_NULL);
    ECORD_RET_IF(num_steps < 0 || rows <= 0, ECORD_E_SIZE);
    k_t_opt_steps<<<ceil_div(rows, 128), 128, 0, as_stream(stream)>>>(score_log, init_score, num_steps, rows, out_steps);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_best_state_snapshot(const ecord_graph *g, const ecord_env *env, const int8_t *init_state,
                                         const int32_t *action_log, int32_t num_steps, uint8_t *out, void *stream) {
    ECORD_RET_IF(!g || !env || !init_state || !action_log || !out, ECORD_E_NULL);
    ECORD_RET_IF(num_steps < 0, ECORD_E_SIZE);
    const int T = env->num_tradj;
    const long long NT = (long long)g->num_nodes * T;
    k_snapshot_init<<<ceil_div(NT, 256), 256, 0, as_stream(stream)>>>(*g, T, init_state, out);
    ECORD_LAUNCH_CHECK();
    k_snapshot_replay<<<ceil_div(g->num_graphs * T, 64), 64, 0, as_stream(stream)>>>(*g, *env, action_log, num_steps, out);
    ECORD_LAUNCH_CHECK();
    return 0;
}

extern "C" int ecord_version(void) { return ECORD_ABI_VERSION; }

extern "C" const char *ecord_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case ECORD_E_NULL: return "ECORD_E_NULL: required pointer is NULL";
        case ECORD_E_SIZE: return "ECORD_E_SIZE: size/shape argument out of range";
        case ECORD_E_ALIGN: return "ECORD_E_ALIGN: pointer not aligned as documented";
        case ECORD_E_UNSUPPORTED: return "ECORD_E_UNSUPPORTED: feature table / option not supported";
        case ECORD_E_DEVICE: return "ECORD_E_DEVICE: not an sm_100 device or driver entry point missing";
        case ECORD_E_PLAN: return "ECORD_E_PLAN: invalid plan handle or state";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown ecord error";
    }
}

extern "C" int64_t ecord_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int64_t)sizeof(ecord_graph);
        case 1: return (int64_t)sizeof(ecord_env);
        case 2: return (int64_t)sizeof(ecord_weights);
        case 3: return (int64_t)sizeof(ecord_policy);
        case 4: return (int64_t)sizeof(ecord_rollout_desc);
        default: return -1;
    }
}
