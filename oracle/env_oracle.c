/*
 * oracle/env_oracle.c -- CPU restatement of the reference's native environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  Not linked into, imported by, or called from the product
 * (ecord_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the timed CPU baseline.
 *
 * Follows /root/reference/ecord/environment/_tradjectory_step_cy.pyx:
 *   - oracle_step_cy()  <->  step_cy() :152-171 (dense `static += 1`, `steps += 1` prologue)
 *                             + _step() :64-150 (the nine ordered per-(graph,trajectory) updates)
 * and /root/reference/ecord/environment/tradjectories.py:
 *   - oracle_scores_and_peeks() <-> Tradjectories._calc_scores_and_peeks() :130-148
 *
 * Pinned: tests/test_oracle.py asserts bit-equality against the compiled reference
 * (oracle/_ref, built by oracle/build_ref.py from the reference .pyx where it lies) and
 * against the committed golden traces in tests/golden/.
 *
 * Layouts are the reference's: node_features f32[N][T][Fn] (C-contiguous),
 * global_features f32[G][T][Fg], actions i64[G][T] holding GLOBAL (flat) vertex ids,
 * edges_by_node i32[2][E] (row 0 = source, row 1 = neighbour), CSR pointer i32[N+1].
 * All arithmetic is single precision, in the reference's order of operations.
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#define NF(v, t, d) node_features[((size_t)(v) * T + (size_t)(t)) * Fn + (size_t)(d)]
#define GF(g, t, d) global_features[((size_t)(g) * T + (size_t)(t)) * Fg + (size_t)(d)]

void oracle_step_cy(const int64_t *actions, float *scores, float *best_scores, float *best_states,
                    float *node_features, float *global_features,
                    const int32_t *node_idx, /* [state, peek, best_state, static, degree] */
                    const int32_t *glob_idx, /* [score_from_best, steps_since_best, num_greedy, steps, max_peek, mean_peek] */
                    const int32_t *edges_by_node, const float *edge_w, const int32_t *edge_ptr,
                    int32_t N, int32_t G, int32_t T, int32_t Fn, int32_t Fg, int32_t E)
{
    const int d_state = node_idx[0], d_peek = node_idx[1], d_best = node_idx[2], d_static = node_idx[3];
    const int g_sfb = glob_idx[0], g_ssb = glob_idx[1], g_ngr = glob_idx[2], g_steps = glob_idx[3];
    const int32_t *e_src = edges_by_node, *e_dst = edges_by_node + E;

    /* step_cy() prologue, pyx:160-164 */
    if (d_static >= 0)
        for (size_t i = 0; i < (size_t)N * T; ++i) node_features[i * Fn + d_static] += 1.0f;
    if (g_steps >= 0)
        for (size_t i = 0; i < (size_t)G * T; ++i) global_features[i * Fg + g_steps] += 1.0f;

    for (int g = 0; g < G; ++g) {
        for (int t = 0; t < T; ++t) {
            const int64_t a = actions[(size_t)g * T + t];
            float *sc = &scores[(size_t)g * T + t], *bs = &best_scores[(size_t)g * T + t];

            /* 1. score, pyx:87 */
            *sc = *sc + NF(a, t, d_peek);
            /* 2. best score, pyx:90-93 */
            int new_best = 0;
            if (*sc > *bs) { *bs = *sc; new_best = 1; }
            /* 3. score from best, pyx:96-97 */
            if (g_sfb >= 0) GF(g, t, g_sfb) = *bs - *sc;
            /* 4. steps since best, pyx:100-105 */
            if (g_ssb >= 0) {
                if (new_best) GF(g, t, g_ssb) = 0.0f;
                else GF(g, t, g_ssb) = GF(g, t, g_ssb) + 1.0f;
            }
            /* 5(i). own peek flips sign, pyx:109-116 */
            NF(a, t, d_peek) = -1.0f * NF(a, t, d_peek);
            if (g_ngr >= 0) {
                if (NF(a, t, d_peek) > 0) GF(g, t, g_ngr) = GF(g, t, g_ngr) + 1.0f;
                if (NF(a, t, d_peek) < 0) GF(g, t, g_ngr) = GF(g, t, g_ngr) - 1.0f;
            }
            /* 5(ii)+6. neighbours, pyx:119-137 (uses the PRE-flip state of a) */
            for (int k = edge_ptr[a]; k < edge_ptr[a + 1]; ++k) {
                const int32_t u = e_src[k], v = e_dst[k];
                const float sw = (2.0f * NF(u, t, d_state) - 1.0f) * (2.0f * NF(v, t, d_state) - 1.0f) * edge_w[k];
                const float old_peek = NF(v, t, d_peek);
                const float new_peek = old_peek - 2.0f * sw;
                if (g_ngr >= 0) {
                    int dg = 0;
                    if (new_peek > 0 && old_peek <= 0) dg = 1;
                    else if (new_peek <= 0 && old_peek > 0) dg = -1;
                    GF(g, t, g_ngr) = GF(g, t, g_ngr) + (float)dg;
                }
                NF(v, t, d_peek) = new_peek;
            }
            /* 7. flip, pyx:140 */
            NF(a, t, d_state) = fmodf(NF(a, t, d_state) + 1.0f, 2.0f);
            /* 8. state at best: only the flipped vertex is recorded, pyx:142-146 */
            if (new_best) {
                best_states[(size_t)a * T + t] = NF(a, t, d_state);
                if (d_best > 0) NF(a, t, d_best) = best_states[(size_t)a * T + t];
            }
            /* 9. steps static, pyx:149-150 */
            if (d_static >= 0) NF(a, t, d_static) = 0.0f;
        }
    }
}

/*
 * scores[g][t] = 0.5 * sum_{directed edges e of g, in order} w_e * ((s_src + s_dst) mod 2)
 * peeks[v][t]  =       sum_{edges e out of v, in CSR order}  w_e * (2 s_v - 1)(2 s_nbr - 1)
 * (tradjectories.py:135-146; torch_scatter.segment_csr sums each segment sequentially in fp32).
 * states i8[N][T] in {0,1}; graph_edge_ptr i64[G+1] = offsets into the directed edge list.
 */
void oracle_scores_and_peeks(const int8_t *states, const int32_t *edges_by_node, const float *edge_w,
                             const int32_t *edge_ptr, const int64_t *graph_edge_ptr,
                             float *scores, float *peeks, int32_t N, int32_t G, int32_t T, int32_t E)
{
    const int32_t *e_src = edges_by_node, *e_dst = edges_by_node + E;
    for (int g = 0; g < G; ++g)
        for (int t = 0; t < T; ++t) {
            float acc = 0.0f;
            for (int64_t k = graph_edge_ptr[g]; k < graph_edge_ptr[g + 1]; ++k) {
                const int s = states[(size_t)e_src[k] * T + t] + states[(size_t)e_dst[k] * T + t];
                acc += edge_w[k] * (float)(s % 2);
            }
            scores[(size_t)g * T + t] = 0.5f * acc;
        }
    for (int v = 0; v < N; ++v)
        for (int t = 0; t < T; ++t) {
            float acc = 0.0f;
            for (int k = edge_ptr[v]; k < edge_ptr[v + 1]; ++k) {
                const float si = (float)(2 * states[(size_t)e_src[k] * T + t] - 1);
                const float sj = (float)(2 * states[(size_t)e_dst[k] * T + t] - 1);
                acc += edge_w[k] * si * sj;
            }
            peeks[(size_t)v * T + t] = acc;
        }
}
