/*
 * ecord_b200.h -- C ABI of libecord_b200.so, the B200 (sm_100a) rollout engine for ECORD.
 *
 * This is the drop-in boundary for the reference's test-time rollout hot path
 * (SURVEY.md section 8).  Plain C: pointers, sizes and PODs only -- no torch / C++ types.
 * Reference citations are relative to the upstream tree (tomdbar/ecord).
 *
 * Conventions
 *   - Every entry point returns int: 0 = ok, <0 = argument error (ECORD_E_*), >0 = cudaError_t.
 *     Nothing throws, exits or prints.  ecord_error_string() names a code.
 *   - Unless a function says "host", every pointer is caller-allocated DEVICE memory; the library
 *     allocates nothing on the data path (plans own only their CUDA graphs).  Kernels are enqueued
 *     on the `stream` argument (a cudaStream_t passed as void*) and do not synchronise.
 *   - Re-entrant: no global mutable state; one host thread per GPU is the intended use.
 *   - Vertex ids: "global" = index into the flat batch [0,N); "local" = index inside its graph.
 *
 * Device data layout (DESIGN.md "Data layout in HBM")
 *   The reference keeps AoS node_features f32[N][T][F] on the host (tradjectories.py:150-166).
 *   Here the per-(vertex,trajectory) state is SoA and TRAJECTORY-MAJOR per graph, so that one
 *   (graph, trajectory) unit is one contiguous row of n_g elements:
 *        plane_offset(g, t, v) = graph_ptr[g] * T + t * n_g + v          (v local)
 *     peek  f32 : cut gain if v flips (reference NodeObservation.PEEK, un-normalised)
 *     sp    i32 : (last_flip_step << 1) | state      static = steps_done - last_flip_step
 *     best  u8  : reference "best_states" (only the flipped vertex is recorded on a new best,
 *                 _tradjectory_step_cy.pyx:142-146)
 *   Per-(graph,trajectory) scalars are indexed m = g * T + t.
 */
#ifndef ECORD_B200_H
#define ECORD_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECORD_ABI_VERSION 3

/* error codes (<0) */
#define ECORD_E_NULL      (-1)  /* required pointer is NULL */
#define ECORD_E_SIZE      (-2)  /* size/shape argument out of range */
#define ECORD_E_ALIGN     (-3)  /* pointer not aligned as documented */
#define ECORD_E_UNSUPPORTED (-4)/* feature table / option not supported by this kernel */
#define ECORD_E_DEVICE    (-5)  /* device is not sm_100 / required driver entry point missing */
#define ECORD_E_PLAN      (-6)  /* invalid plan handle or plan state */

/* network dimensions (canonical ECORD net, experiments/train.py:55-96) */
#define ECORD_DIM_GNN   16
#define ECORD_DIM_NODE  32
#define ECORD_DIM_MSG   64
#define ECORD_DIM_H     1024
#define ECORD_DIM_P1    512
#define ECORD_DIM_Q     64
#define ECORD_STATIC_CLIP 40   /* Tradjectories._clip_features_max, tradjectories.py:16 */

/* GRU / projection arithmetic */
#define ECORD_GEMM_FP32 0   /* CUDA-core fp32 FMA (parity mode; 1e-4 rel vs reference)            */
#define ECORD_GEMM_BF16 1   /* tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM (2e-2 rel) */
#define ECORD_GEMM_BF16X3 2 /* the same tcgen05 kernels with every operand split into a bf16 hi part and a bf16 lo remainder
                               (x ~ hi + lo to 2^-17) and the products hi.hi + lo.hi + hi.lo accumulated in fp32, exact
                               sigmoid / tanh in the GRU cell: fp32-grade recurrent state on the tensor cores (the engine's
                               default; rows_pad must be a multiple of 256, ecord_policy.split = 1) */

/* ---- graph batch: replaces GraphBatcher/Batch.tg (environment/data.py:53-83) + the CSR views
 *      Tradjectories builds (tradjectories.py:38-60).  All arrays read-only. ---- */
typedef struct ecord_graph {
    int32_t num_graphs;            /* G */
    int32_t num_nodes;             /* N = sum n_g */
    int32_t num_edges;             /* E = directed edge count (2 x undirected) */
    int32_t n_max;                 /* max n_g */
    const int32_t *graph_ptr;      /* [G+1] vertex offsets (Batch.ptr) */
    const int32_t *node_graph;     /* [N]   graph id of each vertex (Batch.batch) */
    const int32_t *row_ptr;        /* [N+1] CSR by source (edge_by_node_ptr) */
    const int32_t *col;            /* [E]   neighbour, global id, reference edge order (edges_by_node[1]) */
    const float   *w;              /* [E]   edge weight (edge_attrs_by_node) */
    const int64_t *graph_edge_ptr; /* [G+1] offsets of each graph's directed edges in [0,E] */
    const int32_t *in_ptr;         /* [N+1] incoming-edge lists, in global edge order (for scatter_mean) */
    const int32_t *in_src;         /* [E]   source (global id) of each incoming edge */
    const float   *in_w;           /* [E]   its weight */
    /* vertex tiles of ECORD_QTILE vertices for the fast Q kernel: tile i covers vertices
       [qtile_v0[i], qtile_v0[i] + ECORD_QTILE) (local ids, clipped to n_g) of graph qtile_graph[i] */
    int32_t num_qtiles;
    int32_t int_weights;           /* 1: every edge weight is an integer and sum |w| < 2^24 per graph (sums exact in any order) */
    const int32_t *qtile_graph;    /* [num_qtiles] */
    const int32_t *qtile_v0;       /* [num_qtiles] */
    const float   *graph_wsum;     /* [G] sum of each graph's directed edge weights, or NULL (ecord_env_init then walks the edges) */
} ecord_graph;
#define ECORD_QTILE 128

/* ---- environment state: replaces Tradjectories' arrays (tradjectories.py:78-116) ---- */
typedef struct ecord_env {
    int32_t  num_tradj;            /* T */
    int32_t  true_snapshot;        /* 0: best_state follows the reference (only the flipped vertex is recorded on a new best,
                                      quirk D); 1: the whole spin row is copied on a new best, so cut(best_state) == best_score */
    float   *peek;                 /* [N*T] plane */
    int32_t *sp;                   /* [N*T] plane */
    uint8_t *best_state;           /* [N*T] plane */
    float   *score;                /* [G*T] */
    float   *best_score;           /* [G*T] */
    int32_t *best_step;            /* [G*T] step count (1-based) at which best_score was reached; 0 = initial */
} ecord_env;

/* ---- weights: fp32 device tensors in PyTorch layout [out][in], named by the reference's
 *      state_dict keys.  ---- */
typedef struct ecord_weights {
    /* gnn (models/gnn_embedder.py:105-180) */
    const float *gnn_embed_w, *gnn_embed_b;        /* [16,1],[16] */
    const float *gnn_snd_w, *gnn_snd_b;            /* [16,16],[16] */
    const float *gnn_rec_w, *gnn_rec_b;            /* [32,32],[32] */
    const float *gnn_wih, *gnn_whh, *gnn_bih, *gnn_bhh; /* [48,32],[48,16],[48],[48] */
    /* gnn2rnn (models/gnn_to_rnn.py:6-18), node_encoder (configs.py:174-192) */
    const float *g2r_w, *g2r_b;                    /* [16,16],[16] */
    const float *enc_w, *enc_b;                    /* [16,3],[16] */
    /* rnn (models/rnn_decoder.py:22-136,313-381) */
    const float *msg_w, *msg_b;                    /* [64,34],[64] */
    const float *gru_wih, *gru_whh, *gru_bih, *gru_bhh; /* [3072,64],[3072,1024],[3072],[3072] */
    const float *p1_w, *p1_b, *ln1_w, *ln1_b;      /* [512,1024],[512],[512],[512] */
    const float *p2_w, *p2_b, *ln2_w, *ln2_b;      /* [32,512],[32],[32],[32] */
    const float *v1_w, *v1_b, *v2_w, *v2_b;        /* [32,32],[32],[1,32],[1] */
    const float *q1_w, *q1_b, *qln_w, *qln_b, *q2_w, *q2_b; /* [64,64],[64],[64],[64],[1,64],[1] */
    /* derived by ecord_pack_weights() into caller-allocated buffers */
    void  *gru_w_bf16;   /* bf16, ECORD_GRU_PACK_ELEMS elements: Wh pack [64][48][1024] (rows r16|z16|hn16 of each 16-unit
                            sub-block: a GRU tile of 16, 32 or 64 units is a contiguous row range) as a hi plane (bf16(w)) and a
                            lo plane (bf16(w - hi)); then the Wu_main pack [64][48][64] (rows r16|z16|0), hi plane | lo plane;
                            then Wu_in [1024][64] (W_in), hi plane | lo plane */
    void  *p1_w_bf16;    /* bf16, ECORD_P1_PACK_ELEMS elements: proj_rnn_to_gnn.1 weight [512][1024] hi plane | lo plane, followed
                            by proj_rnn_to_gnn.4 weight [32][512] hi plane | lo plane */
    float *q_tc;         /* f32 [ECORD_QTC_FLOATS] operands of the tensor-core Q kernel (ECORD_Q_TC), or NULL:
                              [0,1280)     reserved
                              [1280,2304)  Wg [64][16] = Wq[:,32:48] centred over the 64 channels (fp32)
                              [2304,2560)  [64][4]: Cc[c][0..2] (C = Wq[:,48:64] W_enc, centred), bconst[c] = centred Wq[:,48:64] b_enc
                              [2560,2816)  reserved
                              [2816,4736)  shared-memory image of the kernel's B operand: 3 swizzled [80][16] fp16 blocks
                                           (W1 hi | W1 lo | W2), 7680 bytes
                              [4736,10148) shared-memory image of the policy tail's constants (one bulk copy per CTA):
                                           LN512 gamma[512] | beta[512] | val_out.1 weight [32][32] | Wq[:, :32] as [64][32] |
                                           Wg [64][16] | Cc,bconst [64][4] | column sums of Wq[:, :32] [32], sum(b_q), 3 pad
                            filled by ecord_pack_weights() */
    const float *q_fold; /* HOST f32 [ECORD_QFOLD_FLOATS] from ecord_fold_q_host():
                              [0,384)   C0|C1|C2 (C = Wq[:,48:64] @ W_enc, 64x3) | LN gamma | LN beta | w_o
                              [384,576) the same C with each column centred over the 64 channels (fast Q kernel)
                              [576,579) LC_k = sum_c w_o[c] gamma[c] Ccentred[c][k];  [579] sum_c w_o[c] beta[c]
                              [580,586) Ccentred^T Ccentred (3x3 symmetric: 00 01 02 11 12 22); [586,588) unused
                            Passed to the Q kernels by value (constant bank), hence a host array. */
#define ECORD_QFOLD_FLOATS 588
#define ECORD_QTC_FLOATS 10160
#define ECORD_GRU_PACK_ELEMS (2 * (3072 * 1024 + 4096 * 64))
#define ECORD_P1_PACK_ELEMS (2 * (512 * 1024 + 32 * 512))
} ecord_weights;

/* ---- per-rollout policy state and scratch (all caller-allocated) ---- */
typedef struct ecord_policy {
    int32_t  rows;        /* M = G*T */
    int32_t  rows_pad;    /* M rounded up to a multiple of 128 (256 selects the CTA-pair tcgen05 GEMM kernels);
                             every [M][*] buffer below has rows_pad rows */
    float   *hidden;      /* [2][rows_pad][1024] double-buffered GRU state (fp32) */
    void    *hidden_bf16; /* [2][rows_pad][1024] bf16 mirror (tcgen05 A operand); split: [2][2][rows_pad][1024] = per state buffer
                             a hi plane and a lo plane (bf16(h - hi)) */
    void    *th_bf16;     /* [rows_pad][1024] bf16 tanh(h) (projection A operand); split: [2][rows_pad][1024] hi | lo */
    void    *u_bf16;      /* [rows_pad][64] bf16 GRU input; split: [2][rows_pad][64] hi | lo */
    float   *u;           /* [rows_pad][64] GRU input  LReLU(msg_in([last_sel; glob])) */
    float   *gates;       /* [rows_pad][6144] fp32-mode scratch: gi | gh */
    float   *th;          /* [rows_pad][1024] fp32-mode tanh(h) */
    float   *y1;          /* [rows_pad][512] projection pre-LayerNorm */
    float   *P;           /* [rows_pad][32]  proj_rnn_to_gnn output */
    float   *A;           /* [rows_pad][64]  Wq[:, :32] @ P + b_q (per-(g,t) part of mlp_out[0]) */
    float   *v;           /* [rows_pad]      state value + mlp_out bias */
    float   *last_sel;    /* [rows_pad][32]  embedding of the last selected vertex */
    const float *node_emb;/* [N][16] gnn2rnn(gnn(x)) (actors/_base.py:218-224) */
    const float *node_B;  /* [N][64] Wq[:,32:48] @ node_emb + Wq[:,48:64] @ b_enc */
    int32_t *actions;     /* [rows] local vertex id chosen at the current step */
    float   *q;           /* [N*T] plane, optional Q-value dump (may be NULL when tau == 0) */
    /* fast Q kernel (q_mode = ECORD_Q_FAST): channel-centred operands and the linear part of the head */
    float   *Ac;          /* [rows_pad][64]  A - mean_c(A) */
    float   *LA;          /* [rows_pad]      sum_c w_o[c] gamma[c] Ac[c] */
    const float *node_Bc; /* [N][64]         node_B - mean_c(node_B) */
    const float *node_LB; /* [N]             sum_c w_o[c] gamma[c] node_Bc[c] */
    unsigned long long *keys; /* [rows]      packed (ordered q << 32 | ~vertex) running arg-max; zero between steps */
    /* tensor-core Q kernel (q_mode = ECORD_Q_TC); with a' = Ac + bconst:
         qa[m] = 92 32-bit words: [0,64) fp16 pairs (hi, lo) of gamma_c a'_c | [64,72) Wg^T a' hi (16 fp16) | [72,80) lo |
                 [80,88) row 64 of W2: fp16 [ca0h ca1h ca2h ca1h 0 0 ca2h 0 | ca0l ca1l ca2l 0 0 0 0 0], ca = Cc^T a' |
                 f32 |a'|^2 | f32 LA | f32 v | 0                                            written by the policy tail
         node_qv[n] = [ |b'|^2 | Cc^T b' (3) | LB | 0 0 0 ],  b' = Wg node_emb[n]             written by ecord_gnn_embed */
    float   *qa;          /* [rows_pad][ECORD_QA_STRIDE] */
    float   *node_qv;     /* [N][8] */
    int32_t  split;       /* 1: the bf16 buffers above carry hi and lo planes (ECORD_GEMM_BF16X3); 0 otherwise */
    int32_t  _pad;
} ecord_policy;

/* Q-head evaluation */
#define ECORD_Q_EXACT 0   /* one CTA per (g,t); LayerNorm/LeakyReLU/dot evaluated literally (parity mode)             */
#define ECORD_QA_STRIDE 92
#define ECORD_Q_FAST  1   /* vertex-tiled, per-vertex operands register-resident across trajectories; centred LN and
                             LeakyReLU(y) = 0.505 y + 0.495 |y| with the linear half folded out (re-associated fp32) */
#define ECORD_Q_TC    2   /* the same folded head with the 64-channel pre-activation of every (vertex, trajectory) row produced
                             by tcgen05.mma kind::tf32 (operands built in shared memory, accumulators in TMEM) and only the
                             |.|-sum left on the CUDA cores.  Every operand is split into fp16 hi + lo parts (3-term products,
                             kind::f16), so the result is fp32-grade (~1e-6 relative) as long as |operands| < 65504.
                             tau == 0 only. */

/* ------------------------------------------------------------------------------------------ */
int         ecord_version(void);
const char *ecord_error_string(int code);
/* sizeof() of the structs above as compiled, so bindings can assert their mirror is in sync:
 * which = 0 graph, 1 env, 2 weights, 3 policy, 4 rollout_desc */
int64_t     ecord_abi_sizeof(int which);
/* Diagnostics (no GPU needed): the tile schedule of the tcgen05 GRU kernel (the a6 update, models/rnn_decoder.py:369-381) for a
 * batch of rows_pad rows (a multiple of 256): out[3 i .. 3 i + 2] = (first row, first hidden unit, hidden units) of tile i, cluster
 * by cluster; *num_clusters (may be NULL) = CTA pairs of the launch.  Returns the number of tiles (which may exceed max_tiles;
 * only the first max_tiles are written), or ECORD_E_SIZE (negative).  tests/ check that the tiles cover [rows_pad] x [1024] exactly once. */
int         ecord_gru_tile_schedule(int32_t rows_pad, int32_t *out, int32_t max_tiles, int32_t *num_clusters);

/* ---- B3: the reference's native seam.  Same arguments, meaning and in-place semantics as
 *      step_cy() (_tradjectory_step_cy.pyx:152-171) called from Tradjectories.step()
 *      (tradjectories.py:387-405); the two trailing sizes are explicit because C has no shapes.
 *      ecord_step_cy takes HOST buffers (numpy arrays as the reference holds them), copies them to
 *      the current device, runs the sm_100a kernel and copies the mutated arrays back.
 *      ecord_step_cy_device is the same on DEVICE buffers in the reference layout, asynchronous. ---- */
int ecord_step_cy(const int64_t *actions,       /* [G][T] global vertex ids */
                  float *scores, float *best_scores,          /* [G][T] */
                  float *best_states,                         /* [N][T] */
                  float *node_features,                       /* [N][T][num_node_features] */
                  float *global_features,                     /* [G][T][num_glob_features] */
                  const int32_t *node_feature_idxs,           /* [5] state,peek,best_state,static,degree (-1 = absent) */
                  const int32_t *glob_feature_idxs,           /* [6] score_from_best,steps_since_best,num_greedy,steps,max_peek,mean_peek */
                  const int32_t *edges_by_node,               /* [2][E] */
                  const float *edge_attrs_by_node,            /* [E] */
                  const int32_t *edge_by_node_ptr,            /* [N+1] */
                  int32_t num_nodes, int32_t num_graphs, int32_t num_tradj,
                  int32_t num_node_features, int32_t num_glob_features);
int ecord_step_cy_device(const int64_t *actions, float *scores, float *best_scores, float *best_states,
                         float *node_features, float *global_features,
                         const int32_t *node_feature_idxs /* host [5] */, const int32_t *glob_feature_idxs /* host [6] */,
                         const int32_t *edges_by_node, const float *edge_attrs_by_node,
                         const int32_t *edge_by_node_ptr,
                         int32_t num_nodes, int32_t num_graphs, int32_t num_tradj,
                         int32_t num_node_features, int32_t num_glob_features, int32_t num_edges, void *stream);

/* B3 with a cached arena: ecord_step_cy() allocates, uploads every array, steps, downloads every array and frees on EVERY
 * call (~2 x the state size over PCIe per step).  A context keeps the graph and the state on the device: create once per
 * Tradjectories object (the CSR arrays are host buffers, uploaded once), upload the arrays after __init__ (any pointer may be
 * NULL = skip), then per step only the actions [G][T] i64 go up and, optionally, the scores [G][T] come back; download()
 * materialises the host arrays (any pointer may be NULL).  Same kernel, same bit-exact results as ecord_step_cy. */
typedef struct ecord_step_cy_ctx ecord_step_cy_ctx;   /* opaque */
int ecord_step_cy_ctx_create(int32_t num_nodes, int32_t num_graphs, int32_t num_tradj, int32_t num_node_features,
                             int32_t num_glob_features, const int32_t *edges_by_node, const float *edge_attrs_by_node,
                             const int32_t *edge_by_node_ptr, ecord_step_cy_ctx **out);
int ecord_step_cy_ctx_upload(ecord_step_cy_ctx *ctx, const float *scores, const float *best_scores, const float *best_states,
                             const float *node_features, const float *global_features);
int ecord_step_cy_ctx_step(ecord_step_cy_ctx *ctx, const int64_t *actions, const int32_t *node_feature_idxs,
                           const int32_t *glob_feature_idxs, float *scores_out /* host [G][T] or NULL */);
int ecord_step_cy_ctx_download(ecord_step_cy_ctx *ctx, float *scores, float *best_scores, float *best_states,
                               float *node_features, float *global_features);
int ecord_step_cy_ctx_destroy(ecord_step_cy_ctx *ctx);

/* ---- a2: initial scores / peeks from spin states; replaces Tradjectories.__init__ +
 *      _calc_scores_and_peeks (tradjectories.py:62-116,130-148).
 *      init_state: i8 [N][T] (reference layout), values in {0,1}. ---- */
int ecord_env_init(const ecord_graph *g, const ecord_env *env, const int8_t *init_state, void *stream);

/* ---- a3: one environment step on the SoA state; actions are LOCAL ids [G*T] (the engine's own
 *      layout; Tradjectories.step adds the batch offset, tradjectories.py:389-390).
 *      step_no = number of steps completed BEFORE this one.  score_log (optional) receives the new
 *      score [G*T] (solver.py:238 log['scores']). ---- */
int ecord_env_step(const ecord_graph *g, const ecord_env *env, const int32_t *actions, int32_t step_no,
                   float *score_log, void *stream);
/* the same step (actions = policy->actions) fused with the NEXT step's msg_in stage, which needs exactly the
 * quantities the flip just produced (best - score) and the cached embedding: saves a launch per step. */
int ecord_env_step_fused(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                         int32_t step_no, float *score_log, void *stream);

/* ---- (8f rank 1) greedy local search on the peeks: greedy_solve (solvers/solver.py:38-95), used by
 *      DQNSolver.rollout(pre_solve_with_greedy / post_solve_with_greedy_from_best) (solver.py:294-295,367-371).
 *      ecord_greedy_select draws, per (g,t), the reference's sample argmin_n -log(U_n) / exp((peek_n - max peek) / max(tau,1e-9))
 *      (solver.py:64, scatter_softmax_sample :30-36; Philox stream (seed, step_no)): uniform over the vertices of maximal
 *      peek at tau = 0.  actions: LOCAL ids [G*T]; at_local_min (optional) [G*T]: 1 where max peek <= 0 (solver.py:45,75-77).
 *      The caller applies the actions with ecord_env_step and loops until every unit has been at a local optimum.
 *      ecord_env_reset_state rebuilds peeks / scores / static counters from spins (i8 [N][T]) or, with NULL, from the
 *      best-state plane (init_from_best, solver.py:39-43) and leaves best_score / best_state / best_step untouched.
 *      ecord_env_rebase_clock moves the environment clock back by `by` steps (last_flip_step -= by, best_step -= by): after a
 *      pre-solve the rollout restarts at step 0 while the static-steps feature keeps counting from the greedy flips. ---- */
int ecord_greedy_select(const ecord_graph *g, const ecord_env *env, float tau, uint64_t seed, int32_t step_no,
                        int32_t *actions, int32_t *at_local_min, void *stream);
int ecord_env_reset_state(const ecord_graph *g, const ecord_env *env, const int8_t *init_state /* or NULL */,
                          int32_t steps_done, void *stream);
int ecord_env_rebase_clock(const ecord_graph *g, const ecord_env *env, int32_t by, void *stream);

/* ---- a4 + parity view: materialise the reference's arrays from the SoA state.
 *      node_features f32 [N][T][3] = (state, peek, static) un-normalised; with normalise != 0 the
 *      canonical observation instead (peek/n_g, min(static,40)/40; tradjectories.py:336-353);
 *      glob_features f32 [G][T][2] = (best-score [/n_g], 0) (quirk C); best_states f32 [N][T].
 *      Any output pointer may be NULL. ---- */
int ecord_env_export(const ecord_graph *g, const ecord_env *env, int32_t steps_done, int32_t normalise,
                     float *node_features, float *glob_features, float *best_states, void *stream);

/* ---- a5: t=0 GNN embedding; replaces _ActorBase.initialise_embeddings -> GNNEmbedder.forward +
 *      GNN2RNN (actors/_base.py:185-237, gnn_embedder.py:133-149,182-199, gnn_to_rnn.py:16-18),
 *      one pass (the reference's 10 extra timing passes, gnn_embedder.py:201-217, are dropped).
 *      LayerNorm statistics are batch-wide (quirk B).  gnn_out, node_emb: f32 [N][16];
 *      node_B: f32 [N][64] (see ecord_policy; may be NULL when only the tensor-core Q kernel, which works from node_emb and
 *      node_qv, will run).  workspace: ecord_gnn_workspace_bytes(N) bytes. ---- */
size_t ecord_gnn_workspace_bytes(int32_t num_nodes);
int ecord_gnn_embed(const ecord_graph *g, const ecord_weights *w, float *gnn_out, float *node_emb, float *node_B,
                    float *node_Bc /* [N][64] or NULL */, float *node_LB /* [N] or NULL */,
                    float *node_qv /* [N][8] or NULL; needs w->q_tc */, void *workspace, void *stream);

/* ---- weight packing for the tensor-core path (device) and the folded Q-head constants (host arrays in/out) ---- */
int ecord_pack_weights(const ecord_weights *w, void *stream);
int ecord_fold_q_host(const float *q1_w /*[64,64]*/, const float *enc_w /*[16,3]*/, const float *qln_w, const float *qln_b,
                      const float *q2_w /*[1,64]*/, float *out /*[ECORD_QFOLD_FLOATS]*/);

/* ---- a6 + a8 (per-(graph,trajectory) part): msg_in -> GRUCell(64->1024) -> proj_rnn_to_gnn -> val_out;
 *      replaces _update_internal_state + GRUDecoder.update_embeddings + the P/state-value half of
 *      _forward_action_value (actors/_base.py:247-269, rnn_decoder.py:369-381,285-311).
 *      cur = index (0/1) of the hidden buffer holding h_{t-1}; h_t is written to buffer 1-cur. ---- */
int ecord_policy_step(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                      int32_t cur, int32_t gemm_mode, int32_t skip_pre, void *stream);
/* the msg_in stage alone: u = LReLU(msg_in([last_sel; glob])).  With skip_pre != 0 ecord_policy_step expects u to be
 * current: ecord_env_step_fused leaves it so (it recomputes u right after the flip), and this entry primes it at t=0. */
int ecord_policy_pre(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p, void *stream);

/* ---- a7 + a8 (per-vertex part) + a9: Q-values over every vertex of every (graph,trajectory),
 *      greedy (tau == 0: argmax, lowest index on ties) or soft (tau > 0: exponential-race sampling,
 *      actors/ecord.py:12-19, Philox stream seeded by (seed, step_no)) selection, and the cache of the
 *      chosen vertex's embedding for the next step.  compat != 0 reproduces the reference's gather
 *      with the LOCAL index into the FLAT array (quirk A: ecord.py:167,183 -> _base.py:305-308).
 *      forced_actions (optional, [G*T] local ids) overrides the selection (teacher forcing).
 *      epsilon in [0,1]: epsilon-greedy (actors/ecord.py:171-180) -- with that probability a unit's action is replaced
 *      by floor(U * n_g), U uniform (Philox stream seeded by (seed, unit, step_no)). ---- */
int ecord_q_select(const ecord_graph *g, const ecord_env *env, const ecord_weights *w, const ecord_policy *p,
                   int32_t steps_done, float tau, uint64_t seed, int32_t compat, const int32_t *forced_actions,
                   int32_t q_mode /* ECORD_Q_*; tau > 0: ECORD_Q_EXACT (exponential race, two passes) or ECORD_Q_TC
                                      (Gumbel-max on the same Philox draws, one pass) */, float epsilon, void *stream);

/* ---- network initialisation (rollout flag use_network_initialisation; solvers/solver.py:256-277,
 *      actors/_base.py:238-245, models/node_classifier.py:8-30): init_state_out i8 [N][T] ~ Bernoulli(sigmoid(
 *      head_w[head] . gnn_out[v] + head_b[head])), gnn_out = the GNN embedding BEFORE gnn2rnn (ecord_gnn_embed).
 *      head_w f32 [num_heads][16], head_b f32 [num_heads] (device).  With num_heads > 1 every spin is re-drawn
 *      uniformly with probability perturb_epsilon (solver.py:268-273); with one head the argument is ignored, as in
 *      the reference.  The draw of (vertex v, trajectory t) depends on (seed, v, tradj_offset + t) only, so a
 *      trajectory shard reproduces its columns of the unsharded draw.  The result feeds ecord_env_init. ---- */
int ecord_node_init_sample(const ecord_graph *g, int32_t num_tradj, int32_t tradj_offset, const float *gnn_out, const float *head_w,
                           const float *head_b, int32_t num_heads, int32_t head, float perturb_epsilon, uint64_t seed,
                           int8_t *init_state_out, void *stream);

/* ---- a10: the S-step loop as CUDA graphs of (policy_step, q_select, env_step); replaces the Python
 *      loop of DQNSolver.rollout (solvers/solver.py:316-362).  No host round trip per step. ---- */
typedef struct ecord_rollout_desc {
    const ecord_graph   *graph;
    const ecord_env     *env;
    const ecord_weights *weights;
    const ecord_policy  *policy;
    int32_t  gemm_mode;       /* ECORD_GEMM_* */
    int32_t  q_mode;          /* ECORD_Q_* */
    int32_t  compat;          /* quirk A on/off */
    float    epsilon;         /* epsilon-greedy: probability of a uniformly random action per (g,t) and step
                                 (actors/ecord.py:171-180); 0 = greedy / tau sampling only */
    float    tau;
    int32_t  steps_per_graph; /* steps unrolled into one CUDA graph (e.g. 16) */
    uint64_t seed;
    int32_t *step_counter;    /* device i32[1]: steps completed; the plan advances it */
    float   *score_log;       /* device f32 [max_steps][G*T] or NULL */
    int32_t *action_log;      /* device i32 [max_steps][G*T] or NULL */
    int32_t  max_steps;       /* capacity of the logs */
    int32_t  _pad;
    uint64_t *step_time_ns;   /* device u64 [max_steps] or NULL: %globaltimer (ns) when step s has been applied -- the per-step
                                 timings DQNSolver.rollout logs as t_step and uses for t_opt / max_time (solver.py:241,318-341) */
} ecord_rollout_desc;

typedef struct ecord_rollout_plan ecord_rollout_plan;   /* opaque */
int ecord_rollout_plan_create(const ecord_rollout_desc *desc, void *stream, ecord_rollout_plan **out);
/* enqueue `num_steps` more steps on the plan's stream; asynchronous.  launches_out (optional, host)
 * receives the number of kernels enqueued. */
int ecord_rollout_plan_run(ecord_rollout_plan *plan, int32_t num_steps, int64_t *launches_out);
/* a new rollout on the same buffers (the caller has re-initialised the environment and zeroed step_counter):
 * the plan's captured graphs stay valid, only its count of enqueued steps (log overrun check) restarts. */
int ecord_rollout_plan_reset(ecord_rollout_plan *plan);
int ecord_rollout_plan_destroy(ecord_rollout_plan *plan);

/* ---- a11 (device part): best score per (g,t) over steps 1..S is env.best_score restricted to s>=1;
 *      since best_score also covers the initial score, the log-based maximum is provided:
 *      out_best[g*T+t] = max_s score_log[s][g*T+t]  (validate/testing.py:201). ---- */
int ecord_score_log_max(const float *score_log, int32_t num_steps, int32_t rows, float *out_best, void *stream);

/* log['t_opt'] (solver.py:241-242): out_steps[g*T+t] = the last step (1-based) at which the unit's score rose and equals its
 * running best (init_score included); 0 if the initial score was never improved on. */
int ecord_t_opt_steps(const float *score_log, const float *init_score, int32_t num_steps, int32_t rows,
                      int32_t *out_steps, void *stream);
/* ---- log_stats views (solver.py:322-327; tradjectories.py:378-385): spins and peeks of the current state in the reference's
 *      padded layout f32 [G][n_max][T] (padding -1 / -1e10).  Either output may be NULL. ---- */
int ecord_env_snapshot_padded(const ecord_graph *g, const ecord_env *env, float *state_out, float *peek_out, void *stream);
/* one %globaltimer sample (ns) on the stream: the time base of ecord_rollout_desc.step_time_ns */
int ecord_stamp_time(uint64_t *out_ns, void *stream);

/* ---- layout helper: trajectory-major plane f32 [N*T] -> the reference's [N][T] ---- */
int ecord_plane_to_nt(const ecord_graph *g, int32_t num_tradj, const float *plane, float *out, void *stream);

/* ---- true best-state snapshot (what quirk D loses): replays action_log up to best_step.
 *      out: u8 plane [N*T] ---- */
int ecord_best_state_snapshot(const ecord_graph *g, const ecord_env *env, const int8_t *init_state,
                              const int32_t *action_log, int32_t num_steps, uint8_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ECORD_B200_H */
