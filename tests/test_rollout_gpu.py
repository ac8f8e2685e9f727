"""Fused rollout (CUDA-graph plan) and the solver / validate API on the GPU."""
import numpy as np
import pytest
import torch

import ecord_oracle as orc
from ecord_b200.data import GraphBatcher
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from util import load_gold, unpack_graphs

pytestmark = pytest.mark.gpu


def _solver(sd, **kw):
    from ecord_b200.solver import B200Solver
    return B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], **kw)


def test_free_running_greedy_reproduces_reference_trace():
    """fp32 mode, no teacher forcing: 24 greedy steps through the plan give the reference's actions and scores."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("model_trace.npz")
    sd = random_state_dicts(int(z["weight_seed"]), perturb_norms=True)
    dg = DeviceGraph(GraphBatcher()(unpack_graphs(z)).tg)
    T, S = int(z["T"]), int(z["S"])
    st = RolloutState(dg, DeviceWeights(sd), T, gemm="fp32", max_steps=S, steps_per_graph=5)   # 4 graphs of 5 + 4 of 1
    st.gnn_embed()
    st.env_init(z["init_state"])
    st.run(S)
    acts = st.action_log[:S].view(S, dg.G, T).cpu().numpy()
    assert np.array_equal(acts, z["actions"].astype(np.int32))
    assert np.array_equal(st.score_log[:S].view(S, dg.G, T).cpu().numpy(), z["scores"])
    assert np.array_equal(st.best_states_batched().numpy(), z["best_state"])
    assert int(st.step_counter.item()) == S
    # the true snapshot (what quirk D loses) must have exactly the best score as its cut value
    snap = st.best_states_batched(true_snapshot=True).numpy()
    ob = orc.build_batch(unpack_graphs(z))
    flat = np.concatenate([snap[g, :ob.num_nodes_list[g]] for g in range(ob.G)]).astype(np.int8)
    env = orc.EnvOracle(ob, T, flat)
    assert np.array_equal(env.scores, st.best_score.view(dg.G, T).cpu().numpy())


@pytest.mark.parametrize("case", range(6))
def test_free_running_rollout_matches_oracle_on_random_shapes(case):
    """Shape fuzz against the oracle rollout (itself pinned to the unmodified reference): ragged batches with graphs smaller
    and larger than a 128-vertex tile, single graphs / single trajectories, isolated vertices, signed and unit weights,
    with and without quirk A.  fp32 mode, greedy, free running: the actions and scores of every (graph, trajectory)
    must equal the oracle's step for step.  A unit may leave the oracle's trajectory only at a near-tie of its best
    Q-values: the random-init network's Q is nearly flat and structurally equivalent vertices tie to the last bit, so the
    fp32 re-association of the folded head can pick the other one.  Such a deviation must start at a step where the
    oracle's own Q of the vertex the GPU chose lies within 1e-5 relative of the maximum; a real defect shows gaps of
    1e-2 and more."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    rs = np.random.RandomState(100 + case)
    shapes = [([5, 1, 37], 3), ([130], 1), ([64, 200, 17, 129], 5), ([3, 3, 3, 3, 3, 3, 3, 3, 3], 2), ([260, 40], 7), ([128, 127], 4)][case]
    ns, T = shapes
    graphs = []
    for i, n in enumerate(ns):
        p = [0.5, 0.1, 0.25][i % 3] if n > 1 else 0.0
        graphs += synthetic_set("er", n, 1, p=p, signed=bool((case + i) % 2), seed0=1000 * case + i) if n > 1 else \
            [orc_single_vertex()]
    compat = case % 2 == 0
    sd = random_state_dicts(20 + case, perturb_norms=True)
    S = 14
    ob = orc.build_batch(graphs)
    init = rs.randint(0, 2, (ob.N, T)).astype(np.int8)
    ref = orc.rollout(orc.ModelOracle(sd, compat=compat), ob, T, S, init, tau=0.0, record=True)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    st = RolloutState(dg, DeviceWeights(sd), T, gemm="fp32", compat=compat, max_steps=S, steps_per_graph=4)
    st.gnn_embed(); st.env_init(init); st.run(S)
    acts = st.action_log[:S].view(S, dg.G, T).cpu().numpy()
    scores = st.score_log[:S].view(S, dg.G, T).cpu().numpy()
    ref_acts = torch.stack(ref["actions"]).numpy()
    ref_scores = torch.stack(ref["scores"]).numpy()
    same = (acts == ref_acts).all(0) & (scores == ref_scores).all(0)              # per unit, over all steps
    assert same.mean() >= 0.5, f"{(~same).sum()} of {same.size} units left the oracle's trajectory"
    for g, t in zip(*np.nonzero(~same)):                                          # a deviation must start at a near-tie
        s0 = int(np.nonzero(acts[:, g, t] != ref_acts[:, g, t])[0][0])
        assert (scores[:s0, g, t] == ref_scores[:s0, g, t]).all()
        q = ref["q"][s0][ob.ptr[g]:ob.ptr[g + 1], t].double()
        gap = float(q.max() - q[int(acts[s0, g, t])])
        assert gap <= 1e-5 * max(1.0, float(q.abs().max())), (g, t, s0, gap)


def orc_single_vertex():
    from ecord_b200.data import EdgeListGraph
    return EdgeListGraph(1, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))


@pytest.mark.parametrize("gemm", ["fp32", "bf16x3", "bf16"])
def test_plan_equals_stepwise(gemm):
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    graphs = synthetic_set("er", 60, 7, p=0.2)
    sd = random_state_dicts(5, perturb_norms=True)
    w = DeviceWeights(sd)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    T, S = 9, 37
    init = np.random.RandomState(2).randint(0, 2, (dg.N, T)).astype(np.int8)
    a = RolloutState(dg, w, T, gemm=gemm, max_steps=S, steps_per_graph=8)
    b = RolloutState(dg, w, T, gemm=gemm, max_steps=S)
    for st in (a, b):
        st.gnn_embed(); st.env_init(init)
    a.run(20); a.run(S - 20)
    for _ in range(S):
        b.step()
    assert torch.equal(a.action_log[:S], b.action_log[:S])
    assert torch.equal(a.score_log[:S], b.score_log[:S])
    assert torch.equal(a.hidden_state(), b.hidden_state())
    assert torch.equal(a.peek, b.peek) and torch.equal(a.sp, b.sp)


def test_config1_best_cut_parity():
    """BASELINE config 1: ER40 validation set (100 graphs), 80 steps, 50 trajectories, graph batch 50, greedy.
    Bar (north_star): best cut equal to the reference's on >= 99% of instances (fp32 mode)."""
    from ecord_b200.testing import TestGraphsConfig, test_solver
    z = load_gold("config1.npz")
    graphs = unpack_graphs(z)
    T, S, bsz = int(z["T"]), int(z["S"]), int(z["bsz"])
    solver = _solver(random_state_dicts(int(z["weight_seed"])), gemm="fp32")
    cfg = TestGraphsConfig(graphs=graphs, opt_scores=torch.tensor(z["opts"]), num_steps=S, tradj_per_graph=T, tau=0,
                           graphs_bsz=bsz, tradj_bsz=T)
    np.random.seed(int(z["state_seed"]))       # the solver draws initial spins exactly as the reference does
    log = test_solver(solver, GraphBatcher(), cfg)
    assert np.array_equal(log["init_score"].numpy(), z["init_score"].astype(np.float32))
    agree = float((log["scores_max"].numpy() == z["scores_max"]).mean())
    assert agree >= 0.99, f"best-cut agreement {agree:.3f}"
    assert log["scores"].shape == (100, T, S) and log["t_step"].shape == (2 * S,)   # merge_log concatenates graph chunks (testing.py:120-143)
    assert np.allclose(log["apx_max"].numpy(), log["scores_max"].numpy() / z["opts"])


@pytest.mark.parametrize("gemm,bar", [("bf16x3", 0.99), ("bf16", 0.97)])
def test_config1_best_cut_tensor_core_modes(gemm, bar):
    """The same config through the tensor-core paths (tcgen05 GRU / projection / tail, fp16-split tensor-core Q).
    "bf16x3" (the default and the benchmarked mode: bf16 hi / lo operand pairs, fp32-grade gate functions) is held to the
    north_star bar of 99 %.  The opt-in fast mode "bf16" rounds the recurrent update to bf16, which can change a near-tie
    arg-max: 97 % (measured 99-100 %).  Both: per-graph mean-over-trajectories within 1 %."""
    from ecord_b200.testing import TestGraphsConfig, test_solver
    z = load_gold("config1.npz")
    graphs = unpack_graphs(z)
    T, S, bsz = int(z["T"]), int(z["S"]), int(z["bsz"])
    solver = _solver(random_state_dicts(int(z["weight_seed"])), gemm=gemm)
    cfg = TestGraphsConfig(graphs=graphs, opt_scores=torch.tensor(z["opts"]), num_steps=S, tradj_per_graph=T, tau=0,
                           graphs_bsz=bsz, tradj_bsz=T)
    np.random.seed(int(z["state_seed"]))
    log = test_solver(solver, GraphBatcher(), cfg)
    agree = float((log["scores_max"].numpy() == z["scores_max"]).mean())
    print(f"{gemm}: best-cut agreement with the reference on {agree:.3f} of the graphs")
    assert agree >= bar, f"best-cut agreement {agree:.3f}"
    if "scores_mean" in z.files:
        rel = np.abs(log["scores_mean"].numpy() - z["scores_mean"]) / np.maximum(np.abs(z["scores_mean"]), 1.0)
        assert float(rel.mean()) < 0.01


def test_solver_log_contract_and_checkpoint_roundtrip(tmp_path):
    """The keys / shapes downstream code reads (SURVEY.md 8b "Log contract"), and DQNSolver.save/load ingestion."""
    graphs = synthetic_set("er", 30, 3, p=0.3) + synthetic_set("er", 20, 2, p=0.3, seed0=7)     # ragged
    sd = random_state_dicts(1)
    solver = _solver(sd, gemm="bf16")
    batch = GraphBatcher()(graphs)
    T = 4
    log = solver.rollout(batch, num_tradj=T, num_steps=-2, use_epsilon=False, tau=0, log_stats=False)
    S = 2 * 30
    assert len(log["scores"]) == S and log["scores"][0].shape == (5, T)
    assert log["init_score"].shape == (5, T) and log["t_opt"].shape == (5, T)
    assert log["best_state"].shape == (5, 30, T) and (log["best_state"][3:, 20:] == -1).all()
    for k in ("t_setup", "t_init_emb"):
        assert isinstance(log[k], list) and log[k][0].shape == (1,)
    assert len(log["t_step"]) == S and log["t_tot"].shape == (1,) and log["t_rollout"].shape == (1,)
    # log_stats adds per-step actions / peeks / state in the reference's padded layout
    log2 = solver.rollout(batch, num_tradj=T, num_steps=5, use_epsilon=False, tau=0, log_stats=True,
                          init_state=np.zeros((batch.tg.num_nodes, T), np.int8))
    assert log2["actions"][0].shape == (5, T) and log2["peeks"][0].shape == (5, 30, T) and log2["state"][0].shape == (5, 30, T)
    assert (log2["peeks"][0][3:, 20:] == -1e10).all()
    # checkpoint in the reference's format -> load
    f = solver.save(str(tmp_path / "ck"))
    other = _solver(random_state_dicts(99), gemm="bf16")
    other.load(f, quiet=True)
    a = solver.rollout(batch, num_tradj=T, num_steps=8, use_epsilon=False, tau=0, init_state=np.ones((batch.tg.num_nodes, T), np.int8))
    b = other.rollout(batch, num_tradj=T, num_steps=8, use_epsilon=False, tau=0, init_state=np.ones((batch.tg.num_nodes, T), np.int8))
    assert torch.equal(torch.stack(a["scores"]), torch.stack(b["scores"]))


@pytest.mark.parametrize("gemm", ["fp32", "bf16x3", "bf16"])
def test_rollout_state_reuse_is_invisible(gemm):
    """The solver keeps the device state of a graph batch and re-initialises it for the next trajectory chunk
    (testing.py:174-189 calls rollout repeatedly on one batch): every call must equal a call on a fresh solver."""
    graphs = synthetic_set("er", 40, 4, p=0.2)
    sd = random_state_dicts(3)
    batch = GraphBatcher()(graphs)
    T, N = 6, batch.tg.num_nodes
    inits = [np.random.RandomState(s).randint(0, 2, (N, T), dtype=np.int8) for s in (1, 2, 1)]
    steps = [40, 24, 40]                       # the second call reuses the state with fewer steps than its logs hold
    kw = dict(num_tradj=T, use_epsilon=False, tau=0)
    reused = _solver(sd, gemm=gemm)
    logs = [reused.rollout(batch, num_steps=S, init_state=i0, post_solve_with_greedy_from_best=(k == 1), **kw)
            for k, (S, i0) in enumerate(zip(steps, inits))]
    assert len(reused._state_cache) == 1
    for k, (S, i0) in enumerate(zip(steps, inits)):
        fresh = _solver(sd, gemm=gemm).rollout(batch, num_steps=S, init_state=i0,
                                               post_solve_with_greedy_from_best=(k == 1), **kw)
        assert len(logs[k]["scores"]) == len(fresh["scores"])
        assert torch.equal(torch.stack(logs[k]["scores"]), torch.stack(fresh["scores"]))
        assert torch.equal(logs[k]["best_state"], fresh["best_state"])
        assert torch.equal(logs[k]["t_opt"] > 0, fresh["t_opt"] > 0)
        assert torch.equal(logs[k]["best_score_by_tradj"], fresh["best_score_by_tradj"])
    assert torch.equal(torch.stack(logs[0]["scores"]), torch.stack(logs[2]["scores"]))


def test_log_stats_matches_oracle_environment_replay():
    """log_stats (solver.py:244-249,322-327): per-step `state` and `peeks` are the environment BEFORE the step's flip in
    the padded [G, n_max, T] layout (fill -1 / -1e10), `actions` the local vertex ids.  Replaying the logged actions in
    the oracle environment must reproduce every logged plane and score bit for bit (ragged batch, fp32 mode)."""
    from util import oracle_batch
    graphs = synthetic_set("er", 24, 2, p=0.3) + synthetic_set("er", 17, 2, p=0.3, seed0=5)
    sd = random_state_dicts(2)
    batch = GraphBatcher()(graphs)
    T, S, N = 3, 12, batch.tg.num_nodes
    init = np.random.RandomState(4).randint(0, 2, (N, T), dtype=np.int8)
    log = _solver(sd, gemm="fp32").rollout(batch, num_tradj=T, num_steps=S, use_epsilon=False, tau=0, log_stats=True,
                                           init_state=init)
    ob = oracle_batch(graphs)
    env = orc.EnvOracle(ob, T, init)
    assert len(log["actions"]) == len(log["peeks"]) == len(log["state"]) == S
    for s in range(S):
        for gi in range(ob.G):
            lo, hi = ob.ptr[gi], ob.ptr[gi + 1]
            np.testing.assert_array_equal(log["state"][s][gi, :hi - lo].numpy(), env.node_features[lo:hi, :, 0])
            np.testing.assert_array_equal(log["peeks"][s][gi, :hi - lo].numpy(), env.node_features[lo:hi, :, 1])
            assert (log["state"][s][gi, hi - lo:] == -1).all() and (log["peeks"][s][gi, hi - lo:] == -1e10).all()
        env.step(log["actions"][s].numpy())
        np.testing.assert_array_equal(log["scores"][s].numpy(), env.scores)
    np.testing.assert_array_equal(log["best_state"].numpy(), env.best_states_batched().numpy())


def test_use_epsilon_follows_the_reference_schedule():
    """solver.py:184-186,334: rollout(use_epsilon=True) explores with epsilon = get_epsilon() (linear in num_param_steps),
    also outside training; a fresh solver therefore acts fully at random, one past final_epsilon_step mostly greedily."""
    graphs = synthetic_set("er", 30, 2, p=0.3)
    sd = random_state_dicts(1)
    batch = GraphBatcher()(graphs)
    T, S = 64, 20
    init = np.random.RandomState(0).randint(0, 2, (batch.tg.num_nodes, T), dtype=np.int8)
    solver = _solver(sd, gemm="fp32", final_epsilon=0.05, final_epsilon_step=100)
    assert solver.get_epsilon() == 1.0
    kw = dict(num_tradj=T, num_steps=S, tau=0, init_state=init, log_stats=True)
    greedy = torch.stack(solver.rollout(batch, use_epsilon=False, **kw)["actions"])
    rnd = torch.stack(solver.rollout(batch, use_epsilon=True, **kw)["actions"])
    assert float((rnd != greedy).float().mean()) > 0.8                     # epsilon = 1: uniform random vertices
    solver.num_param_steps = 1000
    assert abs(solver.get_epsilon() - 0.05) < 1e-12
    some = torch.stack(solver.rollout(batch, use_epsilon=True, **kw)["actions"])
    first = float((some[0] != greedy[0]).float().mean())                  # step 0: same state in both runs
    assert 0.0 < first < 0.2


def test_network_initialisation_samples_the_node_classifier(tmp_path):
    """use_network_initialisation (solver.py:256-277): init spins ~ Bernoulli(sigmoid(head(gnn_embedding))).  The
    reference samples with torch's RNG, so the per-vertex frequencies are what can match: they are checked against
    sigmoid(W gnn_out + b) computed on the host from the (parity-tested) GNN embedding; a solver without the head falls
    back to the random default, and the head survives a checkpoint round trip."""
    graphs = synthetic_set("er", 20, 2, p=0.3)
    sd = random_state_dicts(6, perturb_norms=True)
    batch = GraphBatcher()(graphs)
    gen = torch.Generator().manual_seed(0)
    heads = {"heads.0.0.weight": torch.randn(1, 16, generator=gen) * 2, "heads.0.0.bias": torch.zeros(1),
             "heads.1.0.weight": torch.randn(1, 16, generator=gen) * 2, "heads.1.0.bias": torch.randn(1, generator=gen)}
    one = {k: v for k, v in heads.items() if k.startswith("heads.0")}
    T = 4096
    solver = _solver(sd, gemm="fp32", node_classifier=one)
    log = solver.rollout(batch, num_tradj=T, num_steps=1, tau=0, use_epsilon=False, use_network_initialisation=True,
                         log_stats=True)
    state0 = torch.cat([log["state"][0][g, :20] for g in range(2)]).numpy()            # [N, T] spins before step 0
    gnn_out = solver._last_state.gnn_out.cpu().double()
    prob = torch.sigmoid(gnn_out @ one["heads.0.0.weight"].double().view(-1) + one["heads.0.0.bias"].double()).numpy()
    freq = state0.mean(1)
    assert np.all(np.abs(freq - prob) < 5 * np.sqrt(np.maximum(prob * (1 - prob), 1e-4) / T) + 1e-3), np.abs(freq - prob).max()
    assert prob.max() - prob.min() > 0.05                      # the head is not trivially constant
    # a second call is a fresh draw; the init score is the cut of the drawn spins
    log_b = solver.rollout(batch, num_tradj=T, num_steps=1, tau=0, use_epsilon=False, use_network_initialisation=True,
                           log_stats=True)
    assert not torch.equal(log["state"][0], log_b["state"][0])
    # two heads + use_epsilon: spins are re-drawn uniformly with probability epsilon (= 1 for a fresh solver)
    two = _solver(sd, gemm="fp32", node_classifier=heads)
    log2 = two.rollout(batch, num_tradj=T, num_steps=1, tau=0, use_epsilon=True, use_network_initialisation=True, log_stats=True)
    s2 = torch.cat([log2["state"][0][g, :20] for g in range(2)]).numpy()
    assert np.all(np.abs(s2.mean(1) - 0.5) < 5 * 0.5 / np.sqrt(T))
    # no head -> the reference's fallback: random default initialisation
    plain = _solver(sd, gemm="fp32")
    np.random.seed(3)
    a = plain.rollout(batch, num_tradj=8, num_steps=2, tau=0, use_epsilon=False, use_network_initialisation=True)
    np.random.seed(3)
    b = plain.rollout(batch, num_tradj=8, num_steps=2, tau=0, use_epsilon=False)
    assert torch.equal(a["init_score"], b["init_score"])
    # checkpoint round trip (actors/ecord.py:250-251,269)
    f = two.save(str(tmp_path / "ck"))
    other = _solver(random_state_dicts(9), gemm="fp32")
    other.load(f, quiet=True)
    assert other.node_classifier is not None and torch.equal(other.node_classifier[0].cpu(), two.node_classifier[0].cpu())


def test_unsupported_options_fail_loudly():
    sd = random_state_dicts(1)
    from ecord_b200.solver import B200Solver
    from ecord_b200.observations import NodeObservation
    with pytest.raises(NotImplementedError):
        B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'],
                   node_features=[NodeObservation.STATE, NodeObservation.PEEK])
    with pytest.raises(NotImplementedError):           # the ECO-DQN actor and the double-Q actor are outside the hot path
        _solver(sd, use_ecodqn=True)
    with pytest.raises(NotImplementedError):
        _solver(sd, use_double_q_networks=True)
    with pytest.raises(NotImplementedError):           # extra rewards of the ECO-DQN training setting (environment.py:116-139)
        _solver(sd, intermediate_reward=1 / 40).train()
    # in test mode the reference ignores callbacks (solver.py:348: `if callbacks is not None and self.is_training`)
    solver = _solver(sd)
    batch = GraphBatcher()(synthetic_set("er", 10, 1, p=0.5))
    calls = []
    solver.rollout(batch, num_tradj=2, num_steps=2, callbacks=[(lambda *a, **k: calls.append(1), 1, None)], use_epsilon=False)
    assert calls == []


@pytest.mark.parametrize("kind,n,G,T,S,gen", [
    ("er", 500, 50, 50, 64, dict(p=0.15)),                                   # BASELINE configs[1]
    ("ba", 500, 50, 50, 48, dict()),                                         # configs[2]: power-law degrees (dmax > 100)
    ("er", 10000, 3, 128, 40, dict(avg_degree=10, signed=False)),            # configs[3]: ER10000 x 128 trajectories
    ("er", 4000, 2, 128, 32, dict(avg_degree=40, signed=True)),              # configs[4]: GSet-shaped, degree 40, +-1
], ids=["er500", "ba500", "er10000", "gset4000d40"])
def test_full_size_properties(kind, n, G, T, S, gen):
    """BASELINE config sizes (steps truncated): size-independent properties of the fused throughput-mode rollout --
    the final score is the cut of the final spins, the peeks equal a from-scratch recomputation, the best score is the
    maximum of the log, actions stay inside their graph, and the static counters equal the steps since the last flip."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    graphs = synthetic_set(kind, n, G, **gen)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    st = RolloutState(dg, DeviceWeights(random_state_dicts(0)), T, gemm="bf16x3", max_steps=S)
    init = np.random.RandomState(1).randint(0, 2, (dg.N, T)).astype(np.int8)
    st.gnn_embed(); st.env_init(init); st.run(S)
    nf, gf, bs = st.export()
    ob = orc.build_batch(graphs)
    fresh = orc.EnvOracle(ob, T, nf[..., 0].cpu().numpy().astype(np.int8))
    assert np.array_equal(fresh.scores, st.scores().cpu().numpy())                 # score == cut of the final spins
    assert np.array_equal(fresh.node_features[..., 1], nf[..., 1].cpu().numpy())   # peeks == recomputed
    log = st.score_log[:S].view(S, dg.G, T)
    init_sc = orc.EnvOracle(ob, T, init).scores
    assert np.array_equal(np.maximum(log.max(0).values.cpu().numpy(), init_sc), st.best_score.view(dg.G, T).cpu().numpy())
    acts = st.action_log[:S].view(S, dg.G, T)
    assert int(acts.min()) >= 0 and int(acts.max()) < n
    # static feature == steps since the vertex was last flipped
    last = torch.zeros(dg.G, T, n, dtype=torch.long)
    for s in range(S):
        last.scatter_(2, acts[s].cpu().long().unsqueeze(-1), s + 1)
    static = nf[..., 2].cpu().view(dg.G, n, T).permute(0, 2, 1)
    assert torch.equal(static.long(), S - last)


# ---- greedy pre-/post-solve (reference solvers/solver.py:38-95, 294-295, 367-371) ----
def _greedy_state(gemm="fp32"):
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("greedy_trace.npz")
    graphs = unpack_graphs(z)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    st = RolloutState(dg, DeviceWeights(random_state_dicts(0)), int(z["T"]), gemm=gemm, max_steps=16)
    st.gnn_embed(); st.env_init(z["init_state"])
    return z, graphs, dg, st


def test_greedy_solve_matches_reference_on_tie_free_graphs():
    """Distinct dyadic weights -> no ties -> greedy_solve is deterministic: the engine must reproduce the unmodified
    reference's scores after the solve, after 7 scripted steps + greedy_solve(init_from_best=True), and its best scores
    (fixture written by oracle/check_greedy.py, which also asserts reference == oracle)."""
    z, graphs, dg, st = _greedy_state()
    T = int(z["T"])
    iters = st.greedy_solve()
    assert iters == int(z["steps"])
    assert np.array_equal(st.scores().cpu().numpy(), z["scores_after_greedy"])
    rs = np.random.RandomState(int(z["scripted_seed"]))
    for i in range(int(z["scripted_steps"])):
        a = np.stack([rs.randint(0, int(n), T) for n in dg.num_nodes_list])
        st.env_step(a.astype(np.int32).reshape(-1))
    iters2 = st.greedy_solve(init_from_best=True)
    assert iters2 == int(z["steps_post"])
    assert np.array_equal(st.scores().cpu().numpy(), z["scores_after_post"])
    assert np.array_equal(st.best_score.view(dg.G, T).cpu().numpy(), z["best_after_post"])
    # the state is self-consistent: score == cut of the spins, peeks == recomputed
    nf, _, _ = st.export()
    fresh = orc.EnvOracle(orc.build_batch(graphs), T, nf[..., 0].cpu().numpy().astype(np.int8))
    assert np.array_equal(fresh.scores, st.scores().cpu().numpy())
    assert np.array_equal(fresh.node_features[..., 1], nf[..., 1].cpu().numpy())


def test_greedy_tie_break_is_uniform_and_reaches_local_optima():
    """+-1 weights: ties everywhere.  (1) On a perfect matching with all spins equal every vertex has peek +1: the first
    action must be (roughly) uniform over the vertices.  (2) After the solve every unit has been at a local optimum:
    its best score is >= the score of a fresh single-flip local search bound, and the state stays self-consistent."""
    import networkx as nx
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    n, T = 16, 4096
    g = nx.Graph(); g.add_nodes_from(range(n))
    for i in range(0, n, 2):
        g.add_edge(i, i + 1, weight=1.0)
    dg = DeviceGraph(GraphBatcher()([g]).tg)
    st = RolloutState(dg, DeviceWeights(random_state_dicts(0)), T, gemm="fp32", max_steps=4)
    st.gnn_embed(); st.env_init(np.zeros((n, T), np.int8))
    acts = torch.empty(T, dtype=torch.int32, device="cuda"); flags = torch.empty_like(acts)
    import ctypes as C
    from ecord_b200._lib import check, ptr
    check(st.lib.ecord_greedy_select(C.byref(dg.c), C.byref(st.env_c), 0.0, 7, 0, ptr(acts), ptr(flags), None))
    torch.cuda.synchronize()
    counts = np.bincount(acts.cpu().numpy(), minlength=n)
    assert counts.min() > 0.7 * T / n and counts.max() < 1.3 * T / n, counts
    assert int(flags.sum()) == 0
    iters = st.greedy_solve()
    assert iters >= n // 2                                     # every edge has to be cut: one flip per pair
    assert torch.all(st.best_score == n // 2)                  # the optimum of a perfect matching
    graphs = synthetic_set("er", 80, 4, p=0.1)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    T = 16
    st = RolloutState(dg, DeviceWeights(random_state_dicts(0)), T, gemm="bf16", max_steps=4)
    init = np.random.RandomState(4).randint(0, 2, (dg.N, T)).astype(np.int8)
    st.gnn_embed(); st.env_init(init)
    s0 = st.scores().clone()
    st.greedy_solve()
    assert torch.all(st.best_score.view(dg.G, T) >= s0)
    nf, _, _ = st.export()
    fresh = orc.EnvOracle(orc.build_batch(graphs), T, nf[..., 0].cpu().numpy().astype(np.int8))
    assert np.array_equal(fresh.scores, st.scores().cpu().numpy())
    assert np.array_equal(fresh.node_features[..., 1], nf[..., 1].cpu().numpy())


@pytest.mark.parametrize("pre,post", [(False, False), (True, False), (False, True), (True, True)])
def test_true_best_state_snapshot_with_greedy_solves(pre, post):
    """compat=False: log['best_state'] is the configuration whose cut IS the best score -- also when the rollout starts
    from greedy-improved spins (pre-solve, rebased clock) and when the best is found by the post-solve's greedy flips,
    neither of which is in the action log (the snapshot is kept on the device by the env step, ecord_env.true_snapshot)."""
    graphs = synthetic_set("er", 50, 4, p=0.15) + synthetic_set("er", 23, 2, p=0.3, seed0=11)      # ragged
    batch = GraphBatcher()(graphs)
    T, S = 6, 30
    init = np.random.RandomState(4).randint(0, 2, (batch.tg.num_nodes, T)).astype(np.int8)
    solver = _solver(random_state_dicts(8), compat=False)
    log = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init,
                         pre_solve_with_greedy=pre, post_solve_with_greedy_from_best=post)
    st = solver._last_state
    ob = orc.build_batch(graphs)
    snap = log['best_state'].numpy()
    assert (snap[4:, 23:] == -1).all()
    flat = np.concatenate([snap[g, :ob.num_nodes_list[g]] for g in range(ob.G)]).astype(np.int8)
    env = orc.EnvOracle(ob, T, flat)
    best = st.best_score.view(ob.G, T).cpu().numpy()
    assert np.array_equal(env.scores, best)
    every = torch.stack(log['scores'] + [log['init_score']], -1).max(-1).values.numpy()
    assert (best >= every).all()


def test_rollout_with_greedy_pre_and_post_solve():
    """DQNSolver.rollout(pre_solve_with_greedy=True, post_solve_with_greedy_from_best=True): init_score is taken after the
    pre-solve, one extra `scores` entry is logged after the post-solve, and both can only help the best cut."""
    graphs = synthetic_set("er", 60, 5, p=0.12)
    batch = GraphBatcher()(graphs)
    T, S = 8, 20
    init = np.random.RandomState(9).randint(0, 2, (batch.tg.num_nodes, T)).astype(np.int8)
    for gemm in ("fp32", "bf16"):
        solver = _solver(random_state_dicts(3), gemm=gemm)
        plain = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)
        both = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init,
                              pre_solve_with_greedy=True, post_solve_with_greedy_from_best=True)
        assert len(plain['scores']) == S and len(both['scores']) == S + 1
        assert both['greedy_pre_steps'][0] > 0
        assert torch.all(both['init_score'] >= plain['init_score'])
        best_plain = torch.stack(plain['scores'], -1).max(-1).values
        best_both = torch.maximum(torch.stack(both['scores'], -1).max(-1).values, both['init_score'])
        assert best_both.mean() >= best_plain.mean()


def test_t_opt_follows_the_reference_definition():
    """log['t_opt'] (solver.py:241-242): the accumulated time (embedding + steps) at the LAST step where a trajectory's score
    rose and equals its running best -- recomputed here on the host from the logged scores and the per-step device times."""
    graphs = synthetic_set("er", 60, 6, p=0.15)
    batch = GraphBatcher()(graphs)
    T, S = 5, 90
    init = np.random.RandomState(12).randint(0, 2, (batch.tg.num_nodes, T)).astype(np.int8)
    log = _solver(random_state_dicts(6)).rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)
    scores = torch.stack(log['scores'], -1).numpy()                       # [G,T,S]
    prev = np.concatenate([log['init_score'].numpy()[..., None], scores[..., :-1]], -1)
    runbest = np.maximum.accumulate(np.concatenate([log['init_score'].numpy()[..., None], scores], -1), -1)[..., 1:]
    hit = (scores > prev) & (scores == runbest)
    t_emb = float(log['t_init_emb'][0])
    cum = t_emb + np.cumsum(np.asarray(log['t_step'], np.float64))
    want = np.full(hit.shape[:2], t_emb)
    for g, t in zip(*np.nonzero(hit.any(-1))):
        want[g, t] = cum[np.nonzero(hit[g, t])[0][-1]]
    assert np.allclose(log['t_opt'].numpy(), want, rtol=1e-5, atol=1e-7)
    assert hit.any() and (np.asarray(log['t_step']) > 0).all()
    assert abs(sum(log['t_step']) - float(log['t_rollout'])) < 0.5 * float(log['t_rollout']) + 1e-3


def test_max_time_stops_the_rollout_at_step_granularity():
    """DQNSolver.rollout(max_time=...) (solver.py:318-320): the loop stops before the first step that starts after embedding +
    step time has reached the budget.  The engine reads the device clock stamped at the end of every step; far from the
    limit it launches whole CUDA graphs, close to it single steps, so the overshoot is at most one step."""
    graphs = synthetic_set("er", 200, 20, p=0.1)
    batch = GraphBatcher()(graphs)
    solver = _solver(random_state_dicts(3), gemm="bf16")
    T, S = 16, 4000
    init = np.random.RandomState(2).randint(0, 2, (batch.tg.num_nodes, T)).astype(np.int8)
    solver.rollout(batch, num_tradj=T, num_steps=32, tau=0, use_epsilon=False, init_state=init)      # warm-up
    log = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init, max_time=0.02)
    n = len(log['scores'])
    assert 16 <= n < S
    t_step = np.asarray(log['t_step'])
    used = float(log['t_init_emb'][0]) + t_step.sum()
    assert used >= 0.02 and used - t_step[-1] < 0.02 + 1e-4, (used, t_step[-1])     # stopped at the first step past the budget
    assert (t_step > 0).all() and t_step[4:].max() < 20 * np.median(t_step)
    assert len(log['t_step']) == n and log['best_score_by_tradj'].shape == (20, T)
    full = solver.rollout(batch, num_tradj=T, num_steps=n, tau=0, use_epsilon=False, init_state=init)
    assert torch.equal(torch.stack(full['scores']), torch.stack(log['scores']))     # same steps as an un-timed rollout


def test_training_mode_rollout_feeds_callbacks_like_the_reference():
    """DQNSolver.rollout in TRAINING mode (solver.py:329-365) -- per step the callbacks get the actor
    state and observation BEFORE the step, the action, the reward (environment.py:113-141) and the done flags.  Fixture:
    the unmodified reference with a recording callback (oracle/gen_golden.py train -> train_trace.npz), greedy so that the
    actions are reproducible; fp32-grade modes must reproduce actions, rewards and dones exactly and the tensors to 1e-4."""
    from util import rel_err
    z = load_gold("train_trace.npz")
    graphs = unpack_graphs(z)
    T, S = int(z["T"]), int(z["S"])
    batch = GraphBatcher()(graphs)
    for gemm in ("bf16x3", "fp32"):
        solver = _solver(random_state_dicts(int(z["weight_seed"]), perturb_norms=True), gemm=gemm)
        solver.train()
        assert solver.is_training
        rec = {k: [] for k in ("hidden", "last_sel", "node_feat", "glob_feat", "action", "reward", "done", "t_step")}

        def cb(slv, actor_state, obs, action, reward, done, t_step):
            assert slv is solver and obs.num_tradj == T and actor_state.init_node_embeddings.shape == (batch.tg.num_nodes, T, 16)
            rec["hidden"].append(actor_state.hidden_embeddings.cpu())
            rec["last_sel"].append(actor_state.last_selected_node_embeddings.cpu())
            rec["node_feat"].append(obs.node_features.cpu()); rec["glob_feat"].append(obs.glob_features.cpu())
            rec["action"].append(action.cpu()); rec["reward"].append(reward.cpu()); rec["done"].append(done.cpu().clone())
            rec["t_step"].append(t_step)

        log = solver.rollout(batch, num_tradj=T, num_steps=-1, callbacks=[(cb, 1, None)], use_epsilon=False, tau=0,
                             init_state=z["init_state"])
        assert len(rec["action"]) == S and rec["t_step"] == list(range(S))
        assert solver.num_env_steps == int(z["num_env_steps"]) and solver.num_rollouts == int(z["num_rollouts"])
        assert np.array_equal(torch.stack(rec["action"]).numpy(), z["action"].astype(np.int64)), gemm
        assert np.array_equal(torch.stack(rec["reward"]).numpy(), z["reward"]), gemm
        assert np.array_equal(torch.stack(rec["done"]).numpy(), z["done"]), gemm
        assert np.array_equal(torch.stack(rec["node_feat"]).numpy(), z["node_feat"]), gemm
        assert np.array_equal(torch.stack(rec["glob_feat"]).numpy(), z["glob_feat"]), gemm
        assert rel_err(torch.stack(rec["hidden"]).numpy()[..., ::16], z["hidden_slice"]) < 1e-4
        assert rel_err(torch.stack(rec["last_sel"]).numpy(), z["last_sel"]) < 1e-5
        assert np.array_equal(torch.stack(log["scores"]).numpy(), z["scores"])
        assert len(log["t_step"]) == S and all(t > 0 for t in log["t_step"])
        # callbacks on every 4th environment step only
        calls = []
        solver.num_env_steps = 0
        solver.rollout(batch, num_tradj=T, num_steps=10, callbacks=[(lambda slv, **kw: calls.append(kw["t_step"]), 4, None)],
                       use_epsilon=False, tau=0, init_state=z["init_state"])
        assert calls == [3, 7]
        solver.test()
        assert not solver.is_training


@pytest.mark.parametrize("gemm,compat,ragged", [("bf16x3", True, False), ("bf16x3", True, True), ("fp32", True, True),
                                                ("bf16x3", False, True)])
def test_graph_sharding_reproduces_the_full_batch_rollout(gemm, compat, ragged):
    """SURVEY.md 8e, graph axis (dist.rollout_graph_sharded): every simulated rank rolls out its own graphs (behind the
    quirk-A prefix) with the GNN embedding of the WHOLE batch (quirk B: batch-wide LayerNorm statistics); the per-trajectory
    scores of every step, the best cuts and the best states of its own graphs must equal the single-GPU rollout bit for bit,
    for 1, 2, 3 and 5 ranks."""
    from ecord_b200.dist import graph_shard, rollout_graph_sharded
    sd = random_state_dicts(5, perturb_norms=True)
    ns = [40, 27, 64, 33, 50, 21, 45] if ragged else [36] * 7
    graphs = [g for n in ns for g in synthetic_set("er", n, 1, seed0=n, p=0.2, signed=True)]
    batch = GraphBatcher()(graphs)
    T, S = 6, 40
    init = np.random.RandomState(7).randint(0, 2, (batch.tg.num_nodes, T), dtype=np.int8)
    solver = _solver(sd, gemm=gemm, compat=compat)
    full = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)
    full_scores = torch.stack(full['scores'], -1)                     # [G, T, S]
    for world in (1, 2, 3, 5):
        seen = []
        for rank in range(world):
            subset, own = graph_shard(ns, world, rank, compat=compat)
            if not own:
                continue
            log = rollout_graph_sharded(solver, batch, T, S, init, world_size=world, rank=rank)
            keep = [subset.index(g) for g in own]
            part = torch.stack(log['scores'], -1)[keep]
            assert torch.equal(part, full_scores[own]), (world, rank)
            assert torch.equal(log['best_score_by_tradj'][keep], full['best_score_by_tradj'][own])
            for k, g in zip(keep, own):
                assert torch.equal(log['best_state'][k, :ns[g]], full['best_state'][g, :ns[g]]), (world, rank, g)
            seen += own
        assert seen == list(range(len(ns)))


@pytest.mark.parametrize("gemm", ["bf16x3", "bf16"])
def test_trajectory_shards_reproduce_the_full_batch_rollout(gemm):
    """SURVEY.md 8e, trajectory axis: a shard of the trajectories (B200Solver.rollout(traj_slice=...), what every rank of
    dist.rollout_sharded runs) gives the single-GPU rollout's scores bit for bit -- a row's arithmetic may not depend on how many
    rows the launch has.  The full batch has 7000 rows (64-row CTAs in the policy tail, several 64-unit GRU tiles and several
    projection tiles per CTA pair: the multi-tile projection kernel), the shards 600-700 (32-row CTAs, narrower GRU tiles, the
    one-tile projection kernel): the case that differed before the tail's sums were given one association for both."""
    sd = random_state_dicts(9, perturb_norms=True)
    graphs = synthetic_set("er", 40, 100, seed0=3, p=0.15, signed=True)
    batch = GraphBatcher()(graphs)
    T, S = 70, 60
    init = np.random.RandomState(5).randint(0, 2, (batch.tg.num_nodes, T), dtype=np.int8)
    solver = _solver(sd, gemm=gemm)
    full = torch.stack(solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)['scores'], -1)
    for lo, hi in ((0, 6), (30, 37), (63, 70)):
        part = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init, traj_slice=(lo, hi))
        assert torch.equal(torch.stack(part['scores'], -1), full[:, lo:hi]), (lo, hi)
