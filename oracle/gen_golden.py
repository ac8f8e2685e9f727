"""oracle/gen_golden.py -- generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, read-only):
    python oracle/gen_golden.py
The reference is imported from where it lies; the uninstalled third-party packages are replaced by
oracle/shims/ (torch_scatter, torch_sparse, torch_geometric stand-ins) and the Cython step is the
reference's own .pyx compiled by oracle/build_ref.py into oracle/_ref/.  Nothing from the reference
is copied into the repo: the fixtures hold graphs as edge arrays (data), seeds, and the reference's
OUTPUTS.  While generating, every fixture is also replayed through oracle/ecord_oracle.py and the two
are asserted equal (bit-exact for the environment; exact or <=1e-6 for the fp32 model, see below).

Fixtures
  env_traces.npz   scripted-action environment traces (canonical + all-features tables, ragged batch)
  model_trace.npz  teacher-free greedy rollout of the canonical network: GNN embeddings, per-step
                   Q-values / P / hidden slice / last_sel / actions / scores (ragged 6-graph batch)
  config1.npz      BASELINE config 1 (ER40 validation, 100 graphs, 80 steps, 50 tradj, batch 50):
                   scores_max / scores_mean / per-trajectory best / init scores
  config2_er200.npz, config2_er500.npz, config3_ba500.npz
                   BASELINE configs 2 and 3 at FULL length (num_steps = -4 -> 4*|V| steps, 50 tradj, greedy tau=0):
                   the unmodified reference's test_solver on 10 graphs of testing/ER_200spin_p15_50graphs.pkl,
                   5 synthesised ER500 graphs (SURVEY.md 8d recipe; the reference's file is not shipped) and 10 graphs
                   of testing/BA_500spin_m4_50graphs.pkl: scores_max / scores_mean / per-trajectory best, best step and
                   final score, and the chosen vertices of the first 8 trajectories at every step
                   (python oracle/gen_golden.py full)
"""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np
import torch

import build_ref

assert build_ref.load_step_cy() is not None, "could not build oracle/_ref"

import ecord_oracle as orc
from ecord_b200.network import random_state_dicts

from ecord.environment import NodeObservation, GlobalObservation, Tradjectories
from ecord.environment.data import GraphBatcher
from ecord.graphs.utils import load_graph_from_file
from ecord.models.rnn_decoder import GRUDecoder, RNNOutput
from ecord.solvers import DQNSolver
from ecord.validate.testing import TestGraphsConfig, test_solver
from experiments.configs import NetworkConfig

GOLD = os.path.join(ROOT, "tests", "golden")
REF_GRAPHS = "/root/reference/graphs/ecodqn"


def load(name):
    graphs, opts = load_graph_from_file(os.path.join(REF_GRAPHS, name), quiet=True)
    return graphs, opts


def pack_graphs(graphs):
    """list of graphs -> flat edge arrays (reference edge order) for storage in an .npz."""
    ns, cnt, src, dst, w = [], [], [], [], []
    for g in graphs:
        n, s, d, ww = orc.graph_to_edges(g)
        ns.append(n); cnt.append(len(ww)); src.append(s); dst.append(d); w.append(ww)
    return dict(g_n=np.asarray(ns, np.int32), g_ne=np.asarray(cnt, np.int32),
                g_src=np.concatenate(src).astype(np.int16), g_dst=np.concatenate(dst).astype(np.int16),
                g_w=np.concatenate(w).astype(np.float32))


def canonical_network():
    return NetworkConfig(
        rnn_model=GRUDecoder,
        node_features=[NodeObservation.STATE, NodeObservation.PEEK_NORM, NodeObservation.STEPS_STATIC_CLIP],
        glob_features=[GlobalObservation.SCORE_FROM_BEST_NORM, GlobalObservation.PEEK_MAX_NORM],
        num_layers=4, dim_embedding=16, use_layer_norm=True, dim_node_out=16,
        use_network_initialisation=False, num_rnn_layers=1, dim_features_embedding=16,
        add_glob_to_nodes=False, add_glob_to_obs=True, dim_internal_embedding=1024, learn_scale=False,
        learn_graph_embedding=False, project_rnn_to_gnn=True, output_type=RNNOutput.MLP_ACTION_STATE,
        output_activation=None)


def make_solver(sd):
    nc = canonical_network()
    gnn, rnn, g2r, enc = nc.get_gnn(), nc.get_rnn(), nc.get_gnn2rnn(), nc.get_node_encoder()
    gnn.load_state_dict(sd['gnn']); rnn.load_state_dict(sd['rnn'])
    g2r.load_state_dict(sd['gnn2rnn']); enc.load_state_dict(sd['node_encoder'])
    solver = DQNSolver(gnn=gnn, rnn=rnn, gnn2rnn=g2r, node_encoder=enc, node_classifier=None, use_ecodqn=False,
                       add_glob_to_nodes=nc.add_glob_to_nodes, add_glob_to_obs=nc.add_glob_to_obs,
                       node_features=nc.node_features, global_features=nc.glob_features,
                       allow_reversible_actions=True, default_actor_idx=None, device='cpu',
                       use_double_q_networks=False)
    solver.test()
    return solver, nc.get_graph_batcher()


# ------------------------------------------------------------------------------------------------
def gen_env_traces():
    er40, _ = load("validation/ER_40spin_p15_100graphs")
    ba20, _ = load("validation/BA_20spin_m4_100graphs")
    ba60, _ = load("testing/BA_60spin_m4_50graphs")
    er100, _ = load("testing/ER_100spin_p15_50graphs")
    cases = {
        # name: (graphs, T, S, node feature list, glob feature list)
        "canon": (er40[:6], 5, 150,
                  [NodeObservation.STATE, NodeObservation.PEEK_NORM, NodeObservation.STEPS_STATIC_CLIP],
                  [GlobalObservation.SCORE_FROM_BEST_NORM, GlobalObservation.PEEK_MAX_NORM]),
        "allfeat": (er40[6:10], 4, 120,
                    [NodeObservation.STEPS_STATIC, NodeObservation.STATE_BEST, NodeObservation.STATE,
                     NodeObservation.PEEK],
                    # STEPS must precede NUM_GREEDY_ACTIONS or the reference never configures it
                    # (tradjectories.py:277-282 tests the wrong flag)
                    [GlobalObservation.STEPS, GlobalObservation.NUM_GREEDY_ACTIONS,
                     GlobalObservation.STEPS_SINCE_BEST, GlobalObservation.SCORE_FROM_BEST]),
        "ragged": ([ba20[0], er100[0], ba60[1], er40[11], ba20[3]], 7, 300,
                   [NodeObservation.STATE, NodeObservation.PEEK_NORM, NodeObservation.STEPS_STATIC_CLIP],
                   [GlobalObservation.SCORE_FROM_BEST_NORM, GlobalObservation.PEEK_MAX_NORM]),
    }
    out = {}
    gb = GraphBatcher()
    for name, (graphs, T, S, nfeat, gfeat) in cases.items():
        batch = gb(graphs)
        ns = np.array([g.number_of_nodes() for g in batch.nx])
        N = int(ns.sum())
        rs = np.random.RandomState(100 + len(name))
        init = rs.randint(0, 2, (N, T)).astype(np.int8)
        actions = np.stack([rs.randint(0, ns[:, None], (len(ns), T)) for _ in range(S)]).astype(np.int64)
        tr = Tradjectories(batch.nx, T, batch.tg, node_features=list(nfeat), glob_features=list(gfeat),
                           init_state=init.copy())
        node_idx = np.array(tr._node_features_idxs, np.int32)
        glob_idx = np.array(tr._glob_features_idxs, np.int32)
        score_log = []
        init_nf, init_gf = tr._node_features.copy(), tr._glob_features.copy()
        init_scores = tr._scores.copy()
        for s in range(S):
            tr.step(torch.from_numpy(actions[s]), apply_batch_offset=True)
            score_log.append(tr._scores.copy())
        # replay through the oracle (C restatement AND compiled reference): must be bit-identical
        ob = orc.build_batch(graphs)
        for use_ref in (False, True):
            env = orc.EnvOracle(ob, T, init, node_idx=node_idx, glob_idx=glob_idx,
                                num_node_features=tr._node_features.shape[-1],
                                num_glob_features=tr._glob_features.shape[-1], use_ref=use_ref)
            assert np.array_equal(env.node_features, init_nf), name
            assert np.array_equal(env.glob_features, init_gf), name
            assert np.array_equal(env.scores, init_scores), name
            for s in range(S):
                env.step(actions[s])
                assert np.array_equal(env.scores, score_log[s]), (name, s)
            assert np.array_equal(env.node_features, tr._node_features), name
            assert np.array_equal(env.glob_features, tr._glob_features), name
            assert np.array_equal(env.best_scores, tr._best_scores), name
            assert np.array_equal(env.best_states, tr._best_states), name
        if name != "allfeat":
            nf_obs, gf_obs = env.observe()
            assert torch.equal(nf_obs, tr.get_node_features(flat=True)), name
            assert torch.equal(gf_obs, tr.get_glob_features()), name
            out[f"{name}/obs_node"] = tr.get_node_features(flat=True).numpy()
            out[f"{name}/obs_glob"] = tr.get_glob_features().numpy()
        assert torch.equal(env.best_states_batched(), tr.get_best_states(flat=False)), name
        for k, v in pack_graphs(graphs).items():
            out[f"{name}/{k}"] = v
        out[f"{name}/T"] = np.int32(T)
        out[f"{name}/init_state"] = init
        out[f"{name}/actions"] = actions.astype(np.int16)
        out[f"{name}/node_idx"] = node_idx
        out[f"{name}/glob_idx"] = glob_idx
        out[f"{name}/init_node_features"] = init_nf
        out[f"{name}/init_glob_features"] = init_gf
        out[f"{name}/init_scores"] = init_scores
        out[f"{name}/score_log"] = np.stack(score_log)
        out[f"{name}/node_features"] = tr._node_features.copy()
        out[f"{name}/glob_features"] = tr._glob_features.copy()
        out[f"{name}/best_scores"] = tr._best_scores.copy()
        out[f"{name}/best_states"] = tr._best_states.copy()
        print(f"env trace '{name}': G={len(graphs)} N={N} T={T} S={S}  oracle == reference (bit-exact)")
    np.savez_compressed(os.path.join(GOLD, "env_traces.npz"), **out)


# ------------------------------------------------------------------------------------------------
def gen_model_trace():
    er40, _ = load("validation/ER_40spin_p15_100graphs")
    ba20, _ = load("validation/BA_20spin_m4_100graphs")
    ba60, _ = load("testing/BA_60spin_m4_50graphs")
    graphs = [er40[20], ba20[5], ba60[2], er40[21], ba60[3], er40[22]]
    T, S, wseed, sseed = 5, 24, 7, 4321
    sd = random_state_dicts(wseed, perturb_norms=True)
    solver, gb = make_solver(sd)
    batch = gb(graphs)
    rec = {k: [] for k in ("q", "hidden", "P", "last_sel", "actions")}
    actor = solver.actor
    orig = actor.action_select

    def wrapped(obs, **kw):
        ret = orig(obs, **kw)
        rec["actions"].append(ret[0].clone())
        rec["q"].append(ret[1].clone())
        h = actor.rnn.hidden_embeddings[-1]
        rec["hidden"].append(h.clone())
        rec["P"].append(actor.rnn.proj_rnn_to_gnn(h).clone())
        rec["last_sel"].append(actor._last_selected_node_embeddings.clone())
        return ret

    actor.action_select = wrapped
    np.random.seed(sseed)
    log = solver.rollout(batch, num_tradj=T, num_steps=S, use_epsilon=False, tau=0)
    gnn_emb = actor.gnn_node_embeddings[:, 0].clone()       # [N,16] GNN output (pre gnn2rnn)
    init_emb = actor.init_node_embeddings[:, 0].clone()     # [N,16]
    N = gnn_emb.shape[0]
    init = np.random.RandomState(sseed).randint(0, 2, (N, T), dtype=np.int8)  # tradjectories.py:64 draws int8

    # oracle replay: free-running greedy, must reproduce the reference
    ob = orc.build_batch(graphs)
    model = orc.ModelOracle(sd, compat=True)
    olog = orc.rollout(model, ob, T, S, init, tau=0.0, record=True)
    d_emb = (model.gnn_out - gnn_emb).abs().max().item()
    print(f"model trace: gnn emb max|oracle-ref| = {d_emb:.3e}")
    assert d_emb < 2e-6
    for s in range(S):
        assert torch.equal(olog['actions'][s], rec['actions'][s]), f"actions differ at step {s}"
        assert torch.equal(olog['scores'][s], log['scores'][s]), s
        dq = (olog['q'][s] - rec['q'][s]).abs().max().item()
        dh = (olog['hidden'][s] - rec['hidden'][s]).abs().max().item()
        assert dq < 1e-5 and dh < 1e-5, (s, dq, dh)
    print(f"model trace: {S} greedy steps, oracle actions == reference actions; "
          f"max|dq| = {max((olog['q'][s] - rec['q'][s]).abs().max().item() for s in range(S)):.2e}")

    out = pack_graphs(graphs)
    out.update(dict(
        T=np.int32(T), S=np.int32(S), weight_seed=np.int32(wseed), init_state=init,
        gnn_emb=gnn_emb.numpy(), init_emb=init_emb.numpy(),
        q=torch.stack(rec["q"]).numpy(),                                 # [S,N,T]
        P=torch.stack(rec["P"]).numpy(),                                 # [S,G,T,32]
        hidden_slice=torch.stack(rec["hidden"])[..., ::16].numpy(),      # [S,G,T,64]
        hidden_norm=torch.stack(rec["hidden"]).norm(dim=-1).numpy(),     # [S,G,T]
        last_sel=torch.stack(rec["last_sel"]).numpy(),                   # [S,G,T,32]
        actions=torch.stack(rec["actions"]).numpy().astype(np.int16),    # [S,G,T] local ids
        scores=torch.stack(log["scores"]).numpy(),                       # [S,G,T]
        init_score=log["init_score"].numpy(),
        best_state=log["best_state"].numpy(),
    ))
    np.savez_compressed(os.path.join(GOLD, "model_trace.npz"), **out)


# ------------------------------------------------------------------------------------------------
def gen_config1():
    name = "validation/ER_40spin_p15_100graphs"
    graphs, opts = load(name)
    wseed, sseed, T, S, bsz = 0, 1234, 50, 80, 50
    sd = random_state_dicts(wseed, perturb_norms=False)
    solver, gb = make_solver(sd)
    cfg = TestGraphsConfig(graphs=graphs, opt_scores=torch.tensor(opts), num_steps=S, tradj_per_graph=T,
                           tau=0, max_time=None, graphs_bsz=bsz, tradj_bsz=T)
    np.random.seed(sseed)   # Tradjectories draws init states from numpy's global RNG (tradjectories.py:64)
    log = test_solver(solver, gb, cfg, verbose=False)
    # the same init states, reproduced: one randint call per graph chunk, in order
    rs = np.random.RandomState(sseed)
    inits = [rs.randint(0, 2, (40 * bsz, T), dtype=np.int8) for _ in range(2)]
    model = orc.ModelOracle(sd, compat=True)
    res = orc.test_solver(model, graphs, S, T, bsz, lambda i, N, T_: inits[i], tau=0.0)
    agree = (res['scores_max'] == log['scores_max']).float().mean().item()
    print(f"config1: oracle best-cut == reference best-cut on {agree * 100:.1f}% of 100 graphs; "
          f"apx_max mean ref={float(log['apx_max'].mean()):.4f}")
    assert torch.equal(res['init_score'], log['init_score'])
    assert agree == 1.0
    assert torch.allclose(res['scores_mean'], log['scores_mean'])
    out = pack_graphs(graphs)
    out.update(dict(T=np.int32(T), S=np.int32(S), bsz=np.int32(bsz), weight_seed=np.int32(wseed),
                    state_seed=np.int32(sseed), opts=np.asarray(opts, np.float32),
                    scores_max=log['scores_max'].numpy(), scores_mean=log['scores_mean'].numpy(),
                    scores_by_tradj=log['scores'].max(-1).values.numpy().astype(np.int16),
                    init_score=log['init_score'].numpy().astype(np.int16),
                    final_score=log['scores'][..., -1].numpy().astype(np.int16)))
    np.savez_compressed(os.path.join(GOLD, "config1.npz"), **out)


# ------------------------------------------------------------------------------------------------
def _to_nx(g):
    """EdgeListGraph (ecord_b200.graphs.synthetic_set) -> networkx graph whose reference edge order
    (GraphData.from_networkx: to_directed().edges) is the (src, dst)-sorted order of the edge list."""
    import networkx as nx
    G = nx.Graph()
    G.add_nodes_from(range(g.n))
    keep = g.src < g.dst
    for i, j, w in zip(g.src[keep], g.dst[keep], g.w[keep]):
        G.add_edge(int(i), int(j), weight=float(w))
    return G


def gen_fullsize(which=("er200", "er500", "ba500")):
    """BASELINE configs 2-3 at the real rollout length, through the reference's own test_solver."""
    from ecord_b200.graphs import synthetic_set
    wseed, sseed, KEEP_T = 0, 4242, 8
    sd = random_state_dicts(wseed, perturb_norms=False)
    for name in which:
        T, num_steps = 50, -4
        if name == "er200":
            graphs, _ = load("testing/ER_200spin_p15_50graphs"); graphs = graphs[:10]; out_name = "config2_er200.npz"
        elif name == "ba500":
            graphs, _ = load("testing/BA_500spin_m4_50graphs"); graphs = graphs[:10]; out_name = "config3_ba500.npz"
        elif name == "er2000":
            # BASELINE config 4's graph family (graphs/generated: G(n, avg degree 10), unit weights) at n = 2000: 16 vertex tiles per
            # graph, so the arg-max of a unit spans 16 CTAs of the Q kernel.  2 graphs x 32 trajectories x 1500 steps (the full
            # 4 |V| = 8000 steps x 128 trajectories would take the CPU reference half an hour per graph)
            graphs = [_to_nx(g) for g in synthetic_set("er", 2000, 2, avg_degree=10, signed=False)]; out_name = "config4_er2000.npz"
            T, num_steps = 32, 1500
        elif name == "er2000s":
            # the same family with +-1 weights.  With UNIT weights the t=0 GNN is degenerate -- every vertex starts from the same
            # feature and a mean over identical neighbour messages is that message, so all vertices share one embedding up to
            # rounding -- and 96 % of the reference's arg-max decisions are ties within 1e-6 of the Q spread (26 % exact): no
            # implementation with another summation order follows its trajectories (oracle/ablate_precision.py er2000: an fp64
            # evaluation leaves them after 1.8 steps).  Signed weights break the symmetry, so this fixture can be held to the
            # trajectory bar of configs 2-3 at n = 2000.
            graphs = [_to_nx(g) for g in synthetic_set("er", 2000, 2, avg_degree=10, signed=True)]; out_name = "config4_er2000_signed.npz"
            T, num_steps = 32, 1500
        elif name == "er10000s":
            # the largest size of config 4 (79 vertex tiles per graph), +-1 weights (see er2000s): 1 graph x 16 trajectories x 400 steps
            graphs = [_to_nx(g) for g in synthetic_set("er", 10000, 1, avg_degree=10, signed=True)]; out_name = "config4_er10000_signed.npz"
            T, num_steps = 16, 400
        else:
            graphs = [_to_nx(g) for g in synthetic_set("er", 500, 5, p=0.15, signed=True)]; out_name = "config2_er500.npz"
        solver, gb = make_solver(sd)
        acts = []
        actor = solver.actor
        orig = actor.action_select

        def wrapped(obs, _orig=orig, _acts=acts, **kw):
            ret = _orig(obs, **kw)
            _acts.append(ret[0][:, :KEEP_T].clone())
            return ret

        actor.action_select = wrapped
        cfg = TestGraphsConfig(graphs=graphs, opt_scores=None, num_steps=num_steps, tradj_per_graph=T, tau=0, max_time=None,
                               graphs_bsz=len(graphs), tradj_bsz=T)
        np.random.seed(sseed)
        import time
        t0 = time.time()
        log = test_solver(solver, gb, cfg, verbose=False)
        dt = time.time() - t0
        scores = log['scores']                                   # [G,T,S]
        by_tradj = scores.max(-1)
        G, _, S = scores.shape
        N = sum(g.number_of_nodes() for g in graphs)
        init = np.random.RandomState(sseed).randint(0, 2, (N, T), dtype=np.int8)
        # the init states the reference drew are reproduced by the line above: check through the initial score
        ob = orc.build_batch(graphs)
        env = orc.EnvOracle(ob, T, init)
        assert np.array_equal(env.scores, log['init_score'].numpy()), "init states not reproduced"
        actions = torch.stack(acts).numpy()                      # [S,G,KEEP_T] local ids
        assert actions.shape == (S, G, KEEP_T)
        out = pack_graphs(graphs)
        out.update(dict(T=np.int32(T), S=np.int32(S), weight_seed=np.int32(wseed), state_seed=np.int32(sseed),
                        scores_max=log['scores_max'].numpy(), scores_mean=log['scores_mean'].numpy(),
                        scores_by_tradj=by_tradj.values.numpy().astype(np.int32),
                        best_step_by_tradj=by_tradj.indices.numpy().astype(np.int32),
                        init_score=log['init_score'].numpy().astype(np.int32),
                        final_score=scores[..., -1].numpy().astype(np.int32),
                        actions_first_tradj=actions.astype(np.int16)))
        np.savez_compressed(os.path.join(GOLD, out_name), **out)
        print(f"{out_name}: G={G} T={T} S={S}  reference run {dt:.0f}s ({G * T * S / dt:.0f} units/s)  "
              f"scores_max={log['scores_max'].tolist()}", flush=True)


def gen_train_trace():
    """Training-time data collection: the unmodified reference's DQNSolver.rollout in TRAINING mode
    (solver.train(); solver.py:329-356) with a recording callback -- what Trainer's replay buffer receives every step:
    actor_state (GRU state and cached embedding BEFORE the step), obs (normalised features BEFORE the step), action,
    reward (environment.py:113-141 with intermediate_reward = revisit_reward = 0, the canonical ECORD setting) and done.
    Greedy (use_epsilon=False) so that the actions are reproducible without torch's RNG stream."""
    er40, _ = load("validation/ER_40spin_p15_100graphs")
    ba20, _ = load("validation/BA_20spin_m4_100graphs")
    ba60, _ = load("testing/BA_60spin_m4_50graphs")
    graphs = [er40[30], ba20[7], ba60[4], er40[31]]
    T, wseed, sseed = 3, 11, 777
    sd = random_state_dicts(wseed, perturb_norms=True)
    solver, gb = make_solver(sd)
    solver.train()
    batch = gb(graphs)
    rec = {k: [] for k in ("hidden", "last_sel", "node_feat", "glob_feat", "action", "reward", "done", "t_step")}

    def cb(slv, actor_state, obs, action, reward, done, t_step):
        rec["hidden"].append(actor_state.hidden_embeddings.clone())
        rec["last_sel"].append(actor_state.last_selected_node_embeddings.clone())
        rec["node_feat"].append(obs.node_features.clone())
        rec["glob_feat"].append(obs.glob_features.clone())
        rec["action"].append(action.clone())
        rec["reward"].append(reward.clone())
        rec["done"].append(done.clone())
        rec["t_step"].append(t_step)

    np.random.seed(sseed)
    steps0 = solver.num_env_steps
    log = solver.rollout(batch, num_tradj=T, num_steps=-1, callbacks=[(cb, 1, None)], use_epsilon=False, tau=0)
    S = len(rec["action"])
    N = sum(g.shape[0] if isinstance(g, np.ndarray) else g.number_of_nodes() for g in graphs)
    init = np.random.RandomState(sseed).randint(0, 2, (N, T), dtype=np.int8)
    out = pack_graphs(graphs)
    out.update(dict(T=np.int32(T), S=np.int32(S), weight_seed=np.int32(wseed), init_state=init,
                    num_env_steps=np.int32(solver.num_env_steps - steps0), num_rollouts=np.int32(solver.num_rollouts),
                    hidden_slice=torch.stack(rec["hidden"])[..., ::16].numpy(), last_sel=torch.stack(rec["last_sel"]).numpy(),
                    node_feat=torch.stack(rec["node_feat"]).numpy(), glob_feat=torch.stack(rec["glob_feat"]).numpy(),
                    action=torch.stack(rec["action"]).numpy().astype(np.int16), reward=torch.stack(rec["reward"]).numpy(),
                    done=torch.stack(rec["done"]).numpy(), t_step=np.asarray(rec["t_step"], np.int32),
                    scores=torch.stack(log["scores"]).numpy(), init_score=log["init_score"].numpy()))
    np.savez_compressed(os.path.join(GOLD, "train_trace.npz"), **out)
    print(f"train_trace.npz: G={len(graphs)} T={T} S={S} (num_steps=-1 -> |V| per graph, loop to the max), "
          f"reward sum {float(torch.stack(rec['reward']).sum()):.4f}, done at the end {torch.stack(rec['done'])[-1].sum().item()}")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        gen_train_trace()
    elif len(sys.argv) > 1 and sys.argv[1] == "full":
        gen_fullsize(tuple(sys.argv[2:]) or ("er200", "er500", "ba500"))
    else:
        gen_env_traces()
        gen_model_trace()
        gen_config1()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")
