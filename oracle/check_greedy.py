"""oracle/check_greedy.py -- pin EnvOracle.greedy_solve against the UNMODIFIED reference greedy_solve
(ecord/solvers/solver.py:38-95).  TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).

On graphs whose peeks never tie (distinct dyadic weights) greedy_solve is deterministic, so the reference and the
restatement must agree exactly on the number of steps, the final scores / spins and the best scores, both for the
plain solve and for init_from_best=True.  Writes tests/golden/greedy_trace.npz (graphs, init spins, the reference's
outputs)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np
import networkx as nx
import torch

import build_ref
assert build_ref.load_step_cy() is not None
import ecord_oracle as orc
from gen_golden import pack_graphs

from ecord.environment import NodeObservation, GlobalObservation, Environment
from ecord.environment.data import GraphBatcher
from ecord.solvers.solver import greedy_solve

NODE = [NodeObservation.STATE, NodeObservation.PEEK_NORM, NodeObservation.STEPS_STATIC_CLIP]
GLOB = [GlobalObservation.SCORE_FROM_BEST_NORM, GlobalObservation.PEEK_MAX_NORM]


def graphs_distinct_weights(sizes, seed):
    rs = np.random.RandomState(seed)
    out = []
    for n in sizes:
        g = nx.erdos_renyi_graph(n, 0.2, seed=int(rs.randint(1 << 30)))
        vals = rs.permutation(4096)[:g.number_of_edges()] + 1          # distinct multiples of 2^-10 in (0, 4]
        for (u, v), k in zip(g.edges(), vals):
            g[u][v]['weight'] = float(k) / 1024.0 * (1 if rs.rand() < 0.5 else -1)
        out.append(g)
    return out


def main():
    graphs = graphs_distinct_weights([30, 47, 64], seed=3)
    T = 6
    batch = GraphBatcher()(graphs)
    N = batch.tg.num_nodes
    init = np.random.RandomState(11).randint(0, 2, (N, T)).astype(np.int8)
    env = Environment(batch.nx, T, batch.tg, node_features=NODE, glob_features=GLOB, init_state=torch.from_numpy(init.copy()).float(),
                      device='cpu')
    torch.manual_seed(0)
    env = greedy_solve(env, init_from_best=False)
    ref_scores = env.tradj.get_scores().numpy().copy()
    ref_best = env.tradj._best_scores.copy() if hasattr(env.tradj, "_best_scores") else None
    ref_state = env.tradj.get_node_features(flat=True)[..., 0].numpy().copy() if hasattr(env.tradj, "get_node_features") else None

    ob = orc.build_batch([orc.graph_to_edges(g) for g in graphs])
    o = orc.EnvOracle(ob, T, init)
    steps, tied = o.greedy_solve()
    assert not tied, "fixture must be tie-free"
    assert np.array_equal(o.scores, ref_scores), (o.scores, ref_scores)
    if ref_best is not None:
        assert np.array_equal(o.best_scores, np.asarray(ref_best, np.float32))
    if ref_state is not None:
        assert np.array_equal(o.node_features[..., 0], ref_state)
    # a few scripted steps away from the optimum, then the post-solve from the best states
    rs = np.random.RandomState(5)
    for _ in range(7):
        a = np.stack([rs.randint(0, n, T) for n in ob.num_nodes_list]) if hasattr(ob, "num_nodes_list") else \
            np.stack([rs.randint(0, int(ob.ptr[g + 1] - ob.ptr[g]), T) for g in range(ob.G)])
        env.step(torch.from_numpy(a), ret_reward=False)
        o.step(a)
    assert np.array_equal(o.scores, env.tradj.get_scores().numpy())
    env = greedy_solve(env, init_from_best=True)
    steps2, tied2 = o.greedy_solve(init_from_best=True)
    assert not tied2
    ref_scores2 = env.tradj.get_scores().numpy().copy()
    assert np.array_equal(o.scores, ref_scores2), (o.scores, ref_scores2)
    assert np.array_equal(o.best_scores, np.asarray(env.tradj._best_scores, np.float32))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "greedy_trace.npz"), T=T, init_state=init,
                        scores_after_greedy=ref_scores, steps=steps, scripted_seed=5, scripted_steps=7,
                        scores_after_post=ref_scores2, steps_post=steps2,
                        best_after_post=np.asarray(env.tradj._best_scores, np.float32), **pack_graphs(graphs))
    print(f"reference greedy_solve == oracle: {steps} steps, then {steps2} steps from the best states; fixture written")


if __name__ == "__main__":
    main()
