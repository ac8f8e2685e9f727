"""Seam B1 proven with the reference's OWN driver: the unmodified `ecord.validate.testing.test_solver` and
`ecord.environment.data.GraphBatcher` (installed under baseline/_ref by baseline/install_ref.py; git-ignored, shipped to
the GPU box) run BASELINE config 1 on top of `ecord_b200.solver.B200Solver`, and the log they produce -- merge_log over
two graph chunks (testing.py:120-143), the keys experiments/validate.py:252-260 reads -- is compared with the fixture the
same driver produced on top of the reference's DQNSolver (tests/golden/config1.npz).

The reference imports torch_scatter / torch_sparse / torch_geometric at module level; oracle/shims/ stands in for them
(the validate path touches: Batch.from_data_list, degree, scatter_max on the degrees).  Only the checker side uses them:
B200Solver itself never imports torch_geometric."""
import os
import sys

import numpy as np
import pytest
import torch

from ecord_b200.network import random_state_dicts
from util import load_gold

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
HAVE_REF = os.path.exists(os.path.join(REF, "ecord", "validate", "testing.py"))


@pytest.fixture(scope="module")
def ref_modules():
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items()
                                               if k == "ecord" or k.startswith(("ecord.", "torch_scatter", "torch_sparse", "torch_geometric"))}
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    for k in list(saved_mods):
        del sys.modules[k]
    try:
        from ecord.environment.data import GraphBatcher
        from ecord.validate.testing import TestGraphsConfig, test_solver
        assert os.path.realpath(sys.modules["ecord"].__file__).startswith(os.path.realpath(REF))
        yield GraphBatcher, TestGraphsConfig, test_solver
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "ecord" or k.startswith(("ecord.", "torch_scatter", "torch_sparse", "torch_geometric"))]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def _nx_graphs(z):
    """The fixture's graphs as the dense adjacency matrices the reference's loader yields for this set (graphs/utils.py:18-23)."""
    import networkx as nx
    out, o = [], 0
    for n, e in zip(z["g_n"], z["g_ne"]):
        a = np.zeros((int(n), int(n)))
        a[z["g_src"][o:o + e], z["g_dst"][o:o + e]] = z["g_w"][o:o + e]
        out.append(nx.from_numpy_array(a))
        o += e
    return out


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed (python baseline/install_ref.py in the build container)")
@pytest.mark.parametrize("gemm", ["bf16x3", "fp32"])
def test_reference_test_solver_drives_b200solver(ref_modules, gemm):
    GraphBatcher, TestGraphsConfig, test_solver = ref_modules
    from ecord_b200.solver import B200Solver
    z = load_gold("config1.npz")
    graphs = _nx_graphs(z)
    T, S, bsz = int(z["T"]), int(z["S"]), int(z["bsz"])
    sd = random_state_dicts(int(z["weight_seed"]))
    solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], gemm=gemm)
    cfg = TestGraphsConfig(graphs=graphs, opt_scores=torch.tensor(z["opts"]), num_steps=S, tradj_per_graph=T, tau=0,
                           max_time=None, graphs_bsz=bsz, tradj_bsz=T)
    np.random.seed(int(z["state_seed"]))        # the solver draws the initial spins as Tradjectories does (tradjectories.py:64)
    log = test_solver(solver, GraphBatcher(), cfg, verbose=False)
    G = len(graphs)
    # values: identical initial scores, best cut equal on >= 99 % of the graphs, means within rounding of the trajectories
    assert np.array_equal(log["init_score"].numpy(), z["init_score"].astype(np.float32))
    agree = float((log["scores_max"].numpy() == z["scores_max"]).mean())
    assert agree >= 0.99, agree
    assert np.abs(log["scores_mean"].numpy() - z["scores_mean"]).max() <= 0.5
    # the log contract (merge_log stacks per-step lists on a new last dim, concatenates graph chunks on dim 0)
    assert log["scores"].shape == (G, T, S) and log["scores0"].shape == (G, T, S)
    assert log["init_score"].shape == (G, T) and log["t_opt"].shape == (G, T)
    assert log["best_state"].shape == (G, 40, T)
    assert log["t_step"].shape == (2 * S,) and log["t_inf"].shape == (2 * S,)
    for k in ("t_setup", "t_init_emb", "t_rollout", "t_tot"):
        assert torch.is_tensor(log[k]) and log[k].numel() == 2, k          # one entry per graph chunk
    assert log["num_graphs"] == G and log["num_tradj"] == T
    assert np.allclose(log["apx_max"].numpy(), log["scores_max"].numpy() / z["opts"])
    assert np.array_equal(log["scores"][..., -1].numpy(), log["scores"].numpy()[..., -1])
    assert float(log["tot_solver_time"]) > 0
    assert solver.is_training is False
