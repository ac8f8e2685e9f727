"""Best-cut parity at the REAL rollout length (BASELINE configs 2 and 3) against fixtures recorded from the unmodified
reference (oracle/gen_golden.py full -> tests/golden/config2_er200.npz, config2_er500.npz, config3_ba500.npz):
4*|V| greedy steps, 50 trajectories, all graphs of a fixture in one rollout batch, random-init canonical network.

Bar (north_star): the greedy rollout reproduces the reference's best cut on >= 99 % of the instances -- with 5-10
instances per fixture that means on every one of them.  It holds for the two fp32-grade modes: "fp32" (CUDA-core
SGEMM) and "bf16x3" (the default: tcgen05 GEMMs on bf16 hi / lo operand pairs, exact gate functions).  The opt-in
"bf16" fast mode (single bf16 products, MUFU.TANH) is held to a statistical bar only and its agreement is printed.

Beyond the best cut the tests compare what the fixtures hold per trajectory: the best score of every one of the 50
trajectories and the chosen vertex at every step of the first eight.  Trajectories are not bit-reproducible between
two correct fp32 implementations (a different summation order flips an arg-max at a near-tie of the random-init
network's flat Q-values: oracle/ablate_precision.py measures 1.4 % of the ER200 trajectories for an fp64 GEMM), so
the bar there is a fraction, not identity.
"""
import numpy as np
import pytest
import torch

from ecord_b200.data import GraphBatcher
from ecord_b200.network import random_state_dicts
from util import load_gold, unpack_graphs

pytestmark = pytest.mark.gpu

FIXTURES = ["config2_er200.npz", "config2_er500.npz", "config3_ba500.npz"]


def _run(fixture, gemm):
    from ecord_b200.solver import B200Solver
    z = load_gold(fixture)
    graphs = unpack_graphs(z)
    T, S = int(z["T"]), int(z["S"])
    sd = random_state_dicts(int(z["weight_seed"]))
    solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], gemm=gemm)
    batch = GraphBatcher()(graphs)
    N = batch.tg.num_nodes
    init = np.random.RandomState(int(z["state_seed"])).randint(0, 2, (N, T), dtype=np.int8)   # tradjectories.py:64
    log = solver.rollout(batch, num_tradj=T, num_steps=-4, tau=0, use_epsilon=False, init_state=init)
    assert len(log["scores"]) == S
    assert np.array_equal(log["init_score"].numpy(), z["init_score"].astype(np.float32))
    scores = torch.stack(log["scores"], -1).numpy()                       # [G,T,S]
    by_tradj = scores.max(-1)
    st = solver._last_state
    acts = st.action_log[:S].view(S, st.g.G, T)[:, :, :z["actions_first_tradj"].shape[2]].cpu().numpy()
    ref_acts = z["actions_first_tradj"].astype(np.int32)
    same_traj = (acts == ref_acts).all(0)                                 # [G, 8]
    first_div = np.where((acts != ref_acts).any(0), (acts != ref_acts).argmax(0), S)
    res = dict(best_equal=float((by_tradj.max(-1) == z["scores_max"]).mean()),
               tradj_best_equal=float((by_tradj == z["scores_by_tradj"]).mean()),
               tradj_identical=float(same_traj.mean()), first_divergence=float(first_div.mean()),
               mean_best=float(by_tradj.max(-1).mean()), ref_mean_best=float(z["scores_max"].mean()),
               mean_by_tradj=float(by_tradj.mean()), ref_mean_by_tradj=float(z["scores_by_tradj"].mean()))
    print(f"{fixture} {gemm}: " + "  ".join(f"{k}={v:.3f}" for k, v in res.items()))
    return res


@pytest.mark.parametrize("fixture", FIXTURES)
@pytest.mark.parametrize("gemm", ["bf16x3", "fp32"])
def test_full_length_best_cut_equals_reference(fixture, gemm):
    r = _run(fixture, gemm)
    assert r["best_equal"] >= 0.99, r
    assert r["tradj_best_equal"] >= 0.95, r
    assert r["tradj_identical"] >= 0.80, r


@pytest.mark.parametrize("fixture", FIXTURES)
def test_full_length_best_cut_fast_mode_is_statistically_equal(fixture):
    """gemm="bf16" (opt-in): trajectories leave the reference's after a few hundred steps; the cut quality must not."""
    r = _run(fixture, "bf16")
    assert abs(r["mean_best"] - r["ref_mean_best"]) <= 0.01 * r["ref_mean_best"], r
    assert abs(r["mean_by_tradj"] - r["ref_mean_by_tradj"]) <= 0.01 * r["ref_mean_by_tradj"], r
    assert r["best_equal"] >= 0.6, r
