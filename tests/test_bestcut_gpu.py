"""Best-cut parity at the REAL rollout length (BASELINE configs 2 and 3) against fixtures recorded from the unmodified
reference (oracle/gen_golden.py full -> tests/golden/config2_er200.npz, config2_er500.npz, config3_ba500.npz):
4*|V| greedy steps, 50 trajectories, all graphs of a fixture in one rollout batch, random-init canonical network.
config4_er2000_signed.npz pins BASELINE config 4's graph family at n = 2000 (16 vertex tiles per graph: a unit's arg-max spans
16 CTAs of the Q kernel): 2 graphs x 32 trajectories x 1500 steps of the unmodified reference, +-1 weights;
config4_er10000_signed.npz the family's largest size (n = 10000, 79 tiles: 1 graph x 16 trajectories x 400 steps).  config4_er2000.npz is
the same with the family's own UNIT weights, where the reference's trajectories cannot be followed by any other implementation
(see test_config4_unit_weights_*): it is held to what CAN be pinned there.

Bar (north_star): the greedy rollout reproduces the reference's best cut on >= 99 % of the instances -- with 5-10
instances per fixture that means on every one of them.  It holds for the two fp32-grade modes: "fp32" (CUDA-core
SGEMM) and "bf16x3" (the default: tcgen05 GEMMs on bf16 hi / lo operand pairs, fp32-grade gate functions).  The opt-in
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

FIXTURES = ["config2_er200.npz", "config2_er500.npz", "config3_ba500.npz", "config4_er2000_signed.npz", "config4_er10000_signed.npz"]


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
    # configs 2-3 were recorded at the full 4 |V| steps (num_steps = -4), config 4 (n = 2000) at a fixed 1500 steps
    num_steps = -4 if S == 4 * int(batch.tg.num_nodes_list.max()) else S
    log = solver.rollout(batch, num_tradj=T, num_steps=num_steps, tau=0, use_epsilon=False, init_state=init)
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
    # whole-trajectory identity: the summation-order floor (an fp64 evaluation of the same network against the fp32 reference,
    # oracle/ablate_precision.py) is 0.97-0.99 for the 800-2000 steps of configs 2-3 and 0.47 for 1500 steps at n = 2000
    # (first divergence after 918 steps on average; this engine: 811), where four times as many vertices compete per arg-max
    assert r["tradj_identical"] >= (0.25 if fixture.startswith("config4") else 0.80), r


@pytest.mark.parametrize("fixture", FIXTURES)
def test_full_length_best_cut_fast_mode_is_statistically_equal(fixture):
    """gemm="bf16" (opt-in): trajectories leave the reference's after a few hundred steps; the cut quality must not."""
    r = _run(fixture, "bf16")
    assert abs(r["mean_best"] - r["ref_mean_best"]) <= 0.01 * r["ref_mean_best"], r
    assert abs(r["mean_by_tradj"] - r["ref_mean_by_tradj"]) <= 0.01 * r["ref_mean_by_tradj"], r
    assert r["best_equal"] >= 0.6, r


def test_config4_unit_weights_reference_choices_are_near_tie_optimal():
    """BASELINE config 4 with its own unit weights (G(2000, avg degree 10), as the shipped generated/ER1000.pkl).  There the
    t=0 GNN is degenerate -- all vertices start from one feature and a mean over identical messages is that message, so every
    vertex has the same embedding up to rounding -- and 96 % of the reference's arg-max decisions are ties within 1e-6 of the
    unit's Q spread, 26 % of them exact (oracle/ablate_precision.py er2000): an fp64 evaluation of the same network leaves
    the reference's trajectories after 1.8 steps, so neither can this engine.  What CAN be pinned is pinned: driven along the
    reference's recorded trajectory (its chosen vertex at every one of the 1500 steps, first eight trajectories), the
    environment reproduces the reference's best and final scores exactly, and at every step the reference's choice is optimal
    under the engine's own Q-values to within 1e-4 of the unit's Q spread (fp32-grade modes)."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("config4_er2000.npz")
    graphs = unpack_graphs(z)
    S = int(z["S"])
    ref_acts = z["actions_first_tradj"].astype(np.int32)                   # [S, G, 8]
    T = ref_acts.shape[2]
    sd = random_state_dicts(int(z["weight_seed"]))
    batch = GraphBatcher()(graphs)
    N = batch.tg.num_nodes
    init = np.random.RandomState(int(z["state_seed"])).randint(0, 2, (N, int(z["T"])), dtype=np.int8)[:, :T].copy()
    dg = DeviceGraph(batch.tg)
    ptr = batch.tg.np_ptr
    for gemm in ("bf16x3", "fp32"):
        st = RolloutState(dg, DeviceWeights(sd), T, gemm=gemm, max_steps=S, dump_q=True)
        st.gnn_embed(); st.env_init(init)
        worst, best = 0.0, None
        for s_ in range(S):
            st.policy_step()
            st.q_select(forced_actions=ref_acts[s_].reshape(-1))
            if s_ % 10 == 0 or s_ < 50:                                    # the Q check on a subset of the steps (each is a D2H of N x T)
                q = st.q_values()                                          # [N, T]
                for g in range(dg.G):
                    qq = q[ptr[g]:ptr[g + 1]]
                    chosen = qq.gather(0, torch.as_tensor(ref_acts[s_, g], device=q.device).long().view(1, -1))[0]
                    gap = (qq.max(0).values - chosen) / (qq.max(0).values - qq.min(0).values)
                    worst = max(worst, float(gap.max()))
            st.env_step()
            sc = st.scores()
            best = sc.clone() if best is None else torch.maximum(best, sc)
        print(f"config4_er2000.npz {gemm}: reference's choice below the engine's max Q by at most {worst:.2e} of the unit's Q spread")
        assert worst <= 1e-4, (gemm, worst)
        assert np.array_equal(st.scores().cpu().numpy(), z["final_score"][:, :T].astype(np.float32))
        assert np.array_equal(best.cpu().numpy(), z["scores_by_tradj"][:, :T].astype(np.float32))


def test_config4_unit_weights_free_running_is_statistically_equal():
    """... and free running, the cut quality is the reference's: mean best cut and mean over the trajectories within 0.5 %."""
    r = _run("config4_er2000.npz", "bf16x3")
    assert abs(r["mean_best"] - r["ref_mean_best"]) <= 0.005 * r["ref_mean_best"], r
    assert abs(r["mean_by_tradj"] - r["ref_mean_by_tradj"]) <= 0.005 * r["ref_mean_by_tradj"], r
