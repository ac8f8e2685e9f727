"""Model kernels on the GPU against the reference's recorded tensors (tests/golden/model_trace.npz).

Tolerances (BASELINE.json north_star): fp32 mode 1e-4 relative; bf16 (tcgen05) mode 2e-2 relative, where
"relative" = max|a-b| / max|b| over the tensor.  Per-step checks are teacher-forced with the reference's
actions so that an argmax flip cannot mask a numerical comparison."""
import os

import numpy as np
import pytest
import torch

import ecord_oracle as orc
from ecord_b200.data import GraphBatcher
from ecord_b200.network import random_state_dicts
from util import load_gold, rel_err, spread_rel_err, unpack_graphs

pytestmark = pytest.mark.gpu
# "bf16x3" (tcgen05 on bf16 hi / lo operand pairs, the default mode) is held to the fp32 tolerance
TOL = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 2e-2}
# error of Q relative to the spread of a unit's Q-values (what decides the arg-max), see util.spread_rel_err
TOL_SPREAD = {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 0.5}
TESTS_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT_DIR = os.path.dirname(TESTS_DIR)


def _setup(gemm, compat=True, dump_q=True, q_mode=None):
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("model_trace.npz")
    sd = random_state_dicts(int(z["weight_seed"]), perturb_norms=True)
    dg = DeviceGraph(GraphBatcher()(unpack_graphs(z)).tg)
    T, S = int(z["T"]), int(z["S"])
    st = RolloutState(dg, DeviceWeights(sd), T, gemm=gemm, compat=compat, max_steps=S, dump_q=dump_q, q_mode=q_mode)
    st.gnn_embed()
    st.env_init(z["init_state"])
    return z, sd, st, dg, T, S


def test_gnn_embedding_matches_reference():
    z, sd, st, dg, T, S = _setup("fp32")
    assert rel_err(st.gnn_out.cpu().numpy(), z["gnn_emb"]) < 1e-4
    assert rel_err(st.node_emb.cpu().numpy(), z["init_emb"]) < 1e-4


def test_gnn_layernorm_is_batch_wide():
    """Quirk B: embeddings of a graph depend on the batch it is embedded with (gnn_embedder.py:178,193-199)."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("model_trace.npz")
    graphs = unpack_graphs(z)
    sd = random_state_dicts(int(z["weight_seed"]), perturb_norms=True)
    w = DeviceWeights(sd)
    model = orc.ModelOracle(sd)
    for sub in (graphs[:1], graphs[:3]):
        dg = DeviceGraph(GraphBatcher()(sub).tg)
        st = RolloutState(dg, w, 1, gemm="fp32")
        st.gnn_embed()
        ref = model.gnn_embed(orc.build_batch(sub)).numpy()
        assert rel_err(st.gnn_out.cpu().numpy(), ref) < 1e-4
    # and the two differ from each other by more than the tolerance, as in the reference
    a = model.gnn_embed(orc.build_batch(graphs[:1])).numpy()
    b = model.gnn_embed(orc.build_batch(graphs[:3])).numpy()[:a.shape[0]]
    assert rel_err(a, b) > 1e-4


@pytest.mark.parametrize("gemm", ["fp32", "bf16x3", "bf16"])
def test_per_step_tensors_teacher_forced(gemm):
    z, sd, st, dg, T, S = _setup(gemm)
    tol = TOL[gemm]
    worst_spread = 0.0
    for s in range(S):
        st.policy_step()
        st.q_select(forced_actions=z["actions"][s].astype(np.int32).reshape(-1))
        q = st.q_values().cpu().numpy()
        h = st.hidden[(s + 1) & 1, :st.M].view(dg.G, T, 1024).cpu().numpy()
        assert rel_err(q, z["q"][s]) < tol, f"q step {s}"
        worst_spread = max(worst_spread, spread_rel_err(q, z["q"][s], dg.tg.np_ptr))
        assert worst_spread < TOL_SPREAD[gemm], f"q step {s}: error / spread of the unit's Q-values = {worst_spread:.3e}"
        assert rel_err(h[..., ::16], z["hidden_slice"][s]) < tol, f"hidden step {s}"
        assert rel_err(np.linalg.norm(h, axis=-1), z["hidden_norm"][s]) < tol
        assert rel_err(st.P[:st.M].view(dg.G, T, 32).cpu().numpy(), z["P"][s]) < tol, f"P step {s}"
        assert rel_err(st.last_sel[:st.M].view(dg.G, T, 32).cpu().numpy(), z["last_sel"][s]) < 1e-5, f"last_sel {s}"
        st.env_step()
        assert np.array_equal(st.scores().cpu().numpy(), z["scores"][s])
    print(f"{gemm}: worst Q error relative to the unit's Q spread over {S} steps: {worst_spread:.3e}")


@pytest.mark.parametrize("compat", [True, False])
def test_fast_q_kernel_matches_exact_kernel(compat):
    """ECORD_Q_FAST is the same function re-associated (centred LayerNorm, LeakyReLU split): Q within 1e-5 relative of
    the literal kernel, identical greedy choices on the reference trace, identical cached embeddings and u."""
    z, sd, a, dg, T, S = _setup("fp32", compat=compat, q_mode="exact")
    _, _, b, _, _, _ = _setup("fp32", compat=compat, q_mode="fast")
    for s in range(S):
        for st in (a, b):
            st.policy_step()
            st.q_select()
        assert rel_err(b.q_values().cpu().numpy(), a.q_values().cpu().numpy()) < 1e-5, s
        assert torch.equal(a.actions, b.actions), s
        assert torch.equal(a.last_sel, b.last_sel) and torch.equal(a.u, b.u), s
        assert int(b.keys.abs().sum()) == 0                    # finish kernel re-arms the arg-max keys
        for st in (a, b):
            st.env_step()
    assert torch.equal(a.score, b.score) and torch.equal(a.peek, b.peek) and torch.equal(a.sp, b.sp)


def test_fast_q_kernel_vertex_tiles_and_trajectory_splits():
    """Graphs larger than one 128-vertex tile, ragged sizes, T not a multiple of the trajectory chunk."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.graphs import synthetic_set
    graphs = synthetic_set("er", 300, 2, p=0.05) + synthetic_set("er", 129, 1, p=0.1, seed0=5) + synthetic_set("er", 40, 2, p=0.2, seed0=9)
    sd = random_state_dicts(21, perturb_norms=True)
    w = DeviceWeights(sd)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    T = 37
    init = np.random.RandomState(3).randint(0, 2, (dg.N, T)).astype(np.int8)
    a = RolloutState(dg, w, T, gemm="fp32", q_mode="exact", dump_q=True)
    b = RolloutState(dg, w, T, gemm="fp32", q_mode="fast", dump_q=True)
    for st in (a, b):
        st.gnn_embed(); st.env_init(init)
    for s in range(12):
        for st in (a, b):
            st.policy_step(); st.q_select()
        assert rel_err(b.q_values().cpu().numpy(), a.q_values().cpu().numpy()) < 1e-5
        # arg-max may legitimately differ only where the two best Q-values are within rounding of each other
        qa = a.q_values()
        diff = (a.actions != b.actions).nonzero().flatten().cpu().tolist()
        for m in diff:
            g_, t_ = divmod(m, T)
            lo = int(dg.tg.np_ptr[g_])
            col = qa[lo:int(dg.tg.np_ptr[g_ + 1]), t_]
            gap = float(col[int(a.actions[m])] - col[int(b.actions[m])])
            assert abs(gap) < 1e-5 * float(col.abs().max())
        b.actions.copy_(a.actions); b.last_sel.copy_(a.last_sel)   # keep the two states in lock-step
        for st in (a, b):
            st.env_step()
    assert torch.equal(a.score, b.score)


TOL_TC = 2e-5   # tf32 hi/lo operand split (3-term products), fp32 accumulation: fp32-grade


def _argmax_diffs_are_near_ties(dg, T, qa, a_actions, b_actions, tol):
    """arg-max may legitimately differ only where the two Q-values are within `tol` (relative) of each other"""
    diff = (a_actions != b_actions).nonzero().flatten().cpu().tolist()
    for m in diff:
        g_, t_ = divmod(m, T)
        lo = int(dg.tg.np_ptr[g_])
        col = qa[lo:int(dg.tg.np_ptr[g_ + 1]), t_]
        gap = float(col[int(a_actions[m])] - col[int(b_actions[m])])
        assert abs(gap) < tol * float(col.abs().max()), (m, gap)
    return len(diff)


@pytest.mark.parametrize("compat", [True, False])
def test_tc_q_kernel_matches_exact_kernel(compat):
    """ECORD_Q_TC (tcgen05 kind::tf32 pre-activations + quadratic-form LayerNorm statistics) vs the literal kernel on
    the reference trace: Q within TOL_TC relative, choices identical except at near-ties."""
    z, sd, a, dg, T, S = _setup("fp32", compat=compat, q_mode="exact")
    _, _, b, _, _, _ = _setup("fp32", compat=compat, q_mode="tc")
    worst = 0.0
    for s in range(S):
        for st in (a, b):
            st.policy_step()
            st.q_select()
        qa, qb = a.q_values(), b.q_values()
        worst = max(worst, rel_err(qb.cpu().numpy(), qa.cpu().numpy()))
        assert worst < TOL_TC, (s, worst)
        _argmax_diffs_are_near_ties(dg, T, qa, a.actions, b.actions, 2 * TOL_TC)
        assert int(b.keys.abs().sum()) == 0
        b.actions.copy_(a.actions); b.last_sel.copy_(a.last_sel); b.u.copy_(a.u)   # lock-step
        if b.u_bf16 is not None:
            b.u_bf16.copy_(a.u_bf16)
        for st in (a, b):
            st.env_step()
    assert torch.equal(a.score, b.score) and torch.equal(a.peek, b.peek) and torch.equal(a.sp, b.sp)


def test_tc_q_kernel_vertex_tiles_and_trajectory_splits():
    """Graphs larger than one 128-vertex tile, ragged sizes, partial tiles, odd trajectory counts."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.graphs import synthetic_set
    graphs = synthetic_set("er", 300, 2, p=0.05) + synthetic_set("er", 129, 1, p=0.1, seed0=5) + synthetic_set("er", 40, 2, p=0.2, seed0=9)
    sd = random_state_dicts(21, perturb_norms=True)
    w = DeviceWeights(sd)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    for T in (37, 1, 130):
        init = np.random.RandomState(3).randint(0, 2, (dg.N, T)).astype(np.int8)
        a = RolloutState(dg, w, T, gemm="fp32", q_mode="exact", dump_q=True)
        b = RolloutState(dg, w, T, gemm="fp32", q_mode="tc", dump_q=True)
        for st in (a, b):
            st.gnn_embed(); st.env_init(init)
        for s in range(8):
            for st in (a, b):
                st.policy_step(); st.q_select()
            qa = a.q_values()
            assert rel_err(b.q_values().cpu().numpy(), qa.cpu().numpy()) < TOL_TC, (T, s)
            _argmax_diffs_are_near_ties(dg, T, qa, a.actions, b.actions, 2 * TOL_TC)
            b.actions.copy_(a.actions); b.last_sel.copy_(a.last_sel)
            for st in (a, b):
                st.env_step()
        assert torch.equal(a.score, b.score)


def test_tcgen05_kernels_against_torch_on_identical_operands():
    """The tensor-core GRU / projection kernels vs torch matmuls fed the SAME bf16-rounded operands: isolates
    descriptor/layout bugs from precision (fp32 accumulation order is the only difference)."""
    z, sd, st, dg, T, S = _setup("bf16")
    r, dev = sd['rnn'], st.device
    wih = r['graph_rnn.0.weight_ih'].to(dev).bfloat16().float()
    whh = r['graph_rnn.0.weight_hh'].to(dev).bfloat16().float()
    w1 = r['proj_rnn_to_gnn.1.weight'].to(dev).bfloat16().float()
    for s in range(4):
        st.policy_step()
        u = st.u_bf16[:st.M].float()
        hb, h32 = st.hidden_bf16[s & 1, :st.M].float(), st.hidden[s & 1, :st.M]
        gi = u @ wih.T + r['graph_rnn.0.bias_ih'].to(dev)
        gh = hb @ whh.T + r['graph_rnn.0.bias_hh'].to(dev)
        rr = torch.sigmoid(gi[:, :1024] + gh[:, :1024])
        zz = torch.sigmoid(gi[:, 1024:2048] + gh[:, 1024:2048])
        nn_ = torch.tanh(gi[:, 2048:] + rr * gh[:, 2048:])
        want = (h32 - nn_) * zz + nn_
        got = st.hidden[(s + 1) & 1, :st.M]
        # the kernel evaluates sigmoid/tanh with MUFU.TANH (rel. error ~2^-11): compare at that level
        assert float((got - want).abs().max()) < 2e-3 * max(1.0, float(want.abs().max()))
        assert torch.equal(st.hidden_bf16[(s + 1) & 1, :st.M], got.bfloat16())
        y1 = torch.tanh(got).bfloat16().float() @ w1.T + r['proj_rnn_to_gnn.1.bias'].to(dev)
        assert float((st.y1[:st.M] - y1).abs().max()) < 2e-3 * float(y1.abs().max())   # tanh rounding to bf16 can differ by 1 ulp
        st.q_select(forced_actions=z["actions"][s].astype(np.int32).reshape(-1))
        st.env_step()


def test_tcgen05_split_kernels_are_fp32_grade():
    """ECORD_GEMM_BF16X3: the GRU / projection kernels on bf16 hi / lo operand pairs against an fp64 evaluation of the same
    cell on the fp32 state -- 1e-5 instead of the 2e-3 of the single-product kernels -- and the hi / lo planes they leave
    for the next step are exactly the split of the fp32 values."""
    z, sd, st, dg, T, S = _setup("bf16x3")
    r, dev = sd['rnn'], st.device
    f64 = lambda k: r[k].to(dev).double()
    wih, whh, w1 = f64('graph_rnn.0.weight_ih'), f64('graph_rnn.0.weight_hh'), f64('proj_rnn_to_gnn.1.weight')
    for s in range(4):
        u = st.u[:st.M].double()
        assert torch.equal(st.u_bf16[0, :st.M], st.u[:st.M].bfloat16())
        assert torch.equal(st.u_bf16[1, :st.M], (st.u[:st.M] - st.u_bf16[0, :st.M].float()).bfloat16())
        h32 = st.hidden[s & 1, :st.M].clone()
        st.policy_step()
        gi = u @ wih.T + f64('graph_rnn.0.bias_ih')
        gh = h32.double() @ whh.T + f64('graph_rnn.0.bias_hh')
        rr = torch.sigmoid(gi[:, :1024] + gh[:, :1024])
        zz = torch.sigmoid(gi[:, 1024:2048] + gh[:, 1024:2048])
        nn_ = torch.tanh(gi[:, 2048:] + rr * gh[:, 2048:])
        want = (h32.double() - nn_) * zz + nn_
        got = st.hidden[(s + 1) & 1, :st.M]
        assert float((got.double() - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max())), s
        hi, lo = st.hidden_bf16[(s + 1) & 1, 0, :st.M], st.hidden_bf16[(s + 1) & 1, 1, :st.M]
        assert torch.equal(hi, got.bfloat16()) and torch.equal(lo, (got - hi.float()).bfloat16())
        th = torch.tanh(got.double())
        assert float((st.th_bf16[0, :st.M].float() + st.th_bf16[1, :st.M].float() - th).abs().max()) < 2e-5
        y1 = th @ w1.T + f64('proj_rnn_to_gnn.1.bias')
        assert float((st.y1[:st.M].double() - y1).abs().max()) < 1e-5 * float(y1.abs().max()), s
        st.q_select(forced_actions=z["actions"][s].astype(np.int32).reshape(-1))
        st.env_step()


def test_quirk_a_switch():
    """compat=False caches the embedding of the vertex actually flipped (global index)."""
    z, sd, st, dg, T, S = _setup("fp32", compat=False)
    ob = orc.build_batch(unpack_graphs(z))
    model = orc.ModelOracle(sd, compat=False)
    log = orc.rollout(model, ob, T, 6, z["init_state"], record=True)
    for s in range(6):
        st.policy_step()
        st.q_select(forced_actions=log["actions"][s].numpy().astype(np.int32).reshape(-1))
        assert rel_err(st.q_values().cpu().numpy(), log["q"][s].numpy()) < 1e-4
        assert rel_err(st.last_sel[:st.M].view(dg.G, T, 32).cpu().numpy(), log["last_sel"][s].numpy()) < 1e-5
        st.env_step()


def test_argmax_ties_resolve_to_lowest_index():
    """All-equal Q-values (zero weights in the Q head) must select vertex 0 everywhere (torch_scatter CPU semantics)."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    z = load_gold("model_trace.npz")
    sd = random_state_dicts(3)
    sd['rnn']['mlp_out.3.weight'].zero_()
    dg = DeviceGraph(GraphBatcher()(unpack_graphs(z)).tg)
    st = RolloutState(dg, DeviceWeights(sd), 4, gemm="fp32")
    st.gnn_embed()
    st.env_init(np.random.RandomState(0).randint(0, 2, (dg.N, 4)).astype(np.int8))
    st.policy_step()
    st.q_select()
    assert int(st.actions.abs().sum()) == 0


def test_tau_sampling_is_distributionally_right():
    """tau > 0 (actors/ecord.py:12-19): exponential-race sampling == softmax(q / tau) sampling.  The reference draws
    from torch's RNG so only the distribution can match: chi-square of empirical counts vs softmax probabilities."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.graphs import synthetic_set
    sd = random_state_dicts(11, perturb_norms=True)
    dg = DeviceGraph(GraphBatcher()(synthetic_set("er", 12, 1, p=0.4)).tg)
    T = 4096
    tau = 0.05
    st = RolloutState(dg, DeviceWeights(sd), T, gemm="fp32", tau=tau, seed=123, dump_q=True)
    st.gnn_embed()
    st.env_init(np.zeros((dg.N, T), np.int8))      # identical state in every trajectory -> identical q
    st.policy_step()
    st.q_select()
    q = st.q_values()[:, 0].double().cpu()
    p = torch.softmax(q / tau, 0).numpy()
    counts = np.bincount(st.actions.cpu().numpy(), minlength=dg.N)
    chi2 = float((((counts - T * p) ** 2) / np.maximum(T * p, 1e-9))[p * T > 5].sum())
    dof = int((p * T > 5).sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    st2 = RolloutState(dg, DeviceWeights(sd), T, gemm="fp32", tau=tau, seed=123, dump_q=True)
    st2.gnn_embed(); st2.env_init(np.zeros((dg.N, T), np.int8)); st2.policy_step(); st2.q_select()
    assert torch.equal(st.actions, st2.actions)     # same seed -> same draw (counter-based Philox)


def test_tau_sampling_in_the_tensor_core_kernel():
    """tau > 0 in the fused throughput path (k_q_tc): Gumbel-max, argmax_n q_n / tau - log(-log U_n), is the exponential race
    of actors/ecord.py:12-19 in one pass.  Checked two ways: chi-square of the empirical counts against softmax(q / tau),
    and -- because both kernels read U_n from the same Philox counter -- vertex-for-vertex agreement with the exact
    kernel's draw on the same state (up to near-ties of the race, where the 1e-6 differences between the two Q
    evaluations decide)."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.graphs import synthetic_set
    sd = random_state_dicts(11, perturb_norms=True)
    w = DeviceWeights(sd)
    dg = DeviceGraph(GraphBatcher()(synthetic_set("er", 12, 1, p=0.4) + synthetic_set("er", 150, 1, p=0.1, seed0=2)).tg)
    T, tau = 4096, 0.05

    def draw(q_mode, seed=123):
        st = RolloutState(dg, w, T, gemm="bf16", tau=tau, seed=seed, dump_q=True, q_mode=q_mode)
        st.gnn_embed()
        st.env_init(np.zeros((dg.N, T), np.int8))      # identical state in every trajectory -> identical q
        st.policy_step()
        st.q_select()
        return st.actions.view(dg.G, T).cpu().numpy(), st.q_values().cpu()

    a_tc, q_tc = draw("tc")
    a_ex, q_ex = draw("exact")
    assert float((q_tc - q_ex).abs().max() / q_ex.abs().max()) < 2e-5
    assert (a_tc == a_ex).mean() > 0.995
    q = q_tc[:12, 0].double()
    pr = torch.softmax(q / tau, 0).numpy()
    counts = np.bincount(a_tc[0], minlength=12)
    chi2 = float((((counts - T * pr) ** 2) / np.maximum(T * pr, 1e-9))[pr * T > 5].sum())
    dof = int((pr * T > 5).sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    assert a_tc[1].min() >= 0 and a_tc[1].max() < 150 and len(np.unique(a_tc[1])) > 3
    assert np.array_equal(a_tc, draw("tc")[0]) and not np.array_equal(a_tc, draw("tc", seed=124)[0])
    # and through the fused loop
    st = RolloutState(dg, w, 8, gemm="bf16", tau=tau, seed=5, max_steps=12)
    st.gnn_embed(); st.env_init(np.zeros((dg.N, 8), np.int8)); st.run(12)
    acts = st.action_log[:12].view(12, dg.G, 8).cpu().numpy()
    assert acts[:, 0].max() < 12 and acts[:, 1].max() < 150 and len(np.unique(acts[:, 1])) > 8


@pytest.mark.parametrize("gemm", ["fp32", "bf16"])
def test_epsilon_greedy_is_distributionally_right(gemm):
    """epsilon-greedy (actors/ecord.py:171-180): with probability epsilon a unit's action is floor(U * n_g), otherwise the
    arg-max.  The reference draws from torch.rand, so the distribution is what can match: the fraction of overridden
    units, uniformity of the random actions over each graph's own vertex range (ragged batch), determinism per seed --
    in the exact selection kernel (fp32 mode) and in the tensor-core path's finish kernel (bf16 mode)."""
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.graphs import synthetic_set
    sd = random_state_dicts(5, perturb_norms=True)
    w = DeviceWeights(sd)
    dg = DeviceGraph(GraphBatcher()(synthetic_set("er", 16, 1, p=0.4) + synthetic_set("er", 10, 1, p=0.4, seed0=3)).tg)
    T = 4096

    def first_actions(eps, seed=7):
        st = RolloutState(dg, w, T, gemm=gemm, seed=seed, epsilon=eps)
        st.gnn_embed()
        st.env_init(np.zeros((dg.N, T), np.int8))
        st.policy_step()
        st.q_select()
        return st.actions.view(dg.G, T).cpu().numpy()

    greedy = first_actions(0.0)
    assert (greedy == greedy[:, :1]).all()                 # identical states -> one arg-max per graph
    full = first_actions(1.0)
    for gi, n in enumerate((16, 10)):
        counts = np.bincount(full[gi], minlength=n)
        assert counts.shape[0] == n                          # never outside the graph's own vertices
        chi2 = float(((counts - T / n) ** 2 / (T / n)).sum())
        assert chi2 < (n - 1) + 6 * np.sqrt(2 * (n - 1)) + 10, (gi, chi2)
    part = first_actions(0.25)
    for gi, n in enumerate((16, 10)):
        changed = float((part[gi] != greedy[gi, 0]).mean())
        expect = 0.25 * (1 - 1 / n)
        assert abs(changed - expect) < 5 * np.sqrt(expect * (1 - expect) / T), (gi, changed, expect)
    assert np.array_equal(part, first_actions(0.25)) and not np.array_equal(part, first_actions(0.25, seed=8))
    with pytest.raises(ValueError):
        RolloutState(dg, w, 2, gemm=gemm, epsilon=1.5)


def test_gemm_variants_agree_in_subprocesses():
    """The GRU kernel with its leftover tiles cut to 64, 32 or 16 hidden units (ECORD_GRU_NARROW = 4 | 2 | 1 sub-blocks; the
    default picks the width per launch) and plain launches (ECORD_PDL=0, no programmatic dependent launch) compute the same
    rollout bit for bit, in both tensor-core modes: the switches are read once per process, so each variant runs in its own
    interpreter and prints a digest of hidden state, scores and actions."""
    import subprocess
    import sys
    code = r'''
import sys, hashlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
graphs = synthetic_set("er", 150, 3, p=0.1)
w = DeviceWeights(random_state_dicts(5, perturb_norms=True))
dg = DeviceGraph(GraphBatcher()(graphs).tg)
T = 70
st = RolloutState(dg, w, T, gemm=sys.argv[1], max_steps=40)
st.gnn_embed(); st.env_init(np.random.RandomState(2).randint(0, 2, (dg.N, T)).astype(np.int8))
st.run(37); torch.cuda.synchronize()
h = hashlib.sha1()
for t in (st.hidden[:, :st.M], st.score, st.action_log[:37], st.peek):
    h.update(t.cpu().numpy().tobytes())
print("DIGEST", h.hexdigest())
''' % (ROOT_DIR, TESTS_DIR)
    for gemm in ("bf16x3", "bf16"):
        digests = {}
        for name, env in (("default", {}), ("w64", {"ECORD_GRU_NARROW": "4"}), ("w32", {"ECORD_GRU_NARROW": "2"}),
                          ("w16", {"ECORD_GRU_NARROW": "1"}), ("no-pdl", {"ECORD_PDL": "0"})):
            e = dict(os.environ); e.update(env)
            out = subprocess.run([sys.executable, "-c", code, gemm], env=e, capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stderr[-2000:]
            digests[name] = [l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0]
        assert len(set(digests.values())) == 1, (gemm, digests)
