"""GPU diagnostics: run every parity check with verbose numbers, one section per process.

    python tools/gpu_diag.py <section> [...]      sections: env model tc rollout perf
    bash tools/gpu_diag.sh                         all sections, each under its own timeout

This is a development aid (prints details where the pytest suite only asserts).  It imports the oracle,
so like tests/ it is test infrastructure and not part of the product path.
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import ctypes as C

import numpy as np
import torch

import ecord_oracle as orc
from ecord_b200 import _lib
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from util import load_gold, unpack_graphs, rel_err


def hdr(s):
    print("\n" + "=" * 100 + "\n" + s + "\n" + "=" * 100, flush=True)


def sec_env():
    lib = _lib.load()
    z = load_gold("env_traces.npz")
    p = lambda a: C.c_void_p(a.ctypes.data)
    hdr("B3 drop-in: ecord_step_cy (host buffers) vs reference traces")
    for name in ("canon", "allfeat", "ragged"):
        graphs = unpack_graphs(z, name + "/")
        ob = orc.build_batch(graphs)
        T = int(z[name + "/T"])
        nf = z[name + "/init_node_features"].copy()
        gf = z[name + "/init_glob_features"].copy()
        sc = z[name + "/init_scores"].copy()
        bsc = sc.copy()
        bst = z[name + "/init_state"].astype(np.float32).copy()
        nidx = z[name + "/node_idx"].astype(np.int32)
        gidx = z[name + "/glob_idx"].astype(np.int32)
        acts = z[name + "/actions"].astype(np.int64)
        bad = 0
        for s in range(acts.shape[0]):
            a = np.ascontiguousarray(acts[s] + ob.ptr[:-1, None])
            rc = lib.ecord_step_cy(p(a), p(sc), p(bsc), p(bst), p(nf), p(gf), p(nidx), p(gidx), p(ob.edges_by_node),
                                   p(ob.edge_w), p(ob.edge_ptr), ob.N, ob.G, T, nf.shape[-1], gf.shape[-1])
            assert rc == 0, rc
            if not np.array_equal(sc, z[name + "/score_log"][s]):
                bad += 1
        print(f"{name}: steps with score mismatch {bad}/{acts.shape[0]}; "
              f"node_features equal={np.array_equal(nf, z[name + '/node_features'])} "
              f"glob equal={np.array_equal(gf, z[name + '/glob_features'])} "
              f"best_scores equal={np.array_equal(bsc, z[name + '/best_scores'])} "
              f"best_states equal={np.array_equal(bst, z[name + '/best_states'])}")

    hdr("SoA env: init/step/export vs oracle")
    sd = random_state_dicts(0)
    w = DeviceWeights(sd)
    for name in ("canon", "ragged"):
        graphs = unpack_graphs(z, name + "/")
        T = int(z[name + "/T"])
        b = GraphBatcher()(graphs)
        dg = DeviceGraph(b.tg)
        acts = z[name + "/actions"].astype(np.int64)
        st = RolloutState(dg, w, T, gemm="fp32", max_steps=acts.shape[0])
        st.env_init(z[name + "/init_state"])
        nf, gf, bs = st.export()
        print(f"{name}: init nf equal={np.array_equal(nf.cpu().numpy()[..., :2], z[name + '/init_node_features'][..., :2])} "
              f"scores equal={np.array_equal(st.scores().cpu().numpy(), z[name + '/init_scores'])}")
        bad = 0
        for s in range(acts.shape[0]):
            st.env_step(acts[s].reshape(-1))
            if not np.array_equal(st.scores().cpu().numpy(), z[name + "/score_log"][s]):
                bad += 1
        nf, gf, bs = st.export()
        nfn, gfn, _ = st.export(normalise=True)
        print(f"{name}: score mismatches {bad}/{acts.shape[0]}; nf equal={np.array_equal(nf.cpu().numpy(), z[name + '/node_features'])} "
              f"glob equal={np.array_equal(gf.cpu().numpy(), z[name + '/glob_features'])} "
              f"best_states equal={np.array_equal(bs.cpu().numpy(), z[name + '/best_states'])} "
              f"best_scores equal={np.array_equal(st.best_score.view(dg.G, T).cpu().numpy(), z[name + '/best_scores'])} "
              f"obs_node equal={np.array_equal(nfn.cpu().numpy(), z[name + '/obs_node'])} "
              f"obs_glob equal={np.array_equal(gfn.cpu().numpy(), z[name + '/obs_glob'])} "
              f"score_log equal={np.array_equal(st.score_log.view(-1, dg.G, T).cpu().numpy(), z[name + '/score_log'])}")

    hdr("SoA env at size: BA500-shaped (power-law rows) and ER200, random actions, vs oracle C")
    for kind, n, G, T, S, kw in (("ba", 500, 6, 8, 300, {}), ("er", 200, 8, 6, 300, {"p": 0.15})):
        graphs = synthetic_set(kind, n, G, **kw)
        b = GraphBatcher()(graphs)
        dg = DeviceGraph(b.tg)
        ob = orc.build_batch([tuple(g) for g in graphs])
        rs = np.random.RandomState(5)
        init = rs.randint(0, 2, (ob.N, T)).astype(np.int8)
        env = orc.EnvOracle(ob, T, init)
        st = RolloutState(dg, w, T, gemm="fp32", max_steps=S)
        st.env_init(init)
        ok0 = np.array_equal(st.scores().cpu().numpy(), env.scores)
        for s in range(S):
            a = rs.randint(0, n, (G, T))
            env.step(a)
            st.env_step(a.reshape(-1))
        nf, gf, bs = st.export()
        print(f"{kind}{n}: init scores equal={ok0}; after {S} steps scores equal={np.array_equal(st.scores().cpu().numpy(), env.scores)} "
              f"nf equal={np.array_equal(nf.cpu().numpy(), env.node_features)} best_states equal={np.array_equal(bs.cpu().numpy(), env.best_states)} "
              f"dmax={int(np.bincount(b.tg.np_src).max())}")


def _model_setup(gemm, dump_q=True, q_mode=None):
    z = load_gold("model_trace.npz")
    graphs = unpack_graphs(z)
    sd = random_state_dicts(int(z["weight_seed"]), perturb_norms=True)
    w = DeviceWeights(sd)
    b = GraphBatcher()(graphs)
    dg = DeviceGraph(b.tg)
    T, S = int(z["T"]), int(z["S"])
    st = RolloutState(dg, w, T, gemm=gemm, compat=True, max_steps=S, dump_q=dump_q, q_mode=q_mode)
    st.gnn_embed()
    st.env_init(z["init_state"])
    return z, st, dg, T, S


def sec_model():
    hdr("GNN embedding vs reference (golden)")
    z, st, dg, T, S = _model_setup("fp32")
    torch.cuda.synchronize()
    print(f"gnn_out rel err {rel_err(st.gnn_out.cpu().numpy(), z['gnn_emb']):.3e}   "
          f"node_emb rel err {rel_err(st.node_emb.cpu().numpy(), z['init_emb']):.3e}")
    hdr("teacher-forced per-step parity, fp32 mode, vs reference (golden)")
    free_ok = True
    for s in range(S):
        st.policy_step()
        st.q_select(forced_actions=z["actions"][s].astype(np.int32).reshape(-1))
        torch.cuda.synchronize()
        q = st.q_values().cpu().numpy()
        h = st.hidden[(st.steps_done + 1) & 1, :st.M].view(dg.G, T, 1024).cpu().numpy()
        P = st.P[:st.M].view(dg.G, T, 32).cpu().numpy()
        ls = st.last_sel[:st.M].view(dg.G, T, 32).cpu().numpy()
        # greedy choice this kernel would have made
        qt = torch.from_numpy(q)
        mine = np.stack([orc._first_argmax(qt[dg.tg.np_ptr[g]:dg.tg.np_ptr[g + 1]]).numpy() for g in range(dg.G)])
        same = (mine == z["actions"][s]).all()
        free_ok &= bool(same)
        print(f"step {s:2d}: q rel {rel_err(q, z['q'][s]):.2e}  hidden rel {rel_err(h[..., ::16], z['hidden_slice'][s]):.2e}  "
              f"|h| rel {rel_err(np.linalg.norm(h, axis=-1), z['hidden_norm'][s]):.2e}  P rel {rel_err(P, z['P'][s]):.2e}  "
              f"last_sel rel {rel_err(ls, z['last_sel'][s]):.2e}  argmax==ref {same}")
        st.env_step()
        assert np.array_equal(st.scores().cpu().numpy(), z["scores"][s]), "scores differ under teacher forcing"
    print("all greedy choices equal to the reference:", free_ok)


def sec_tc():
    hdr("tcgen05 path: bf16 mode vs fp32 mode, teacher-forced with the reference's actions")
    z, st, dg, T, S = _model_setup("bf16")
    z2, st32, _, _, _ = _model_setup("fp32")
    # standalone check of the GRU / PROJ kernels against torch on bf16-rounded operands
    sd = random_state_dicts(int(z["weight_seed"]), perturb_norms=True)
    for s in range(S):
        st.policy_step()
        st32.policy_step()
        torch.cuda.synchronize()
        if s in (0, 1, 5):
            r = sd['rnn']
            dev = st.device
            u = st.u_bf16[:st.M].float()
            hprev = st.hidden_bf16[s & 1, :st.M].float()
            hprev32 = st.hidden[s & 1, :st.M]
            wih = r['graph_rnn.0.weight_ih'].to(dev).bfloat16().float()
            whh = r['graph_rnn.0.weight_hh'].to(dev).bfloat16().float()
            gi = u @ wih.T + r['graph_rnn.0.bias_ih'].to(dev)
            gh = hprev @ whh.T + r['graph_rnn.0.bias_hh'].to(dev)
            rr = torch.sigmoid(gi[:, :1024] + gh[:, :1024])
            zz = torch.sigmoid(gi[:, 1024:2048] + gh[:, 1024:2048])
            nn_ = torch.tanh(gi[:, 2048:] + rr * gh[:, 2048:])
            hnew = (hprev32 - nn_) * zz + nn_
            got = st.hidden[(s + 1) & 1, :st.M]
            y1ref = torch.tanh(got).bfloat16().float() @ r['proj_rnn_to_gnn.1.weight'].to(dev).bfloat16().float().T \
                + r['proj_rnn_to_gnn.1.bias'].to(dev)
            print(f"   step {s}: GRU kernel vs torch(bf16 operands): max abs {float((got - hnew).abs().max()):.3e} "
                  f"(|h|max {float(hnew.abs().max()):.3f});  PROJ kernel vs torch: max abs "
                  f"{float((st.y1[:st.M] - y1ref).abs().max()):.3e} (|y1|max {float(y1ref.abs().max()):.3f});  "
                  f"hidden_bf16 mirror ok {bool(torch.equal(st.hidden_bf16[(s + 1) & 1, :st.M], got.bfloat16()))}  "
                  f"th ok {float((st.th_bf16[:st.M].float() - torch.tanh(got)).abs().max()):.2e}")
        fa = z["actions"][s].astype(np.int32).reshape(-1)
        st.q_select(forced_actions=fa)
        st32.q_select(forced_actions=fa)
        torch.cuda.synchronize()
        q, q32 = st.q_values().cpu().numpy(), st32.q_values().cpu().numpy()
        h = st.hidden[(s + 1) & 1, :st.M].cpu().numpy()
        h32 = st32.hidden[(s + 1) & 1, :st.M].cpu().numpy()
        print(f"step {s:2d}: bf16 vs fp32: hidden rel {rel_err(h, h32):.2e}  y1 rel {rel_err(st.y1[:st.M].cpu().numpy(), st32.y1[:st.M].cpu().numpy()):.2e}  "
              f"P rel {rel_err(st.P[:st.M].cpu().numpy(), st32.P[:st.M].cpu().numpy()):.2e}  q rel {rel_err(q, q32):.2e}  "
              f"q(bf16) vs reference rel {rel_err(q, z['q'][s]):.2e}")
        st.env_step()
        st32.env_step()


def sec_qfast():
    hdr("fast Q kernel vs exact Q kernel (fp32 GEMM for both), teacher-forced")
    z, sa, dg, T, S = _model_setup("fp32", q_mode="exact")
    _, sb, _, _, _ = _model_setup("fp32", q_mode="fast")
    for s in range(S):
        fa = z["actions"][s].astype(np.int32).reshape(-1)
        for st in (sa, sb):
            st.policy_step()
            st.q_select(forced_actions=None)
        torch.cuda.synchronize()
        qa, qb = sa.q_values().cpu().numpy(), sb.q_values().cpu().numpy()
        print(f"step {s:2d}: q fast vs exact rel {rel_err(qb, qa):.2e}; fast vs reference rel {rel_err(qb, z['q'][s]):.2e}; "
              f"argmax equal {bool(torch.equal(sa.actions, sb.actions))}; u equal {bool(torch.equal(sa.u, sb.u))}; "
              f"last_sel rel {rel_err(sb.last_sel.cpu().numpy(), sa.last_sel.cpu().numpy()):.1e}")
        for st in (sa, sb):
            st.q_select(forced_actions=fa)
            st.env_step()
    print("scores equal:", bool(torch.equal(sa.score, sb.score)), " keys all zero:", int(sb.keys.abs().sum()) == 0)


def sec_rollout():
    hdr("free-running greedy rollout through the CUDA-graph plan vs reference (golden model trace)")
    for gemm in ("fp32", "bf16"):
        z, st, dg, T, S = _model_setup(gemm, dump_q=False)
        st.run(S)
        torch.cuda.synchronize()
        acts = st.action_log[:S].view(S, dg.G, T).cpu().numpy()
        scores = st.score_log[:S].view(S, dg.G, T).cpu().numpy()
        same_steps = [(acts[s] == z["actions"][s]).all() for s in range(S)]
        first_div = same_steps.index(False) if False in same_steps else -1
        print(f"{gemm}: actions identical for {sum(same_steps)}/{S} steps (first divergence at step {first_div}); "
              f"trajectories fully identical: {float((acts == z['actions']).all(0).mean()) * 100:.1f}%; "
              f"final scores equal: {np.array_equal(scores[-1], z['scores'][-1])}; launches {st.kernel_launches}")
        if gemm == "fp32":
            bs = st.best_states_batched().numpy()
            print(f"      best_state (compat) equal to reference: {np.array_equal(bs, z['best_state'])}")
            snap = st.best_states_batched(true_snapshot=True).numpy()
            # true snapshot must reproduce best_score as its cut value
            print(f"      true snapshot differs from compat best_state in {int((snap != bs).sum())} entries")

    hdr("BASELINE config 1 (ER40 x100, 80 steps, 50 tradj, batch 50): best-cut agreement with the reference")
    from ecord_b200.solver import B200Solver
    from ecord_b200.testing import TestGraphsConfig, test_solver
    z = load_gold("config1.npz")
    graphs = unpack_graphs(z)
    T, S, bsz = int(z["T"]), int(z["S"]), int(z["bsz"])
    sd = random_state_dicts(int(z["weight_seed"]))
    for gemm in ("fp32", "bf16"):
        solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], gemm=gemm)
        cfg = TestGraphsConfig(graphs=graphs, opt_scores=torch.tensor(z["opts"]), num_steps=S, tradj_per_graph=T, tau=0,
                               graphs_bsz=bsz, tradj_bsz=T)
        np.random.seed(int(z["state_seed"]))
        t0 = time.time()
        log = test_solver(solver, GraphBatcher(), cfg)
        dt = time.time() - t0
        smax = log["scores_max"].numpy()
        by = log["scores"].max(-1).values.numpy()
        print(f"{gemm}: init_score equal {np.array_equal(log['init_score'].numpy(), z['init_score'].astype(np.float32))}; "
              f"best-cut equal on {float((smax == z['scores_max']).mean()) * 100:.1f}% of graphs; "
              f"per-trajectory best equal on {float((by == z['scores_by_tradj']).mean()) * 100:.2f}%; "
              f"final score equal on {float((log['scores'][..., -1].numpy() == z['final_score']).mean()) * 100:.2f}% of trajectories; "
              f"apx_max mean {float(log['apx_max'].mean()):.4f} (ref {float((z['scores_max'] / z['opts']).mean()):.4f}); "
              f"wall {dt:.2f}s t_rollout {float(log['t_rollout'].sum()):.4f}s")


def sec_perf():
    hdr("timing: ER500 (p=0.15, +-1), G=50, T=50")
    sd = random_state_dicts(0)
    w = DeviceWeights(sd)
    graphs = synthetic_set("er", 500, 50, p=0.15)
    b = GraphBatcher()(graphs)
    dg = DeviceGraph(b.tg)
    T = 50
    init = np.random.RandomState(0).randint(0, 2, (dg.N, T)).astype(np.int8)
    for gemm, qm in (("bf16", "fast"), ("bf16", "exact"), ("fp32", "exact")):
        S = 400 if gemm == "bf16" else 48
        st = RolloutState(dg, w, T, gemm=gemm, max_steps=S + 64, q_mode=qm)
        t0 = time.time(); st.gnn_embed(); torch.cuda.synchronize(); t_gnn = time.time() - t0
        t0 = time.time(); st.env_init(init); torch.cuda.synchronize(); t_init = time.time() - t0
        st.run(16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.run(S); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{gemm}/{qm}: {S} steps in {ms:.2f} ms -> {ms / S * 1e3:.1f} us/step, {dg.G * T * S / ms * 1e3 / 1e6:.2f} M units/s "
              f"(gnn {t_gnn * 1e3:.1f} ms, env_init {t_init * 1e3:.1f} ms)")
        # per-kernel coarse timing in step-wise mode
        def tm(fn, reps=20):
            torch.cuda.synchronize()
            a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            bb.record(); bb.synchronize()
            return a.elapsed_time(bb) / reps * 1e3
        print(f"      policy_step {tm(st.policy_step):.1f} us   q_select {tm(st.q_select):.1f} us")
        st.max_steps = 0   # keep env_step from logging while it is timed repeatedly
        sl, st.score_log = st.score_log, None
        print(f"      env_step    {tm(st.env_step):.1f} us")
        st.score_log = sl
        st.close()


SECTIONS = {"env": sec_env, "model": sec_model, "tc": sec_tc, "qfast": sec_qfast, "rollout": sec_rollout, "perf": sec_perf}

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    for name in sys.argv[1:]:
        try:
            SECTIONS[name]()
        except Exception:
            traceback.print_exc()
            print(f"SECTION {name} FAILED", flush=True)
