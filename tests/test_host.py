"""Host-side logic on CPU: batching, loaders/generators, log merging and chunking, weight ingestion, and the
multi-process trajectory sharding + final reduce over gloo (world_size 2)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

import ecord_oracle as orc
from ecord_b200.data import EdgeListGraph, GraphBatcher, to_edge_list
from ecord_b200.dist import shard_bounds
from ecord_b200.graphs import barabasi_albert, erdos_renyi, load_graph_from_file, synthetic_set
from ecord_b200.network import SHAPES, check_state_dicts, random_state_dicts, state_dicts_from_checkpoint
from ecord_b200.testing import TestGraphsConfig, merge_log
from util import load_gold, unpack_graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batcher_matches_oracle_on_all_input_formats():
    import networkx as nx
    import scipy.sparse as sp
    rs = np.random.RandomState(0)
    a = np.triu((rs.rand(17, 17) < 0.3) * (2.0 * rs.randint(0, 2, (17, 17)) - 1), 1)
    a = a + a.T
    g_nx = nx.from_numpy_array(a)
    ref = orc.build_batch([g_nx, g_nx])
    for fmt in (a, g_nx, sp.csr_matrix(a), to_edge_list(a)):
        b = GraphBatcher()([fmt, fmt])
        assert np.array_equal(b.tg.np_src, ref.edge_index[0]) and np.array_equal(b.tg.np_dst, ref.edge_index[1])
        assert np.array_equal(b.tg.np_w, ref.edge_w)
        assert b.tg.ptr.tolist() == [0, 17, 34] and b.tg.batch.tolist() == [0] * 17 + [1] * 17
        assert torch.equal(b.tg.degree, torch.from_numpy(ref.degree))
        assert b.info.batch_offset.tolist() == [0, 17]
    with pytest.raises(NotImplementedError):
        GraphBatcher(add_laplacian_PEs=True)


def test_batcher_on_golden_graphs_and_empty_graph():
    z = load_gold("model_trace.npz")
    graphs = unpack_graphs(z)
    b = GraphBatcher()(graphs)
    ob = orc.build_batch(graphs)
    assert np.array_equal(b.tg.edge_index.numpy(), ob.edge_index)
    assert b.tg.num_edges == ob.E and b.tg.num_nodes == ob.N
    # a graph without edges is legal input (isolated vertices): CSR rows are empty
    e = EdgeListGraph(3, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))
    b = GraphBatcher()([e, graphs[0]])
    assert b.tg.degree[:3].tolist() == [0, 0, 0] and b.tg.num_nodes == 3 + graphs[0].n


def test_generators_are_seeded_symmetric_and_shaped():
    g = erdos_renyi(500, 0.15, seed=3)
    assert g.n == 500 and abs(len(g.w) / 2 - 0.15 * 500 * 499 / 2) < 600
    assert set(np.unique(g.w)) == {-1.0, 1.0}
    assert np.array_equal(g.src, erdos_renyi(500, 0.15, seed=3).src)
    # symmetric: every (i,j,w) has its (j,i,w)
    fwd = set(zip(g.src.tolist(), g.dst.tolist(), g.w.tolist()))
    assert all((j, i, w) in fwd for i, j, w in list(fwd)[:2000])
    assert np.all(np.diff(g.src) >= 0) and not np.any(g.src == g.dst)
    big = erdos_renyi(10000, 10 / 9999, seed=0, signed=False)
    assert abs(len(big.w) / big.n - 10) < 0.5 and np.all(big.w == 1.0)
    ba = barabasi_albert(500, 4, seed=0)
    assert len(ba.w) // 2 == 4 * (500 - 4) and np.bincount(ba.src).max() > 40      # power-law hub
    assert len(synthetic_set("er", 40, 3, p=0.15)) == 3


def test_loader_reads_reference_style_pickles(tmp_path):
    d = tmp_path / "validation"
    (d / "opts").mkdir(parents=True)
    a = np.zeros((4, 4)); a[0, 1] = a[1, 0] = 1.0; a[2, 3] = a[3, 2] = -1.0
    with open(d / "ER_4spin.pkl", "wb") as f:
        pickle.dump([a, a], f)
    with open(d / "opts" / "cuts_ER_4spin.pkl", "wb") as f:
        pickle.dump([1.0, 1.0], f)
    graphs, opts = load_graph_from_file(str(d / "ER_4spin"), quiet=True)
    assert len(graphs) == 2 and opts == [1.0, 1.0]
    cfg = TestGraphsConfig.from_file(str(d / "ER_4spin"), num_steps=-2, tradj_per_graph=5, graphs_bsz=1, tradj_bsz=2)
    assert cfg.get_num_batches() == (2, 3) and cfg.opt_scores.tolist() == [1.0, 1.0]
    assert "2 graphs" in cfg.get_summary() and cfg.get_label() == "ER_4spin"


def test_gset_mtx_loader_and_batcher(tmp_path):
    """graphs/utils.py:36-52: a GSet .mtx file (+ cuts.pkl beside it) -> one sparse graph the batcher accepts, with
    the reference's edge order (source ascending, neighbours ascending) and both directions of every edge."""
    import scipy.io
    import scipy.sparse as sp
    from ecord_b200.graphs import load_gset_graph
    d = tmp_path / "gset"
    d.mkdir()
    iu = np.array([0, 0, 1, 2, 3]); ju = np.array([2, 4, 3, 4, 4]); w = np.array([1, -1, 1, 1, -1], np.float32)
    scipy.io.mmwrite(str(d / "G2.mtx"), sp.coo_matrix((w, (iu, ju)), shape=(5, 5)))     # upper triangle only, as GSet ships
    with open(d / "cuts.pkl", "wb") as f:
        pickle.dump([11624, 11620], f)
    graphs, cuts = load_gset_graph(str(d / "G2"))
    assert cuts == [11620] and len(graphs) == 1
    graphs2, cuts2 = load_graph_from_file(str(d / "G2.mtx"), quiet=True)           # 'gset' in the path -> same loader
    assert cuts2 == [11620]
    tg = GraphBatcher()(graphs + graphs2).tg
    assert tg.num_nodes == 10 and tg.num_edges == 20
    src, dst, ww = tg.np_src[:10], tg.np_dst[:10], tg.np_w[:10]
    assert np.all(np.diff(src) >= 0)
    assert all(np.all(np.diff(dst[src == v]) > 0) for v in range(5))
    assert sorted(zip(src.tolist(), dst.tolist(), ww.tolist())) == sorted(
        [(int(a), int(b), float(c)) for a, b, c in zip(iu, ju, w)] + [(int(b), int(a), float(c)) for a, b, c in zip(iu, ju, w)])
    np.testing.assert_array_equal(tg.np_src[10:] - 5, src)


def test_dump_logs_roundtrip(tmp_path):
    """experiments/validate.py:244-250 + experiments/utils/system.py:74-106: save_to_disk(fname, log, compressed=True)."""
    from ecord_b200.io import load_from_disk, save_to_disk
    log = {"scores": torch.arange(24.0).view(2, 3, 4), "t_step": np.array([0.1, 0.2]), "best_state": [torch.ones(2, 5, 3)],
           "label": "ER_4spin"}
    for compressed in (True, False):
        out = save_to_disk(str(tmp_path / "ER_4spin_run1"), log, compressed=compressed)
        assert os.path.exists(out) and os.path.splitext(out)[-1] in (".snappy", ".pbz2", ".pkl")
        back = load_from_disk(str(tmp_path / "ER_4spin_run1"), compressed=compressed)
        assert torch.equal(back["scores"], log["scores"]) and back["label"] == "ER_4spin"
        assert torch.equal(back["best_state"][0], log["best_state"][0]) and back["t_step"].tolist() == [0.1, 0.2]


def test_merge_log_semantics():
    """testing.py:120-143: list-of-tensor values stack on a new last dim; trajectory chunks concatenate on dim 1,
    graph chunks on dim 0; float lists add element-wise within a graph chunk and concatenate across graph chunks."""
    mk = lambda G, T, S, v: {"scores": [torch.full((G, T), float(v + s)) for s in range(S)], "t_step": [0.5] * S,
                             "init_score": torch.zeros(G, T)}
    lb = merge_log(mk(2, 3, 4, 0), {}, dim=1)
    lb = merge_log(mk(2, 5, 4, 10), lb, dim=1)
    assert lb["scores"].shape == (2, 8, 4) and lb["init_score"].shape == (2, 8)
    assert lb["t_step"].tolist() == [1.0] * 4
    log = merge_log(lb, {}, dim=0)
    log = merge_log(lb, log, dim=0)
    assert log["scores"].shape == (4, 8, 4) and log["t_step"].shape == (8,)
    assert float(log["scores"].max(-1).values.max()) == 13.0


def test_weight_shapes_init_and_checkpoint_ingestion():
    sd = random_state_dicts(0)
    assert sum(v.numel() for m in sd.values() for v in m.values()) == 3902562        # 3760+272+64+3898466
    assert torch.equal(sd["rnn"]["msg_in.0.weight"], random_state_dicts(0)["rnn"]["msg_in.0.weight"])
    assert float(sd["rnn"]["graph_rnn.0.weight_hh"].abs().max()) <= 1 / 32 + 1e-6      # U(+-1/sqrt(1024))
    assert torch.all(sd["rnn"]["mlp_out.1.weight"] == 1) and not torch.all(
        random_state_dicts(0, perturb_norms=True)["rnn"]["mlp_out.1.weight"] == 1)
    ck = {"actor:checkpoint": {f"{m}:state_dict": s for m, s in sd.items()}, "num_env_steps": 3}
    ck["actor:checkpoint"]["rnn_target:state_dict"] = sd["rnn"]                       # extra keys are ignored
    out = state_dicts_from_checkpoint(ck)
    assert set(out) == set(SHAPES) and torch.equal(out["gnn"]["embed.weight"], sd["gnn"]["embed.weight"])
    bad = {m: dict(s) for m, s in sd.items()}
    bad["rnn"]["graph_rnn.0.weight_hh"] = torch.zeros(3072, 512)
    with pytest.raises(ValueError):
        check_state_dicts(bad)
    del bad["gnn"]["embed.bias"]
    with pytest.raises(KeyError):
        check_state_dicts({**bad, "rnn": sd["rnn"]})


def test_node_classifier_head_ingestion():
    """models/node_classifier.py:8-30 (`heads.<i>.0` = Linear(16 -> 1)) -> stacked (weight [H,16], bias [H]); a module, a
    state_dict or None are accepted, wrong shapes and empty dicts are rejected."""
    from ecord_b200.network import node_classifier_heads
    assert node_classifier_heads(None) is None
    sd = {f"heads.{i}.0.weight": torch.full((1, 16), float(i + 1)) for i in range(3)}
    sd.update({f"heads.{i}.0.bias": torch.tensor([0.5 * i]) for i in range(3)})
    w, b = node_classifier_heads(sd)
    assert w.shape == (3, 16) and b.shape == (3,) and w[2, 5] == 3.0 and b[1] == 0.5

    class Heads(torch.nn.Module):            # the reference's module layout
        def __init__(self):
            super().__init__()
            self.heads = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(16, 1)) for _ in range(2)])
    m = Heads()
    w2, b2 = node_classifier_heads(m)
    assert w2.shape == (2, 16) and torch.equal(w2[1], m.heads[1][0].weight.detach().view(-1))
    with pytest.raises(KeyError):
        node_classifier_heads({})
    with pytest.raises(ValueError):
        node_classifier_heads({"heads.0.0.weight": torch.zeros(1, 8), "heads.0.0.bias": torch.zeros(1)})


def test_shard_bounds_partition():
    for total, world in ((50, 8), (128, 8), (7, 2), (3, 4), (400, 1)):
        cuts = [shard_bounds(total, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


def test_graph_shard_covers_quirk_a_prefix():
    """Graph sharding (dist.graph_shard): own ranges partition the graphs; with quirk A the subset starts with the leading
    graphs that cover vertex indices [0, max n of the own graphs), so that local-as-global indices resolve as in the full batch."""
    from ecord_b200.dist import graph_shard
    for sizes, world in (([500] * 100, 8), ([40, 30, 50, 20, 60, 10], 3), ([10, 100, 20], 3), ([5, 5], 4)):
        owned = []
        ptr = np.concatenate([[0], np.cumsum(sizes)])
        for r in range(world):
            subset, own = graph_shard(sizes, world, r, compat=True)
            owned += own
            if not own:
                assert subset == []
                continue
            k = 0
            while k < len(subset) and subset[k] == k:
                k += 1                                         # leading graphs 0..k-1 are in place
            assert ptr[k] >= max(sizes[g] for g in own)
            assert [g for g in subset if g in own] == own and len(set(subset)) == len(subset)
            assert graph_shard(sizes, world, r, compat=False) == (own, own)
        assert owned == list(range(len(sizes)))


def test_tg_batch_select_roundtrip():
    rs = np.random.RandomState(3)
    graphs = []
    for n in (7, 4, 9, 5):
        a = np.triu((rs.rand(n, n) < 0.5) * rs.choice([-1.0, 1.0], (n, n)), 1)
        graphs.append(a + a.T)
    tg = GraphBatcher()(graphs).tg
    sub = tg.select([0, 2, 3])
    ref = GraphBatcher()([graphs[0], graphs[2], graphs[3]]).tg
    assert np.array_equal(sub.np_src, ref.np_src) and np.array_equal(sub.np_dst, ref.np_dst) and np.array_equal(sub.np_w, ref.np_w)
    assert np.array_equal(sub.np_ptr, ref.np_ptr)
    assert np.array_equal(tg.vertex_index([2, 0]), np.concatenate([np.arange(11, 20), np.arange(0, 7)]))


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from ecord_b200.dist import reduce_best, shard_bounds
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
G, T = 5, 7                                          # ragged split: 4 + 3 trajectories
full = torch.arange(G * T, dtype=torch.float32).view(G, T) * (1 - 2 * (torch.arange(G).view(G, 1) % 2))
lo, hi = shard_bounds(T, 2, rank)
best, by_all = reduce_best(full[:, lo:hi].contiguous(), T)
assert torch.equal(by_all, full), (rank, by_all)
assert torch.equal(best, full.max(-1).values)
T2 = 8                                               # equal split: the fast path (no padding)
full2 = torch.arange(G * T2, dtype=torch.float32).view(G, T2) * (1 - 2 * (torch.arange(G).view(G, 1) % 2))
lo, hi = shard_bounds(T2, 2, rank)
best2, by_all2 = reduce_best(full2[:, lo:hi].contiguous(), T2)
assert torch.equal(by_all2, full2) and torch.equal(best2, full2.max(-1).values)
from ecord_b200.dist import gather_graph_shards
glo, ghi = shard_bounds(G, 2, rank)                  # graph sharding: 3 + 2 graphs
assert torch.equal(gather_graph_shards(full[glo:ghi].contiguous(), G), full)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_trajectory_sharding_reduce_over_gloo_world2(tmp_path):
    """the only collective of the path -- all_reduce(MAX) of the best cut + all_gather of the
    per-trajectory best -- gives the single-process answer with 2 ranks (gloo on CPU; NCCL on GPUs)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_double_q_construction_fails_the_way_validate_py_expects():
    """experiments/validate.py:187-200 first tries DQNSolver(use_double_q_networks=True, ...) and falls back to the single-Q
    actor when that raises: the engine's constructor must raise (before touching CUDA), not silently run one network."""
    from ecord_b200.solver import B200Solver
    sd = random_state_dicts(0) if "random_state_dicts" in globals() else None
    if sd is None:
        from ecord_b200.network import random_state_dicts as rsd
        sd = rsd(0)
    for kw in (dict(use_double_q_networks=True), dict(use_ecodqn=True)):
        with pytest.raises(NotImplementedError):
            B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], **kw)
