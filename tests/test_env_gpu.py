"""Environment kernels on the GPU against the reference traces and the oracle: bit-exact (integer/fp32 state)."""
import ctypes as C

import numpy as np
import pytest
import torch

import ecord_oracle as orc
from ecord_b200.data import GraphBatcher
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from util import load_gold, unpack_graphs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weights():
    from ecord_b200.engine import DeviceWeights
    return DeviceWeights(random_state_dicts(0))


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("name", ["canon", "allfeat", "ragged"])
def test_step_cy_dropin_is_bit_exact(lib, name):
    """Seam B3: ecord_step_cy with the reference's own arguments (host numpy buffers, full feature tables)."""
    z = load_gold("env_traces.npz")
    ob = orc.build_batch(unpack_graphs(z, name + "/"))
    T = int(z[name + "/T"])
    nf, gf = z[name + "/init_node_features"].copy(), z[name + "/init_glob_features"].copy()
    sc = z[name + "/init_scores"].copy()
    bsc, bst = sc.copy(), z[name + "/init_state"].astype(np.float32).copy()
    nidx, gidx = z[name + "/node_idx"].astype(np.int32), z[name + "/glob_idx"].astype(np.int32)
    acts = z[name + "/actions"].astype(np.int64)
    for s in range(acts.shape[0]):
        a = np.ascontiguousarray(acts[s] + ob.ptr[:-1, None])
        rc = lib.ecord_step_cy(_p(a), _p(sc), _p(bsc), _p(bst), _p(nf), _p(gf), _p(nidx), _p(gidx), _p(ob.edges_by_node),
                               _p(ob.edge_w), _p(ob.edge_ptr), ob.N, ob.G, T, nf.shape[-1], gf.shape[-1])
        assert rc == 0
        assert np.array_equal(sc, z[name + "/score_log"][s]), s
    assert np.array_equal(nf, z[name + "/node_features"])
    assert np.array_equal(gf, z[name + "/glob_features"])
    assert np.array_equal(bsc, z[name + "/best_scores"])
    assert np.array_equal(bst, z[name + "/best_states"])


@pytest.mark.parametrize("name", ["canon", "allfeat", "ragged"])
def test_step_cy_cached_arena_is_bit_exact(lib, name):
    """Seam B3 with the state kept on the device (ecord_step_cy_ctx_*): graph and arrays uploaded once, per step only the
    actions go up and the scores come back; the arrays downloaded at the end equal the reference's, bit for bit."""
    import ctypes as C
    import time
    z = load_gold("env_traces.npz")
    ob = orc.build_batch(unpack_graphs(z, name + "/"))
    T = int(z[name + "/T"])
    nf, gf = z[name + "/init_node_features"].copy(), z[name + "/init_glob_features"].copy()
    sc = z[name + "/init_scores"].copy()
    bsc, bst = sc.copy(), z[name + "/init_state"].astype(np.float32).copy()
    nidx, gidx = z[name + "/node_idx"].astype(np.int32), z[name + "/glob_idx"].astype(np.int32)
    acts = z[name + "/actions"].astype(np.int64)
    ctx = C.c_void_p()
    assert lib.ecord_step_cy_ctx_create(ob.N, ob.G, T, nf.shape[-1], gf.shape[-1], _p(ob.edges_by_node), _p(ob.edge_w),
                                        _p(ob.edge_ptr), C.byref(ctx)) == 0
    assert lib.ecord_step_cy_ctx_upload(ctx, _p(sc), _p(bsc), _p(bst), _p(nf), _p(gf)) == 0
    out = np.empty_like(sc)
    t0 = time.perf_counter()
    for s in range(acts.shape[0]):
        a = np.ascontiguousarray(acts[s] + ob.ptr[:-1, None])
        assert lib.ecord_step_cy_ctx_step(ctx, _p(a), _p(nidx), _p(gidx), _p(out)) == 0
        assert np.array_equal(out, z[name + "/score_log"][s]), s
    dt = (time.perf_counter() - t0) / acts.shape[0]
    assert lib.ecord_step_cy_ctx_download(ctx, _p(sc), _p(bsc), _p(bst), _p(nf), _p(gf)) == 0
    assert lib.ecord_step_cy_ctx_destroy(ctx) == 0
    assert np.array_equal(nf, z[name + "/node_features"]) and np.array_equal(gf, z[name + "/glob_features"])
    assert np.array_equal(bsc, z[name + "/best_scores"]) and np.array_equal(bst, z[name + "/best_states"])
    assert np.array_equal(sc, z[name + "/score_log"][-1])
    print(f"{name}: {dt * 1e6:.0f} us per cached-arena step (host call incl. actions H2D, scores D2H)")


def test_step_cy_rejects_bad_tables(lib):
    a = np.zeros((1, 1), np.int64)
    f = np.zeros(4, np.float32)
    ptr_ = np.zeros(2, np.int32)
    bad = np.array([-1, 1, -1, -1, -1], np.int32)       # STATE not configured (tradjectories.py:226)
    g = -np.ones(6, np.int32)
    assert lib.ecord_step_cy(_p(a), _p(f), _p(f), _p(f), _p(f), _p(f), _p(bad), _p(g), _p(ptr_), _p(f), _p(ptr_),
                             1, 1, 1, 2, 1) == -4


@pytest.mark.parametrize("name", ["canon", "ragged"])
def test_soa_env_matches_reference_trace(weights, name):
    from ecord_b200.engine import DeviceGraph, RolloutState
    z = load_gold("env_traces.npz")
    graphs = unpack_graphs(z, name + "/")
    T = int(z[name + "/T"])
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    acts = z[name + "/actions"].astype(np.int64)
    st = RolloutState(dg, weights, T, gemm="fp32", max_steps=acts.shape[0])
    st.env_init(z[name + "/init_state"])
    nf, gf, bs = st.export()
    assert np.array_equal(nf.cpu().numpy(), z[name + "/init_node_features"])
    assert np.array_equal(st.scores().cpu().numpy(), z[name + "/init_scores"])
    for s in range(acts.shape[0]):
        st.env_step(acts[s].reshape(-1))
    nf, gf, bs = st.export()
    assert np.array_equal(st.score_log.view(-1, dg.G, T).cpu().numpy(), z[name + "/score_log"])
    assert np.array_equal(nf.cpu().numpy(), z[name + "/node_features"])
    assert np.array_equal(gf.cpu().numpy(), z[name + "/glob_features"])
    assert np.array_equal(bs.cpu().numpy(), z[name + "/best_states"])
    assert np.array_equal(st.best_score.view(dg.G, T).cpu().numpy(), z[name + "/best_scores"])
    nfn, gfn, _ = st.export(normalise=True)          # a4: the canonical observation
    assert np.array_equal(nfn.cpu().numpy(), z[name + "/obs_node"])
    assert np.array_equal(gfn.cpu().numpy(), z[name + "/obs_glob"])


@pytest.mark.parametrize("kind,n,G,T,S,kw", [
    ("ba", 500, 8, 10, 400, {}),                       # power-law rows (dmax ~100): lanes stride the CSR row
    ("er", 200, 8, 6, 300, {"p": 0.15}),
    ("er", 2000, 2, 4, 200, {"avg_degree": 10, "signed": False}),   # unit weights, sparse
    ("er", 33, 5, 1, 100, {"p": 0.3}),                 # T = 1, odd sizes
])
def test_soa_env_matches_oracle_at_size(weights, kind, n, G, T, S, kw):
    from ecord_b200.engine import DeviceGraph, RolloutState
    graphs = synthetic_set(kind, n, G, **kw)
    dg = DeviceGraph(GraphBatcher()(graphs).tg)
    ob = orc.build_batch(graphs)
    rs = np.random.RandomState(n)
    init = rs.randint(0, 2, (ob.N, T)).astype(np.int8)
    env = orc.EnvOracle(ob, T, init)
    st = RolloutState(dg, weights, T, gemm="fp32", max_steps=S)
    st.env_init(init)
    assert np.array_equal(st.scores().cpu().numpy(), env.scores)
    for s in range(S):
        a = rs.randint(0, n, (G, T))
        env.step(a)
        st.env_step(a.reshape(-1))
    nf, gf, bs = st.export()
    assert np.array_equal(st.scores().cpu().numpy(), env.scores)
    assert np.array_equal(nf.cpu().numpy(), env.node_features)
    assert np.array_equal(bs.cpu().numpy(), env.best_states)
    assert np.array_equal(st.best_score.view(G, T).cpu().numpy(), env.best_scores)
    # size-independent invariants (SURVEY.md section 4 ii): recompute from the final spins
    fresh = orc.EnvOracle(ob, T, nf.cpu().numpy()[..., 0].astype(np.int8))
    assert np.array_equal(fresh.scores, st.scores().cpu().numpy())
    assert np.array_equal(fresh.node_features[..., 1], nf.cpu().numpy()[..., 1])
    assert (st.best_score >= st.score).all()


def test_env_init_rejects_wrong_shape(weights):
    from ecord_b200.engine import DeviceGraph, RolloutState
    dg = DeviceGraph(GraphBatcher()(synthetic_set("er", 20, 2, p=0.3)).tg)
    st = RolloutState(dg, weights, 3, gemm="fp32")
    with pytest.raises(AssertionError):
        st.env_init(np.zeros((dg.N, 4), np.int8))
