"""The C-ABI library: loads, exports every symbol the header declares, mirrors the structs, and rejects
bad arguments without touching a GPU (no compute calls here)."""
import ctypes as C
import os
import re

from ecord_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ecord_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecord_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ecord_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in ecord_b200/_lib.py"
    assert sorted(_lib.SYMBOLS) == declared


def test_struct_mirrors_match(lib):
    for which, st in enumerate((_lib.Graph, _lib.Env, _lib.Weights, _lib.Policy, _lib.RolloutDesc)):
        assert lib.ecord_abi_sizeof(which) == C.sizeof(st)
    assert lib.ecord_abi_sizeof(99) == -1


def test_version_and_error_strings(lib):
    assert lib.ecord_version() == 3
    assert lib.ecord_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert lib.ecord_error_string(code).startswith(b"ECORD_E_")


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.ecord_env_init(None, None, None, None) == -1
    g = _lib.Graph(num_graphs=0, num_nodes=0, num_edges=0, n_max=0)
    e = _lib.Env(num_tradj=1)
    assert lib.ecord_env_init(C.byref(g), C.byref(e), None, None) == -2        # sizes checked before pointers
    assert lib.ecord_step_cy(*([None] * 11), 1, 1, 1, 1, 1) == -1
    assert lib.ecord_rollout_plan_create(None, None, None) == -1
    assert lib.ecord_rollout_plan_run(None, 1, None) == -6
    assert lib.ecord_rollout_plan_reset(None) == -6
    assert lib.ecord_rollout_plan_destroy(None) == 0
    assert lib.ecord_gnn_workspace_bytes(0) == 0 and lib.ecord_gnn_workspace_bytes(100) >= 3 * 100 * 16 * 4
    assert lib.ecord_fold_q_host(None, None, None, None, None, None) == -1


def test_fold_q_host_matches_numpy(lib):
    import numpy as np
    import torch
    rs = np.random.RandomState(0)
    q1, enc = rs.randn(64, 64).astype(np.float32), rs.randn(16, 3).astype(np.float32)
    g, b, wo = rs.randn(64).astype(np.float32), rs.randn(64).astype(np.float32), rs.randn(1, 64).astype(np.float32)
    out = torch.empty(_lib.QFOLD_FLOATS)
    p = lambda a: C.c_void_p(a.ctypes.data)
    assert lib.ecord_fold_q_host(p(q1), p(enc), p(g), p(b), p(wo), C.c_void_p(out.data_ptr())) == 0
    Cm = q1[:, 48:64].astype(np.float64) @ enc.astype(np.float64)
    o = out.numpy()
    assert np.allclose(o[:192].reshape(3, 64).T, Cm, atol=1e-5)
    assert np.array_equal(o[192:256], g) and np.array_equal(o[256:320], b) and np.array_equal(o[320:384], wo[0])
    Cc = Cm - Cm.mean(0, keepdims=True)                                   # channel-centred planes (fast Q kernel)
    assert np.allclose(o[384:576].reshape(3, 64).T, Cc, atol=1e-5)
    assert np.allclose(o[576:579], (wo[0] * g) @ Cc, atol=1e-4)
    assert np.isclose(o[579], float(wo[0] @ b), atol=1e-5)
    CC = Cc.T @ Cc                                                        # tensor-core Q kernel: 3x3 Gram matrix, upper triangle
    assert np.allclose(o[580:586], [CC[0, 0], CC[0, 1], CC[0, 2], CC[1, 1], CC[1, 2], CC[2, 2]], atol=1e-5)


def test_sass_is_blackwell_native():
    """tcgen05 / TMA / TMEM mnemonics must be present in the built library (guide: B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    # tcgen05.mma (single CTA and cta_group::2), TMA loads, TMEM load / store, multicast commit, packed fp32 FMA
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2D.2CTA", "LDTM", "STTM", "UTCBAR.2CTA.MULTICAST", "FFMA2",
                     "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout


def test_gru_tile_schedule_covers_every_batch_size_exactly_once(lib):
    """The tcgen05 GRU kernel's tiles (64-unit tiles in full rounds over the CTA pairs, the leftover ones cut to 32 or 16 units;
    the enumeration is the kernel's own host + device function) must cover [rows_pad] x [1024 hidden units] exactly once for
    every batch size, use at most 74 CTA pairs and give every pair at most one tile more than any other per width class."""
    import numpy as np
    for rt in list(range(1, 41)) + [74, 75, 100, 157, 200]:
        rows_pad = 256 * rt
        cap = 16 * rt * 4 + 8
        out = np.zeros(3 * cap, np.int32)
        ncl = C.c_int32(0)
        n = lib.ecord_gru_tile_schedule(rows_pad, out.ctypes.data_as(C.c_void_p), cap, C.byref(ncl))
        assert 0 < n <= cap and 1 <= ncl.value <= 74, (rt, n, ncl.value)
        t = out[:3 * n].reshape(n, 3)
        cover = np.zeros((rt, 64), np.int32)                       # (row tile, 16-unit sub-block)
        for m0, u0, w in t:
            assert m0 % 256 == 0 and 0 <= m0 < rows_pad and w in (16, 32, 64) and u0 % w == 0 and u0 + w <= 1024, (rt, m0, u0, w)
            cover[m0 // 256, u0 // 16:(u0 + w) // 16] += 1
        assert (cover == 1).all(), rt
    assert lib.ecord_gru_tile_schedule(100, None, 0, None) < 0
