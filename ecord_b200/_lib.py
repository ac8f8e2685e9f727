"""ctypes binding of libecord_b200.so (C ABI in include/ecord_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
The library is built in-tree by `python -m ecord_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libecord_b200.so")

GEMM_FP32 = 0
GEMM_BF16 = 1
GEMM_BF16X3 = 2
Q_EXACT = 0
Q_FAST = 1
Q_TC = 2
QTC_FLOATS = 10160
QA_STRIDE = 92
QTILE = 128
QFOLD_FLOATS = 588
GRU_PACK_ELEMS = 2 * (3072 * 1024 + 4096 * 64)
P1_PACK_ELEMS = 2 * (512 * 1024 + 32 * 512)

_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class EcordError(RuntimeError):
    pass


class Graph(C.Structure):
    _fields_ = [("num_graphs", _i32), ("num_nodes", _i32), ("num_edges", _i32), ("n_max", _i32),
                ("graph_ptr", _vp), ("node_graph", _vp), ("row_ptr", _vp), ("col", _vp), ("w", _vp),
                ("graph_edge_ptr", _vp), ("in_ptr", _vp), ("in_src", _vp), ("in_w", _vp),
                ("num_qtiles", _i32), ("int_weights", _i32), ("qtile_graph", _vp), ("qtile_v0", _vp), ("graph_wsum", _vp)]


class Env(C.Structure):
    _fields_ = [("num_tradj", _i32), ("true_snapshot", _i32), ("peek", _vp), ("sp", _vp), ("best_state", _vp),
                ("score", _vp), ("best_score", _vp), ("best_step", _vp)]


WEIGHT_FIELDS = [
    # (struct field, module, state_dict key)
    ("gnn_embed_w", "gnn", "embed.weight"), ("gnn_embed_b", "gnn", "embed.bias"),
    ("gnn_snd_w", "gnn", "node_model.msg_snd.0.weight"), ("gnn_snd_b", "gnn", "node_model.msg_snd.0.bias"),
    ("gnn_rec_w", "gnn", "node_model.msg_rec.0.weight"), ("gnn_rec_b", "gnn", "node_model.msg_rec.0.bias"),
    ("gnn_wih", "gnn", "node_model.rnn.weight_ih"), ("gnn_whh", "gnn", "node_model.rnn.weight_hh"),
    ("gnn_bih", "gnn", "node_model.rnn.bias_ih"), ("gnn_bhh", "gnn", "node_model.rnn.bias_hh"),
    ("g2r_w", "gnn2rnn", "node_projection.0.weight"), ("g2r_b", "gnn2rnn", "node_projection.0.bias"),
    ("enc_w", "node_encoder", "0.weight"), ("enc_b", "node_encoder", "0.bias"),
    ("msg_w", "rnn", "msg_in.0.weight"), ("msg_b", "rnn", "msg_in.0.bias"),
    ("gru_wih", "rnn", "graph_rnn.0.weight_ih"), ("gru_whh", "rnn", "graph_rnn.0.weight_hh"),
    ("gru_bih", "rnn", "graph_rnn.0.bias_ih"), ("gru_bhh", "rnn", "graph_rnn.0.bias_hh"),
    ("p1_w", "rnn", "proj_rnn_to_gnn.1.weight"), ("p1_b", "rnn", "proj_rnn_to_gnn.1.bias"),
    ("ln1_w", "rnn", "proj_rnn_to_gnn.2.weight"), ("ln1_b", "rnn", "proj_rnn_to_gnn.2.bias"),
    ("p2_w", "rnn", "proj_rnn_to_gnn.4.weight"), ("p2_b", "rnn", "proj_rnn_to_gnn.4.bias"),
    ("ln2_w", "rnn", "proj_rnn_to_gnn.5.weight"), ("ln2_b", "rnn", "proj_rnn_to_gnn.5.bias"),
    ("v1_w", "rnn", "val_out.1.weight"), ("v1_b", "rnn", "val_out.1.bias"),
    ("v2_w", "rnn", "val_out.3.weight"), ("v2_b", "rnn", "val_out.3.bias"),
    ("q1_w", "rnn", "mlp_out.0.weight"), ("q1_b", "rnn", "mlp_out.0.bias"),
    ("qln_w", "rnn", "mlp_out.1.weight"), ("qln_b", "rnn", "mlp_out.1.bias"),
    ("q2_w", "rnn", "mlp_out.3.weight"), ("q2_b", "rnn", "mlp_out.3.bias"),
]


class Weights(C.Structure):
    _fields_ = [(f, _vp) for f, _, _ in WEIGHT_FIELDS] + [("gru_w_bf16", _vp), ("p1_w_bf16", _vp), ("q_tc", _vp), ("q_fold", _vp)]


class Policy(C.Structure):
    _fields_ = [("rows", _i32), ("rows_pad", _i32), ("hidden", _vp), ("hidden_bf16", _vp), ("th_bf16", _vp),
                ("u_bf16", _vp), ("u", _vp), ("gates", _vp), ("th", _vp), ("y1", _vp), ("P", _vp), ("A", _vp),
                ("v", _vp), ("last_sel", _vp), ("node_emb", _vp), ("node_B", _vp), ("actions", _vp), ("q", _vp),
                ("Ac", _vp), ("LA", _vp), ("node_Bc", _vp), ("node_LB", _vp), ("keys", _vp), ("qa", _vp), ("node_qv", _vp),
                ("split", _i32), ("_pad", _i32)]


class RolloutDesc(C.Structure):
    _fields_ = [("graph", C.POINTER(Graph)), ("env", C.POINTER(Env)), ("weights", C.POINTER(Weights)),
                ("policy", C.POINTER(Policy)), ("gemm_mode", _i32), ("q_mode", _i32), ("compat", _i32), ("epsilon", _f32),
                ("tau", _f32),
                ("steps_per_graph", _i32), ("seed", _u64), ("step_counter", _vp), ("score_log", _vp),
                ("action_log", _vp), ("max_steps", _i32), ("_pad", _i32), ("step_time_ns", _vp)]


# every symbol include/ecord_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "ecord_version": (_i32, []),
    "ecord_error_string": (C.c_char_p, [_i32]),
    "ecord_abi_sizeof": (_i64, [_i32]),
    "ecord_gru_tile_schedule": (_i32, [_i32, _vp, _i32, _vp]),
    "ecord_step_cy": (_i32, [_vp] * 11 + [_i32] * 5),
    "ecord_step_cy_device": (_i32, [_vp] * 11 + [_i32] * 6 + [_vp]),
    "ecord_step_cy_ctx_create": (_i32, [_i32] * 5 + [_vp] * 3 + [_P(_vp)]),
    "ecord_step_cy_ctx_upload": (_i32, [_vp] * 6),
    "ecord_step_cy_ctx_step": (_i32, [_vp] * 5),
    "ecord_step_cy_ctx_download": (_i32, [_vp] * 6),
    "ecord_step_cy_ctx_destroy": (_i32, [_vp]),
    "ecord_env_init": (_i32, [_P(Graph), _P(Env), _vp, _vp]),
    "ecord_env_step": (_i32, [_P(Graph), _P(Env), _vp, _i32, _vp, _vp]),
    "ecord_env_step_fused": (_i32, [_P(Graph), _P(Env), _P(Weights), _P(Policy), _i32, _vp, _vp]),
    "ecord_greedy_select": (_i32, [_P(Graph), _P(Env), _f32, _u64, _i32, _vp, _vp, _vp]),
    "ecord_env_reset_state": (_i32, [_P(Graph), _P(Env), _vp, _i32, _vp]),
    "ecord_env_rebase_clock": (_i32, [_P(Graph), _P(Env), _i32, _vp]),
    "ecord_env_export": (_i32, [_P(Graph), _P(Env), _i32, _i32, _vp, _vp, _vp, _vp]),
    "ecord_gnn_workspace_bytes": (C.c_size_t, [_i32]),
    "ecord_gnn_embed": (_i32, [_P(Graph), _P(Weights), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ecord_pack_weights": (_i32, [_P(Weights), _vp]),
    "ecord_fold_q_host": (_i32, [_vp] * 6),
    "ecord_policy_step": (_i32, [_P(Graph), _P(Env), _P(Weights), _P(Policy), _i32, _i32, _i32, _vp]),
    "ecord_policy_pre": (_i32, [_P(Graph), _P(Env), _P(Weights), _P(Policy), _vp]),
    "ecord_q_select": (_i32, [_P(Graph), _P(Env), _P(Weights), _P(Policy), _i32, _f32, _u64, _i32, _vp, _i32, _f32, _vp]),
    "ecord_node_init_sample": (_i32, [_P(Graph), _i32, _i32, _vp, _vp, _vp, _i32, _i32, _f32, _u64, _vp, _vp]),
    "ecord_rollout_plan_create": (_i32, [_P(RolloutDesc), _vp, _P(_vp)]),
    "ecord_rollout_plan_run": (_i32, [_vp, _i32, _P(_i64)]),
    "ecord_rollout_plan_reset": (_i32, [_vp]),
    "ecord_rollout_plan_destroy": (_i32, [_vp]),
    "ecord_score_log_max": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "ecord_plane_to_nt": (_i32, [_P(Graph), _i32, _vp, _vp, _vp]),
    "ecord_t_opt_steps": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "ecord_env_snapshot_padded": (_i32, [_P(Graph), _P(Env), _vp, _vp, _vp]),
    "ecord_stamp_time": (_i32, [_vp, _vp]),
    "ecord_best_state_snapshot": (_i32, [_P(Graph), _P(Env), _vp, _vp, _i32, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises EcordError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ECORD_B200_LIB") or LIB_PATH      # diagnostics: a variant build of the same library (build.py -D...)
    if not os.path.exists(path):
        raise EcordError(f"{path} not found: build it with `python -m ecord_b200.build` "
                         "(there is no CPU / PyTorch fallback for the rollout path)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError here = header/library drift
        fn.restype, fn.argtypes = res, args
    if lib.ecord_version() != 3:
        raise EcordError(f"ABI version mismatch: library reports {lib.ecord_version()}")
    for which, st in enumerate((Graph, Env, Weights, Policy, RolloutDesc)):
        if lib.ecord_abi_sizeof(which) != C.sizeof(st):
            raise EcordError(f"struct {st.__name__}: ctypes mirror is {C.sizeof(st)} bytes, "
                             f"library has {lib.ecord_abi_sizeof(which)}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ecord_error_string(int(rc)).decode()
        raise EcordError(f"{what or 'ecord call'} failed: [{rc}] {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
