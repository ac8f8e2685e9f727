"""Device-side objects of the rollout engine: graph batch, weights, and per-rollout state.

PyTorch is used for device memory, streams and (in dist.py) torch.distributed only; every computation
on the hot path is a kernel of libecord_b200.so reached through the C ABI (include/ecord_b200.h).
There is no CPU or eager-PyTorch fallback: without a CUDA device these classes raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GEMM_BF16, GEMM_BF16X3, GEMM_FP32, Q_EXACT, Q_FAST, Q_TC, check, ptr
from .network import check_state_dicts

# "bf16x3" (the default): tcgen05 GEMMs on bf16 hi/lo operand pairs, three products, fp32-grade gate functions -- fp32-grade
# recurrent state on the tensor cores; "bf16": single bf16 products + MUFU.TANH (fast, 2e-2 tolerance); "fp32": CUDA-core SGEMM
_GEMM_MODES = {"fp32": GEMM_FP32, "bf16": GEMM_BF16, "bf16x3": GEMM_BF16X3}
_Q_MODES = {"exact": Q_EXACT, "fast": Q_FAST, "tc": Q_TC}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise _lib.EcordError("ecord_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_ROLLOUT_STREAMS = {}


def _rollout_stream(device):
    """One long-lived side stream per device for all RolloutStates: torch's caching allocator keeps a pool per
    stream, so a fresh stream per rollout would turn every allocation made under it into a cudaMalloc."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _ROLLOUT_STREAMS:
        _ROLLOUT_STREAMS[key] = torch.cuda.Stream(device=device)
    return _ROLLOUT_STREAMS[key]


def _on_stream(fn):
    """Run a RolloutState method on the state's own (non-default) stream.  CUDA graphs cannot be captured on
    the legacy default stream, so every kernel of a rollout is enqueued on `self.stream`; the caller's current
    stream is ordered before and after the call, which keeps plain torch code around it correct."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        cur = torch.cuda.current_stream(self.device)
        if cur == self.stream:
            return fn(self, *a, **k)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = fn(self, *a, **k)
        cur.wait_stream(self.stream)
        return out
    return wrapped


class DeviceGraph:
    """Uploads one batch (data.TgBatch) and fills the C struct `ecord_graph`."""

    def __init__(self, tg, device=None):
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.tg = tg
        N, G, E = tg.num_nodes, tg.num_graphs, tg.num_edges
        src, dst, w = tg.np_src, tg.np_dst, tg.np_w
        deg = np.bincount(src, minlength=N)
        row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        # incoming lists in GLOBAL EDGE ORDER: stable sort of the edge list by destination
        order = np.argsort(dst, kind="stable")
        in_ptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=N))]).astype(np.int32)
        gep = np.concatenate([[0], np.cumsum(tg.np_edge_counts)]).astype(np.int64)
        dev = self.device
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.graph_ptr = up(tg.np_ptr.astype(np.int32))
        self.node_graph = up(np.repeat(np.arange(G), tg.num_nodes_list).astype(np.int32))
        self.row_ptr = up(row_ptr)
        self.col = up(dst.astype(np.int32))
        self.w = up(w.astype(np.float32))
        self.graph_edge_ptr = up(gep)
        self.in_ptr = up(in_ptr)
        self.in_src = up(src[order].astype(np.int32))
        self.in_w = up(w[order].astype(np.float32))
        self.N, self.G, self.E = N, G, E
        self.n_max = int(tg.num_nodes_list.max())
        self.num_nodes_list = tg.num_nodes_list
        # vertex tiles for the fast Q kernel
        tiles_per_graph = (tg.num_nodes_list + _lib.QTILE - 1) // _lib.QTILE
        tg_ids = np.repeat(np.arange(G), tiles_per_graph).astype(np.int32)
        starts = np.concatenate([[0], np.cumsum(tiles_per_graph)])[:-1]
        tv0 = ((np.arange(tg_ids.shape[0]) - np.repeat(starts, tiles_per_graph)) * _lib.QTILE).astype(np.int32)
        self.qtile_graph, self.qtile_v0 = up(tg_ids), up(tv0)
        # per-graph sum of the directed edge weights (env init: cut = (sum w - sum peek) / 4, exact for integer weights)
        w64 = w.astype(np.float64)
        edge_graph = np.repeat(np.arange(G), tg.np_edge_counts)
        wsum = np.bincount(edge_graph, weights=w64, minlength=G)
        wabs = np.bincount(edge_graph, weights=np.abs(w64), minlength=G)
        int_weights = int(bool(np.all(w64 == np.rint(w64)) and (wabs.max() if G else 0) < 2 ** 24))
        self.graph_wsum = up(wsum.astype(np.float32))
        self.c = _lib.Graph(G, N, E, self.n_max, ptr(self.graph_ptr), ptr(self.node_graph), ptr(self.row_ptr),
                            ptr(self.col), ptr(self.w), ptr(self.graph_edge_ptr), ptr(self.in_ptr),
                            ptr(self.in_src), ptr(self.in_w), int(tg_ids.shape[0]), int_weights, ptr(self.qtile_graph),
                            ptr(self.qtile_v0), ptr(self.graph_wsum))

    def padded_row_index(self):
        """int64 [N] on the device: row of vertex v in a [G * n_max, ...] padded array (graph * n_max + local id)."""
        if getattr(self, "_pad_idx", None) is None:
            ptrs = torch.from_numpy(self.tg.np_ptr.astype(np.int64))
            gid = torch.repeat_interleave(torch.arange(self.G), torch.from_numpy(self.num_nodes_list.astype(np.int64)))
            self._pad_idx = (gid * self.n_max + (torch.arange(self.N) - ptrs[:-1][gid])).to(self.device)
        return self._pad_idx

    def pinned_buffer(self, shape):
        """A cached page-locked float32 host buffer (results are copied out of it by the caller)."""
        cache = self.__dict__.setdefault("_pinned", {})
        if shape not in cache:
            cache[shape] = torch.empty(shape, dtype=torch.float32).pin_memory()
        return cache[shape]

    def plane_index(self, T):
        """int64 [N,T]: plane offset of every (vertex, trajectory) -- host-side layout helper for views."""
        ptrs = torch.from_numpy(self.tg.np_ptr)
        n_of = torch.from_numpy(self.num_nodes_list)[self.tg.batch]
        lo = ptrs[:-1][self.tg.batch]
        v_local = torch.arange(self.N) - lo
        t = torch.arange(T)
        return (lo * T).unsqueeze(1) + t.unsqueeze(0) * n_of.unsqueeze(1) + v_local.unsqueeze(1)


class DeviceWeights:
    """fp32 weights on the device under the reference's names + the packed / folded forms."""

    def __init__(self, state_dicts, device=None, pack_bf16=True):
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        sd = check_state_dicts(state_dicts)
        self.host = sd
        self.t = {}
        kw = {}
        for field, mod, key in _lib.WEIGHT_FIELDS:
            self.t[field] = sd[mod][key].to(self.device).contiguous()
            kw[field] = ptr(self.t[field])
        # folded Q-head constants (host array, passed to the kernel by value)
        r = sd['rnn']
        self.q_fold = torch.empty(_lib.QFOLD_FLOATS, dtype=torch.float32)
        check(self.lib.ecord_fold_q_host(ptr(r['mlp_out.0.weight']), ptr(sd['node_encoder']['0.weight']),
                                         ptr(r['mlp_out.1.weight']), ptr(r['mlp_out.1.bias']),
                                         ptr(r['mlp_out.3.weight']), ptr(self.q_fold)), "ecord_fold_q_host")
        self.gru_w_bf16 = self.p1_w_bf16 = self.q_tc = None
        if pack_bf16:
            self.q_tc = torch.zeros(_lib.QTC_FLOATS, dtype=torch.float32, device=self.device)
            self.gru_w_bf16 = torch.empty(_lib.GRU_PACK_ELEMS, dtype=torch.bfloat16, device=self.device)
            self.p1_w_bf16 = torch.empty(_lib.P1_PACK_ELEMS, dtype=torch.bfloat16, device=self.device)
        self.c = _lib.Weights(gru_w_bf16=ptr(self.gru_w_bf16), p1_w_bf16=ptr(self.p1_w_bf16), q_tc=ptr(self.q_tc),
                              q_fold=ptr(self.q_fold), **kw)
        if pack_bf16:
            check(self.lib.ecord_pack_weights(C.byref(self.c), _stream()), "ecord_pack_weights")


class RolloutState:
    """Environment + policy state of one rollout batch (G graphs x T trajectories) on one GPU.

    Step-wise methods (`policy_step`, `q_select`, `env_step`) expose the individual kernels for parity
    tests and `log_stats`; `run()` executes whole steps through the CUDA-graph plan."""

    def __init__(self, dgraph, weights, num_tradj, gemm="bf16x3", compat=True, tau=0.0, seed=0,
                 max_steps=0, log_scores=True, log_actions=True, steps_per_graph=16, dump_q=False, q_mode=None,
                 stream=None, epsilon=0.0):
        self.lib = _lib.load()
        self.g, self.w = dgraph, weights
        dev = dgraph.device
        self.device = dev
        self.stream = stream if stream is not None else _rollout_stream(dev)
        self.T = T = int(num_tradj)
        self.gemm = _GEMM_MODES[gemm]
        if self.gemm != GEMM_FP32 and weights.gru_w_bf16 is None:
            raise ValueError("weights were created with pack_bf16=False")
        self.compat, self.tau, self.seed = bool(compat), float(tau), int(seed)
        self.epsilon = float(epsilon)     # epsilon-greedy (actors/ecord.py:171-180); fixed for the life of the CUDA-graph plan
        if not 0.0 <= self.epsilon <= 1.0:
            raise ValueError("epsilon must lie in [0, 1]")
        # parity mode = (fp32 GEMM, exact Q); throughput mode = (bf16 tcgen05 GEMM, tensor-core Q).  tau > 0 (sampling) is
        # implemented by the exact kernel (exponential race, two passes) and by the tensor-core kernel (Gumbel-max, one pass).
        if q_mode is None:
            q_mode = "tc" if gemm != "fp32" else "exact"
        self.q_mode = _Q_MODES[q_mode]
        if self.q_mode == Q_FAST and self.tau > 0:
            raise ValueError("tau > 0 (sampling) requires q_mode='exact' or 'tc'")
        if self.q_mode == Q_TC and weights.q_tc is None:
            raise ValueError("weights were created with pack_bf16=False")
        self.fused = self.q_mode != Q_EXACT
        N, G = dgraph.N, dgraph.G
        M = G * T
        Mp = (M + 255) // 256 * 256          # 256: the tcgen05 GEMMs run on CTA pairs (cta_group::2, 256-row tiles)
        self.M, self.Mp = M, Mp
        f32 = dict(dtype=torch.float32, device=dev)
        # environment (trajectory-major planes)
        self.peek = torch.empty(N * T, **f32)
        self.sp = torch.empty(N * T, dtype=torch.int32, device=dev)
        self.best_state = torch.empty(N * T, dtype=torch.uint8, device=dev)
        self.score = torch.empty(M, **f32)
        self.best_score = torch.empty(M, **f32)
        self.best_step = torch.empty(M, dtype=torch.int32, device=dev)
        # compat=False also lifts quirk D: the best-state plane is a true snapshot (copied whole on every new best)
        self.env_c = _lib.Env(T, int(not self.compat), ptr(self.peek), ptr(self.sp), ptr(self.best_state), ptr(self.score),
                              ptr(self.best_score), ptr(self.best_step))
        # policy
        self.hidden = torch.zeros(2, Mp, 1024, **f32)
        self.u = torch.zeros(Mp, 64, **f32)
        self.y1 = torch.zeros(Mp, 512, **f32)
        self.P = torch.zeros(Mp, 32, **f32)
        self.A = torch.zeros(Mp, 64, **f32)
        self.v = torch.zeros(Mp, **f32)
        self.last_sel = torch.zeros(Mp, 32, **f32)
        self.actions = torch.zeros(M, dtype=torch.int32, device=dev)
        self.hidden_bf16 = self.th_bf16 = self.u_bf16 = self.gates = self.th = None
        self.split = int(self.gemm == GEMM_BF16X3)
        if self.gemm == GEMM_BF16:
            self.hidden_bf16 = torch.zeros(2, Mp, 1024, dtype=torch.bfloat16, device=dev)
            self.th_bf16 = torch.zeros(Mp, 1024, dtype=torch.bfloat16, device=dev)
            self.u_bf16 = torch.zeros(Mp, 64, dtype=torch.bfloat16, device=dev)
        elif self.gemm == GEMM_BF16X3:          # hi and lo planes (include/ecord_b200.h: ecord_policy)
            self.hidden_bf16 = torch.zeros(2, 2, Mp, 1024, dtype=torch.bfloat16, device=dev)
            self.th_bf16 = torch.zeros(2, Mp, 1024, dtype=torch.bfloat16, device=dev)
            self.u_bf16 = torch.zeros(2, Mp, 64, dtype=torch.bfloat16, device=dev)
        else:
            self.gates = torch.zeros(Mp, 6144, **f32)
            self.th = torch.zeros(Mp, 1024, **f32)
        self.q = torch.zeros(N * T, **f32) if (dump_q or (self.tau > 0 and self.q_mode == Q_EXACT)) else None
        self.gnn_out = torch.empty(N, 16, **f32)
        self.node_emb = torch.empty(N, 16, **f32)
        self.node_B = torch.empty(N, 64, **f32) if self.q_mode != Q_TC else None      # read by the exact / fast Q kernels only
        self.Ac = self.LA = self.node_Bc = self.node_LB = self.keys = self.qa = self.node_qv = None
        if self.q_mode == Q_FAST:
            self.Ac = torch.zeros(Mp, 64, **f32)
            self.LA = torch.zeros(Mp, **f32)
            self.node_Bc = torch.empty(N, 64, **f32)
            self.node_LB = torch.empty(N, **f32)
        if self.q_mode == Q_TC:
            self.qa = torch.zeros(Mp, _lib.QA_STRIDE, **f32)
            self.node_qv = torch.empty(N, 8, **f32)
        if self.q_mode != Q_EXACT:
            self.keys = torch.zeros(M, dtype=torch.int64, device=dev)
        self.policy_c = _lib.Policy(M, Mp, ptr(self.hidden), ptr(self.hidden_bf16), ptr(self.th_bf16), ptr(self.u_bf16),
                                    ptr(self.u), ptr(self.gates), ptr(self.th), ptr(self.y1), ptr(self.P), ptr(self.A),
                                    ptr(self.v), ptr(self.last_sel), ptr(self.node_emb), ptr(self.node_B),
                                    ptr(self.actions), ptr(self.q), ptr(self.Ac), ptr(self.LA), ptr(self.node_Bc),
                                    ptr(self.node_LB), ptr(self.keys), ptr(self.qa), ptr(self.node_qv), self.split, 0)
        # logs + step counter
        self.max_steps = int(max_steps)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.score_log = torch.empty(max(self.max_steps, 1), M, **f32) if log_scores and max_steps > 0 else None
        self.action_log = (torch.empty(max(self.max_steps, 1), M, dtype=torch.int32, device=dev)
                           if log_actions and max_steps > 0 else None)
        # device clock (%globaltimer, ns) at the end of every step: per-step t_step / t_opt / max_time without host syncs
        self.step_time = torch.zeros(max(self.max_steps, 1), dtype=torch.int64, device=dev) if max_steps > 0 else None
        self.time_base = torch.zeros(1, dtype=torch.int64, device=dev)
        self.steps_per_graph = int(steps_per_graph)
        self.steps_done = 0
        self.plan = None
        self.init_state_dev = None
        self.kernel_launches = 0
        self.env_clock_extra = 0
        self.greedy_steps = 0
        self._best_out = torch.empty(M, **f32)

    def reset(self):
        """Make the state reusable for another rollout on the same graph batch: `env_init` re-initialises every
        buffer a rollout reads before writing, the captured CUDA graphs stay valid."""
        self.kernel_launches = 0
        self.env_clock_extra = 0
        self.greedy_steps = 0
        if self.plan is not None:
            check(self.lib.ecord_rollout_plan_reset(self.plan), "ecord_rollout_plan_reset")

    # ---- set-up -----------------------------------------------------------------------------
    @_on_stream
    def gnn_embed(self):
        """t=0 embedding (a5): fills gnn_out, node_emb and node_B."""
        ws = torch.empty(self.lib.ecord_gnn_workspace_bytes(self.g.N), dtype=torch.uint8, device=self.device)
        check(self.lib.ecord_gnn_embed(C.byref(self.g.c), C.byref(self.w.c), ptr(self.gnn_out), ptr(self.node_emb),
                                       ptr(self.node_B), ptr(self.node_Bc), ptr(self.node_LB), ptr(self.node_qv), ptr(ws),
                                       _stream()),
              "ecord_gnn_embed")
        self.kernel_launches += 10
        self._ws = ws

    @_on_stream
    def gnn_embed_from(self, full_graph, vertex_index):
        """t=0 embedding of a SUB-batch (graph sharding, dist.py): the GNN runs over `full_graph` (the DeviceGraph of the whole
        batch, whose batch-wide LayerNorm statistics the reference's embedding depends on -- quirk B) and the rows of this
        state's vertices (`vertex_index` i64 [N_sub] on the device, indices into the full batch) are gathered."""
        fg = full_graph
        f32 = dict(dtype=torch.float32, device=self.device)
        full = dict(gnn_out=torch.empty(fg.N, 16, **f32), node_emb=torch.empty(fg.N, 16, **f32), node_B=torch.empty(fg.N, 64, **f32) if self.node_B is not None else None,
                    node_Bc=torch.empty(fg.N, 64, **f32) if self.node_Bc is not None else None,
                    node_LB=torch.empty(fg.N, **f32) if self.node_LB is not None else None,
                    node_qv=torch.empty(fg.N, 8, **f32) if self.node_qv is not None else None)
        ws = torch.empty(self.lib.ecord_gnn_workspace_bytes(fg.N), dtype=torch.uint8, device=self.device)
        check(self.lib.ecord_gnn_embed(C.byref(fg.c), C.byref(self.w.c), ptr(full['gnn_out']), ptr(full['node_emb']),
                                       ptr(full['node_B']), ptr(full['node_Bc']), ptr(full['node_LB']), ptr(full['node_qv']),
                                       ptr(ws), _stream()), "ecord_gnn_embed")
        self.kernel_launches += 10
        for name, src in full.items():
            if src is not None:
                torch.index_select(src, 0, vertex_index, out=getattr(self, name))

    @_on_stream
    def env_init(self, init_state):
        """a2: init_state int8 [N,T] (host numpy / torch, reference layout)."""
        if torch.is_tensor(init_state) and init_state.is_cuda:
            s = init_state.to(torch.int8).contiguous()
        else:
            s = torch.as_tensor(np.ascontiguousarray(np.asarray(init_state), dtype=np.int8))
        if tuple(s.shape) != (self.g.N, self.T):
            raise AssertionError(f"init_state must have shape {(self.g.N, self.T)}.  Got {tuple(s.shape)}.")
        self.init_state_dev = s.to(self.device, non_blocking=False)
        check(self.lib.ecord_env_init(C.byref(self.g.c), C.byref(self.env_c), ptr(self.init_state_dev), _stream()),
              "ecord_env_init")
        self.kernel_launches += 2
        self.hidden.zero_()
        if self.hidden_bf16 is not None:
            self.hidden_bf16.zero_()
        self.last_sel.zero_()
        self.step_counter.zero_()
        self.steps_done = 0
        if self.keys is not None:
            self.keys.zero_()
        if self.fused:   # prime u = msg_in([0; glob_0]); afterwards the fused env step keeps it current
            check(self.lib.ecord_policy_pre(C.byref(self.g.c), C.byref(self.env_c), C.byref(self.w.c),
                                            C.byref(self.policy_c), _stream()), "ecord_policy_pre")
            self.kernel_launches += 1

    @_on_stream
    def network_init_state(self, head_w, head_b, head=0, perturb_epsilon=0.0, tradj_offset=0, draw=0):
        """Initial spins drawn from the node classifier on the GNN embedding (use_network_initialisation,
        solver.py:256-277); needs gnn_embed() first.  head_w [H,16], head_b [H] device tensors.  Returns int8 [N,T]
        on the device, the form env_init takes.  `draw` numbers the call: each one is a fresh sample."""
        out = torch.empty(self.g.N, self.T, dtype=torch.int8, device=self.device)
        check(self.lib.ecord_node_init_sample(C.byref(self.g.c), self.T, int(tradj_offset), ptr(self.gnn_out), ptr(head_w), ptr(head_b),
                                              int(head_w.shape[0]), int(head), float(perturb_epsilon),
                                              (self.seed + 0x51ED270B + 0x9E3779B97F4A7C15 * int(draw)) & (2 ** 64 - 1), ptr(out),
                                              _stream()), "ecord_node_init_sample")
        self.kernel_launches += 1
        return out

    # ---- step-wise API (tests, log_stats) -----------------------------------------------------
    @_on_stream
    def policy_step(self):
        check(self.lib.ecord_policy_step(C.byref(self.g.c), C.byref(self.env_c), C.byref(self.w.c), C.byref(self.policy_c),
                                         self.steps_done & 1, self.gemm, int(self.fused), _stream()), "ecord_policy_step")
        self.kernel_launches += (6 if self.gemm == GEMM_FP32 else 4) - int(self.fused)

    @_on_stream
    def q_select(self, forced_actions=None):
        fa = None
        if forced_actions is not None:
            fa = torch.as_tensor(forced_actions).to(self.device, torch.int32).contiguous().view(-1)
            assert fa.numel() == self.M
        check(self.lib.ecord_q_select(C.byref(self.g.c), C.byref(self.env_c), C.byref(self.w.c), C.byref(self.policy_c),
                                      self.steps_done, self.tau, self.seed, int(self.compat), ptr(fa), self.q_mode,
                                      self.epsilon, _stream()), "ecord_q_select")
        self.kernel_launches += 1 if self.q_mode == Q_EXACT else 2
        self._fa = fa

    @_on_stream
    def env_step(self, actions=None):
        a = self.actions if actions is None else torch.as_tensor(actions).to(self.device, torch.int32).contiguous().view(-1)
        row = None
        if self.score_log is not None and self.steps_done < self.max_steps:
            row = self.score_log[self.steps_done]
            if self.action_log is not None:
                self.action_log[self.steps_done].copy_(a)
        if self.fused:
            if actions is not None:
                self.actions.copy_(a)
            check(self.lib.ecord_env_step_fused(C.byref(self.g.c), C.byref(self.env_c), C.byref(self.w.c),
                                                C.byref(self.policy_c), self.steps_done, ptr(row), _stream()),
                  "ecord_env_step_fused")
        else:
            check(self.lib.ecord_env_step(C.byref(self.g.c), C.byref(self.env_c), ptr(a), self.steps_done, ptr(row),
                                          _stream()), "ecord_env_step")
        self.kernel_launches += 1
        if self.step_time is not None and self.steps_done < self.max_steps:      # the plan's env kernel stamps this itself
            check(self.lib.ecord_stamp_time(C.c_void_p(self.step_time.data_ptr() + 8 * self.steps_done), _stream()), "ecord_stamp_time")
        self.steps_done += 1
        self.step_counter.fill_(self.steps_done)

    # ---- greedy local search on the peeks (reference solvers/solver.py:38-95) -----------
    @_on_stream
    def greedy_solve(self, init_from_best=False, tau=0.0, max_iters=None, rebase=False):
        """greedy_solve(env, tau, init_from_best): every trajectory flips a vertex of maximal peek (uniformly at random
        among ties) until each of them has been at a local optimum at some point; trajectories that got there early
        keep stepping, exactly as the reference's `while not done.all()` loop does.  The flips are environment steps
        (scores, best scores and the static counters move) but not rollout steps: nothing is logged and the policy
        state is untouched.  rebase=True (pre-solve) puts the environment clock back to the rollout's step 0
        afterwards.  Returns the number of steps taken."""
        M = self.M
        clock = self.steps_done
        if init_from_best:
            check(self.lib.ecord_env_reset_state(C.byref(self.g.c), C.byref(self.env_c), None, clock, _stream()),
                  "ecord_env_reset_state")
        acts = torch.empty(M, dtype=torch.int32, device=self.device)
        flags = torch.empty(M, dtype=torch.int32, device=self.device)

        def select(step_no):
            check(self.lib.ecord_greedy_select(C.byref(self.g.c), C.byref(self.env_c), float(tau), self.seed + 0x9E3779B9,
                                               step_no, ptr(acts), ptr(flags), _stream()), "ecord_greedy_select")
            self.kernel_launches += 1

        select(clock)
        done = flags.clone()
        iters = 0
        while not bool(done.all()):                       # one 4-byte D2H per step: the loop length is data dependent
            if max_iters is not None and iters >= max_iters:
                break
            check(self.lib.ecord_env_step(C.byref(self.g.c), C.byref(self.env_c), ptr(acts), clock + iters, None, _stream()),
                  "ecord_env_step")
            self.kernel_launches += 1
            iters += 1
            select(clock + iters)
            done |= flags
        self.greedy_steps += iters
        if rebase and iters:
            check(self.lib.ecord_env_rebase_clock(C.byref(self.g.c), C.byref(self.env_c), iters, _stream()),
                  "ecord_env_rebase_clock")
            self.kernel_launches += 1
        elif not rebase:
            self.env_clock_extra += iters
        if self.fused:      # the msg_in input of the next policy step depends on (best - score): refresh it
            check(self.lib.ecord_policy_pre(C.byref(self.g.c), C.byref(self.env_c), C.byref(self.w.c),
                                            C.byref(self.policy_c), _stream()), "ecord_policy_pre")
            self.kernel_launches += 1
        return iters

    def step(self, forced_actions=None):
        self.policy_step()
        self.q_select(forced_actions)
        self.env_step()

    # ---- fused loop -------------------------------------------------------------------------
    def _make_plan(self):
        desc = _lib.RolloutDesc(C.pointer(self.g.c), C.pointer(self.env_c), C.pointer(self.w.c), C.pointer(self.policy_c),
                                self.gemm, self.q_mode, int(self.compat), self.epsilon, self.tau, self.steps_per_graph, self.seed,
                                ptr(self.step_counter), ptr(self.score_log), ptr(self.action_log), self.max_steps, 0,
                                ptr(self.step_time))
        out = C.c_void_p()
        check(self.lib.ecord_rollout_plan_create(C.byref(desc), _stream(), C.byref(out)), "ecord_rollout_plan_create")
        self.plan = out

    @_on_stream
    def run(self, num_steps):
        """Enqueue `num_steps` whole steps (asynchronous).  Uses CUDA graphs of `steps_per_graph` steps."""
        if self.plan is None:
            self._make_plan()
        n = C.c_int64(0)
        check(self.lib.ecord_rollout_plan_run(self.plan, int(num_steps), C.byref(n)), "ecord_rollout_plan_run")
        self.kernel_launches += n.value
        self.steps_done += int(num_steps)

    @_on_stream
    def stamp_time_base(self):
        """Sample the device clock on the rollout stream: the origin of `step_time` (call right before the first run())."""
        check(self.lib.ecord_stamp_time(ptr(self.time_base), _stream()), "ecord_stamp_time")

    def step_times_host(self, num_steps):
        """Seconds each of the first `num_steps` steps took (device clock; the first one counts from the time base)."""
        S = int(num_steps)
        if S <= 0:
            return torch.zeros(0, dtype=torch.float64)
        t = torch.cat([self.time_base, self.step_time[:S]]).cpu().double()
        return (t[1:] - t[:-1]) * 1e-9

    @_on_stream
    def t_opt_steps(self, num_steps, init_score):
        """Per unit, the last step (1-based) at which the score rose to its running best (0: never) -- what the reference's
        log['t_opt'] marks (solver.py:241-242)."""
        out = torch.empty(self.M, dtype=torch.int32, device=self.device)
        check(self.lib.ecord_t_opt_steps(ptr(self.score_log), ptr(init_score.contiguous().view(-1)), int(num_steps), self.M,
                                         ptr(out), _stream()), "ecord_t_opt_steps")
        return out.view(self.g.G, self.T)

    # ---- log_stats (solvers/solver.py:244-249,322-327) ----------------------------------------
    @_on_stream
    def clone_env(self):
        """A copy of the environment planes as they are now (the spins / peeks the rollout starts from, after any greedy
        pre-solve): the base from which replay_stats() reconstructs the per-step views."""
        t = dict(peek=self.peek.clone(), sp=self.sp.clone(), best_state=self.best_state.clone(), score=self.score.clone(),
                 best_score=self.best_score.clone(), best_step=self.best_step.clone())
        c = _lib.Env(self.T, 0, ptr(t['peek']), ptr(t['sp']), ptr(t['best_state']), ptr(t['score']), ptr(t['best_score']),
                     ptr(t['best_step']))
        return t, c, self.steps_done

    def replay_stats(self, base, num_steps, chunk=8):
        """Per-step `peeks` and `state` of a finished rollout in the reference's layout (f32 [G, n_max, T] on the host,
        padding -1e10 / -1): yields (step, peeks, state) for step = 0 .. num_steps - 1, each the view BEFORE that step's flip,
        exactly what DQNSolver.rollout appends with log_stats=True.

        The rollout itself ran through the CUDA-graph plan with no host round trip; here the action log is replayed onto a
        copy of the initial environment (one snapshot kernel + one env-step kernel per step, no policy kernels), `chunk`
        steps at a time into one of two device buffers whose device->host copy into pinned memory runs on a side stream
        while the next chunk is being replayed.  The reference copies both arrays to the host every step inside its loop."""
        tensors, env_c, step0 = base
        G, nmax, T = self.g.G, self.g.n_max, self.T
        S = int(num_steps)
        chunk = max(1, min(int(chunk), S)) if S else 1
        dev_buf = [torch.empty(chunk, 2, G, nmax, T, dtype=torch.float32, device=self.device) for _ in range(2)]
        host_buf = [torch.empty(chunk, 2, G, nmax, T, dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=self.device)
        done = [torch.cuda.Event() for _ in range(2)]          # D2H of buffer i complete
        filled = [torch.cuda.Event() for _ in range(2)]        # replay of buffer i complete
        pending = []                                           # (buffer, first step, count)

        def drain(entry):
            b, s0, cnt = entry
            done[b].synchronize()
            for i in range(cnt):
                yield s0 + i, host_buf[b][i, 1].clone(), host_buf[b][i, 0].clone()

        st_ptr = C.c_void_p(self.stream.cuda_stream)

        def enqueue(c0, b, cnt):          # replay `cnt` steps into device buffer b, then its D2H on the side stream
            for i in range(cnt):
                s = c0 + i
                check(self.lib.ecord_env_snapshot_padded(C.byref(self.g.c), C.byref(env_c), ptr(dev_buf[b][i, 0]),
                                                         ptr(dev_buf[b][i, 1]), st_ptr), "ecord_env_snapshot_padded")
                check(self.lib.ecord_env_step(C.byref(self.g.c), C.byref(env_c), ptr(self.action_log[s]), step0 + s, None,
                                              st_ptr), "ecord_env_step")
            filled[b].record(self.stream)
            copy_stream.wait_event(filled[b])
            with torch.cuda.stream(copy_stream):
                host_buf[b][:cnt].copy_(dev_buf[b][:cnt], non_blocking=True)
                done[b].record(copy_stream)
            self.kernel_launches += 2 * cnt

        for c0 in range(0, S, chunk):
            b = (c0 // chunk) & 1
            cnt = min(chunk, S - c0)
            if len(pending) == 2:         # buffer b still belongs to the copy issued two chunks ago: hand its steps out first
                yield from drain(pending.pop(0))
            enqueue(c0, b, cnt)
            pending.append((b, c0, cnt))
        for entry in pending:
            yield from drain(entry)
        del tensors

    def close(self):
        if self.plan is not None:
            self.lib.ecord_rollout_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- views in the reference's layouts (parity checks, logs) ------------------------------
    @_on_stream
    def export(self, normalise=False, nodes=True):
        """(node_features [N,T,3], glob_features [G,T,2], best_states [N,T]) as torch CUDA tensors
        (nodes=False skips the node features: None is returned in their place)."""
        N, G, T = self.g.N, self.g.G, self.T
        nf = torch.empty(N, T, 3, dtype=torch.float32, device=self.device) if nodes else None
        gf = torch.empty(G, T, 2, dtype=torch.float32, device=self.device)
        if nodes:
            bs = torch.empty(N, T, dtype=torch.float32, device=self.device)
        else:                   # the end-of-rollout view: a cached buffer (consumed before the next export)
            if getattr(self, "_bs_buf", None) is None:
                self._bs_buf = torch.empty(N, T, dtype=torch.float32, device=self.device)
            bs = self._bs_buf
        check(self.lib.ecord_env_export(C.byref(self.g.c), C.byref(self.env_c), self.steps_done, int(normalise),
                                        ptr(nf), ptr(gf), ptr(bs), _stream()), "ecord_env_export")
        return nf, gf, bs

    @_on_stream
    def plane_to_nt(self, plane):
        out = torch.empty(self.g.N, self.T, dtype=torch.float32, device=self.device)
        check(self.lib.ecord_plane_to_nt(C.byref(self.g.c), self.T, ptr(plane), ptr(out), _stream()), "ecord_plane_to_nt")
        return out

    def q_values(self):
        """Q-values of the last q_select in the reference layout [N,T] (needs dump_q=True)."""
        return self.plane_to_nt(self.q)

    def hidden_state(self):
        """Current GRU state [G,T,1024]."""
        return self.hidden[self.steps_done & 1, :self.M].view(self.g.G, self.T, 1024)

    def scores(self):
        return self.score.view(self.g.G, self.T)

    @_on_stream
    def score_log_host(self, num_steps):
        """The first `num_steps` rows of the score log as a host tensor [S, G, T]: one D2H into a page-locked tensor
        owned by the caller."""
        S, G, T = int(num_steps), self.g.G, self.T
        host = torch.empty((S, self.M), dtype=torch.float32, pin_memory=True)
        host.copy_(self.score_log[:S], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.view(S, G, T)

    @_on_stream
    def best_from_log(self, num_steps):
        out = self._best_out
        check(self.lib.ecord_score_log_max(ptr(self.score_log), int(num_steps), self.M, ptr(out), _stream()),
              "ecord_score_log_max")
        return out.view(self.g.G, self.T)

    @_on_stream
    def best_states_device(self, true_snapshot=False):
        """[G, n_max, T] f32 on the device, -1 padded (tradjectories.py:324-328,378-385).  true_snapshot=False gives the
        reference's semantics (quirk D); True replays the action log up to each trajectory's best step.  The result is a
        cached buffer: consume (copy) it before the next call."""
        N, G, T = self.g.N, self.g.G, self.T
        if true_snapshot and not self.compat:
            # compat=False keeps the best-state plane as a true snapshot on the device (ecord_env.true_snapshot): greedy
            # pre- / post-solve flips and rebased clocks are covered, nothing has to be replayed
            flat = self.export(nodes=False)[2]
        elif true_snapshot:
            if self.action_log is None or self.init_state_dev is None:
                raise ValueError("true_snapshot needs log_actions=True")
            if self.greedy_steps:
                raise ValueError("greedy pre-/post-solve flips are not in the action log: use compat=False, which keeps a "
                                 "true best-state snapshot on the device")
            plane = torch.empty(N * T, dtype=torch.uint8, device=self.device)
            check(self.lib.ecord_best_state_snapshot(C.byref(self.g.c), C.byref(self.env_c), ptr(self.init_state_dev),
                                                     ptr(self.action_log), min(self.steps_done, self.max_steps),
                                                     ptr(plane), _stream()), "ecord_best_state_snapshot")
            flat = self.plane_to_nt(plane.float())
        else:
            flat = self.export(nodes=False)[2]
        if G * self.g.n_max == N:          # equal-sized graphs: [N, T] is [G, n_max, T] already
            return flat.view(G, self.g.n_max, T)
        # ragged batch: pad on the device (one index_copy into a cached buffer whose padding rows stay -1)
        if getattr(self, "_pad", None) is None:
            self._pad = torch.full((G * self.g.n_max, T), -1.0, dtype=torch.float32, device=self.device)
        pad = self._pad
        pad.index_copy_(0, self.g.padded_row_index(), flat)
        return pad.view(G, self.g.n_max, T)

    def best_states_batched(self, true_snapshot=False):
        """best_states_device() on the host (a page-locked tensor owned by the caller)."""
        dev = self.best_states_device(true_snapshot)
        host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host
