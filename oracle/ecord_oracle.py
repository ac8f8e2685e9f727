"""oracle/ecord_oracle.py -- CPU restatement of the reference's test-time rollout (ECORD).

TEST INFRASTRUCTURE ONLY.  The product package (ecord_b200/) never imports this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
only as the checker or as the timed CPU baseline.

What is restated (citations are into /root/reference):
  a1  build_batch()          ecord/environment/data.py:53-83,114-128   (networkx -> batched COO/CSR)
  a2  EnvOracle.__init__     ecord/environment/tradjectories.py:18-148 (init scores / peeks / features)
  a3  EnvOracle.step         ecord/environment/_tradjectory_step_cy.pyx:9-171 via env_oracle.c
                             (or the compiled reference itself, oracle/_ref, with use_ref=True)
  a4  EnvOracle.observe      tradjectories.py:336-376, environment.py:94-111
  a5  ModelOracle.embed      models/gnn_embedder.py:133-149,182-199; models/gnn_to_rnn.py:16-18;
                             solvers/actors/_base.py:185-237
  a6  ModelOracle.update     solvers/actors/_base.py:247-269; models/rnn_decoder.py:369-381
  a7  ModelOracle.node_emb   solvers/actors/_base.py:279-294; solvers/actors/ecord.py:78-91
  a8  ModelOracle.q_values   models/rnn_decoder.py:195-236,285-311
  a9  ModelOracle.select     solvers/actors/ecord.py:123-193,12-19; actors/_base.py:296-308
  a10 rollout()              solvers/solver.py:196-377
  a11 test_solver()          validate/testing.py:120-230

Third-party arithmetic that is NOT under /root/reference (torch-scatter 2.0.5, torch-geometric
1.7.0; requirements.txt:7,9) is restated from the published semantics: scatter_max/min tie ->
lowest index; scatter_mean = sum / max(count,1); PyG LayerNorm(batch=None) = whole-tensor
(x-mean)/(std_pop+eps).  The reference has no tests for these boundaries => "parity unpinned"
at exact ties / the LN formula.  Everything else is pinned: tests/test_oracle.py checks this
file against tests/golden/*.npz, which oracle/gen_golden.py produced by running the UNMODIFIED
reference (with oracle/shims for the uninstalled packages and oracle/_ref for the Cython step).
"""
import ctypes
import os
import subprocess
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
CLIP_MAX = 40.0  # Tradjectories._clip_features_max, tradjectories.py:16

# canonical feature tables (experiments/train.py:62-70 -> tradjectories.py:189-300)
NODE_IDX_CANON = np.array([0, 1, -1, 2, -1], dtype=np.intc)      # state, peek, best_state, static, degree
GLOB_IDX_CANON = np.array([0, -1, -1, -1, -1, -1], dtype=np.intc)  # score_from_best, ...; 2nd feature dead (quirk C)


# --------------------------------------------------------------------------------------------
# native pieces
# --------------------------------------------------------------------------------------------
_LIB = None


def build_native(force=False):
    """Compile env_oracle.c -> libenv_oracle.so (gcc).  Building the checker is not using it."""
    out = os.path.join(HERE, "libenv_oracle.so")
    src = os.path.join(HERE, "env_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-ffp-contract=off",
                               "-o", out, src, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_native())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------
# a1: batching
# --------------------------------------------------------------------------------------------
class OracleBatch:
    """Arrays the reference's GraphBatcher + Tradjectories derive from a list of graphs."""

    def __init__(self, edge_index, edge_w, num_nodes_list, num_dir_edges_list):
        self.edge_index = np.ascontiguousarray(edge_index, dtype=np.int64)   # [2,E] global ids
        self.edge_w = np.ascontiguousarray(edge_w, dtype=np.float32)          # [E]
        self.num_nodes_list = np.asarray(num_nodes_list, dtype=np.int64)
        self.G = len(num_nodes_list)
        self.N = int(self.num_nodes_list.sum())
        self.E = self.edge_index.shape[1]
        self.ptr = np.concatenate([[0], np.cumsum(self.num_nodes_list)]).astype(np.int64)
        self.batch = np.repeat(np.arange(self.G), self.num_nodes_list).astype(np.int64)
        self.degree = np.bincount(self.edge_index[0], minlength=self.N).astype(np.int64)
        self.edge_ptr = np.concatenate([[0], np.cumsum(self.degree)]).astype(np.int32)  # CSR by source
        self.graph_edge_ptr = np.concatenate([[0], np.cumsum(num_dir_edges_list)]).astype(np.int64)
        self.edges_by_node = np.ascontiguousarray(self.edge_index.astype(np.int32))
        self.n_max = int(self.num_nodes_list.max())


def graph_to_edges(g):
    """One graph -> (n, src[], dst[], w[]) in the reference's edge order:
    GraphData.from_networkx (data.py:114-128) relabels nodes to integers in G.nodes order, makes
    the graph directed and lists G.edges; weights come from the 'weight' attribute in that order."""
    if isinstance(g, tuple) and len(g) == 4:   # (n, src, dst, w): directed edge arrays already in reference order
        return int(g[0]), np.asarray(g[1], np.int64), np.asarray(g[2], np.int64), np.asarray(g[3], np.float32)
    import networkx as nx
    if isinstance(g, np.ndarray):
        g = nx.from_numpy_array(g)          # graphs/utils.py:19-21
    g = nx.convert_node_labels_to_integers(g)
    g = g.to_directed() if not nx.is_directed(g) else g
    edges = list(g.edges)
    w = list(nx.get_edge_attributes(g, "weight").values())
    if len(w) != len(edges):                # unweighted graph: the reference would fail; use 1
        w = [1.0] * len(edges)
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    return g.number_of_nodes(), e[:, 0], e[:, 1], np.asarray(w, dtype=np.float32)


def build_batch(graphs):
    srcs, dsts, ws, ns, nes = [], [], [], [], []
    off = 0
    for g in graphs:
        n, s, d, w = graph_to_edges(g)
        srcs.append(s + off)
        dsts.append(d + off)
        ws.append(w)
        ns.append(n)
        nes.append(len(w))
        off += n
    ei = np.stack([np.concatenate(srcs), np.concatenate(dsts)])
    # the reference splits edge_index by source degree (data.py:57-58): requires sources ascending
    assert np.all(np.diff(ei[0]) >= 0), "edge list must be grouped by ascending source"
    return OracleBatch(ei, np.concatenate(ws), ns, nes)


# --------------------------------------------------------------------------------------------
# a2-a4: environment
# --------------------------------------------------------------------------------------------
class EnvOracle:
    """Tradjectories + Environment of the reference for the canonical feature set
    node = [STATE, PEEK_NORM, STEPS_STATIC_CLIP], glob = [SCORE_FROM_BEST_NORM, PEEK_MAX_NORM]
    (or any index tables in the reference's intc[5]/intc[6] format, un-normalised features)."""

    def __init__(self, ob, num_tradj, init_state, node_idx=NODE_IDX_CANON, glob_idx=GLOB_IDX_CANON,
                 num_node_features=3, num_glob_features=2, use_ref=False):
        self.ob, self.T = ob, int(num_tradj)
        N, G, T = ob.N, ob.G, self.T
        states = np.ascontiguousarray(np.asarray(init_state).astype(np.int8))
        assert states.shape == (N, T)
        self.node_idx = np.ascontiguousarray(node_idx, dtype=np.intc)
        self.glob_idx = np.ascontiguousarray(glob_idx, dtype=np.intc)
        self.Fn, self.Fg = int(num_node_features), int(num_glob_features)
        scores = np.zeros((G, T), np.float32)
        peeks = np.zeros((N, T), np.float32)
        _lib().oracle_scores_and_peeks(_p(states), _p(ob.edges_by_node), _p(ob.edge_w), _p(ob.edge_ptr),
                                       _p(ob.graph_edge_ptr), _p(scores), _p(peeks),
                                       ctypes.c_int32(N), ctypes.c_int32(G), ctypes.c_int32(T),
                                       ctypes.c_int32(ob.E))
        self.scores = scores
        self.best_scores = scores.copy()
        nf = np.zeros((N, T, self.Fn), np.float32)          # tradjectories.py:150-166
        nf[..., self.node_idx[0]] = states
        nf[..., self.node_idx[1]] = peeks
        if self.node_idx[2] >= 0:
            nf[..., self.node_idx[2]] = states
        self.node_features = nf
        gf = np.zeros((G, T, self.Fg), np.float32)          # tradjectories.py:168-187
        if self.glob_idx[2] >= 0:
            ng = np.zeros((G, T), np.float32)
            np.add.at(ng, ob.batch, (peeks > 0).astype(np.float32))
            gf[..., self.glob_idx[2]] = ng
        self.glob_features = gf
        self.best_states = states.astype(np.float32).copy()  # tradjectories.py:116
        self.init_scores = scores.copy()
        self._step_ref = None
        if use_ref:
            import build_ref
            self._step_ref = build_ref.load_step_cy()
            assert self._step_ref is not None, "oracle/_ref is not built"

    def step(self, actions_local):
        """actions_local: int64 [G,T] per-graph vertex ids (tradjectories.py:387-405 adds the offset)."""
        ob = self.ob
        a = np.ascontiguousarray(np.asarray(actions_local, dtype=np.int64) + ob.ptr[:-1, None])
        if self._step_ref is not None:
            self._step_ref(a, self.scores, self.best_scores, self.best_states, self.node_features,
                           self.glob_features, self.node_idx, self.glob_idx, ob.edges_by_node, ob.edge_w,
                           ob.edge_ptr, ob.N, ob.G, self.T)
        else:
            _lib().oracle_step_cy(_p(a), _p(self.scores), _p(self.best_scores), _p(self.best_states),
                                  _p(self.node_features), _p(self.glob_features), _p(self.node_idx),
                                  _p(self.glob_idx), _p(ob.edges_by_node), _p(ob.edge_w), _p(ob.edge_ptr),
                                  ctypes.c_int32(ob.N), ctypes.c_int32(ob.G), ctypes.c_int32(self.T),
                                  ctypes.c_int32(self.Fn), ctypes.c_int32(self.Fg), ctypes.c_int32(ob.E))

    # ---- greedy local search on the peeks (solvers/solver.py:38-95 greedy_solve), SURVEY.md 8f rank 1 ----
    def reinit_from_best(self):
        """init_from_best=True (solver.py:39-43): spins <- best_states, scores and peeks recomputed, the node
        features re-initialised (static counters back to 0); best_scores / best_states are left alone."""
        ob = self.ob
        states = np.ascontiguousarray(self.best_states.astype(np.int8))
        scores = np.zeros((ob.G, self.T), np.float32)
        peeks = np.zeros((ob.N, self.T), np.float32)
        _lib().oracle_scores_and_peeks(_p(states), _p(ob.edges_by_node), _p(ob.edge_w), _p(ob.edge_ptr),
                                       _p(ob.graph_edge_ptr), _p(scores), _p(peeks), ctypes.c_int32(ob.N),
                                       ctypes.c_int32(ob.G), ctypes.c_int32(self.T), ctypes.c_int32(ob.E))
        self.scores = scores
        nf = np.zeros_like(self.node_features)
        nf[..., self.node_idx[0]] = states
        nf[..., self.node_idx[1]] = peeks
        if self.node_idx[2] >= 0:
            nf[..., self.node_idx[2]] = states
        self.node_features = nf

    def greedy_solve(self, rng=None, init_from_best=False, max_iters=100000):
        """Every trajectory keeps flipping a vertex of maximal peek -- uniformly at random among ties, which is what
        the reference's scatter_softmax_sample(peeks / 1e-9) amounts to at tau = 0 -- until every one of them has
        been at a local optimum (max peek <= 0) at some point; trajectories that got there early keep stepping
        (solver.py:62-86).  Returns (number of steps, whether any arg-max was tied)."""
        rng = rng or np.random.RandomState(0)
        ob = self.ob
        if init_from_best:
            self.reinit_from_best()

        def select():
            pk = self.node_features[..., self.node_idx[1]]                      # [N,T]
            acts = np.zeros((ob.G, self.T), np.int64)
            lm = np.zeros((ob.G, self.T), bool)
            tie = False
            for g in range(ob.G):
                blk = pk[ob.ptr[g]:ob.ptr[g + 1]]                               # [n,T]
                mx = blk.max(0)
                lm[g] = mx <= 0
                for t in range(self.T):
                    cand = np.flatnonzero(blk[:, t] == mx[t])
                    tie |= cand.size > 1
                    acts[g, t] = cand[rng.randint(cand.size)] if cand.size > 1 else cand[0]
            return acts, lm, tie

        acts, done, tied = select()
        it = 0
        while not done.all() and it < max_iters:
            self.step(acts)
            acts, lm, tie = select()
            tied |= tie
            done |= lm
            it += 1
        return it, tied

    # canonical observation normalisation (tradjectories.py:336-376): peek / n_g, min(static,40)/40,
    # (best - score) / n_g; glob column 1 stays 0 (quirk C).
    def observe(self):
        ob = self.ob
        nf = self.node_features.copy()
        n_exp = ob.num_nodes_list[ob.batch].astype(np.float32)[:, None]
        nf[..., 1] = nf[..., 1] / n_exp
        nf[..., 2] = np.clip(nf[..., 2], None, CLIP_MAX) / CLIP_MAX
        gf = self.glob_features.copy()
        gf[..., 0] = gf[..., 0] / ob.num_nodes_list.astype(np.float32)[:, None]
        return torch.from_numpy(nf), torch.from_numpy(gf)

    def best_states_batched(self):
        """get_best_states(flat=False): [G, n_max, T], -1 padded (tradjectories.py:324-328,378-385)."""
        ob = self.ob
        out = -np.ones((ob.G, ob.n_max, self.T), np.float32)
        for g in range(ob.G):
            out[g, :ob.num_nodes_list[g]] = self.best_states[ob.ptr[g]:ob.ptr[g + 1]]
        return torch.from_numpy(out)


# --------------------------------------------------------------------------------------------
# a5-a9: model
# --------------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, 0.01)


def _pyg_layer_norm(x, eps=1e-5):
    x = x - x.mean()
    return x / (x.std(unbiased=False) + eps)


class ModelOracle:
    """Canonical ECORD network (experiments/train.py:55-96 via configs.py:83-209), fp32 on CPU.
    `sd` = {'gnn','gnn2rnn','node_encoder','rnn'} -> state_dict with the reference's key names."""

    def __init__(self, sd, compat=True):
        g, r = sd['gnn'], sd['rnn']
        f = lambda t: t.detach().to(torch.float32).cpu().clone()
        self.compat = compat
        self.embed_w, self.embed_b = f(g['embed.weight']), f(g['embed.bias'])
        self.snd_w, self.snd_b = f(g['node_model.msg_snd.0.weight']), f(g['node_model.msg_snd.0.bias'])
        self.rec_w, self.rec_b = f(g['node_model.msg_rec.0.weight']), f(g['node_model.msg_rec.0.bias'])
        self.gnn_cell = torch.nn.GRUCell(32, 16)
        self.gnn_cell.load_state_dict({k.split('.')[-1]: f(v) for k, v in g.items() if k.startswith('node_model.rnn.')})
        self.g2r_w, self.g2r_b = f(sd['gnn2rnn']['node_projection.0.weight']), f(sd['gnn2rnn']['node_projection.0.bias'])
        self.enc_w, self.enc_b = f(sd['node_encoder']['0.weight']), f(sd['node_encoder']['0.bias'])
        self.msg_w, self.msg_b = f(r['msg_in.0.weight']), f(r['msg_in.0.bias'])
        self.cell = torch.nn.GRUCell(64, 1024)
        self.cell.load_state_dict({k.split('.')[-1]: f(v) for k, v in r.items() if k.startswith('graph_rnn.0.')})
        self.p1_w, self.p1_b = f(r['proj_rnn_to_gnn.1.weight']), f(r['proj_rnn_to_gnn.1.bias'])
        self.ln1_w, self.ln1_b = f(r['proj_rnn_to_gnn.2.weight']), f(r['proj_rnn_to_gnn.2.bias'])
        self.p2_w, self.p2_b = f(r['proj_rnn_to_gnn.4.weight']), f(r['proj_rnn_to_gnn.4.bias'])
        self.ln2_w, self.ln2_b = f(r['proj_rnn_to_gnn.5.weight']), f(r['proj_rnn_to_gnn.5.bias'])
        self.v1_w, self.v1_b = f(r['val_out.1.weight']), f(r['val_out.1.bias'])
        self.v2_w, self.v2_b = f(r['val_out.3.weight']), f(r['val_out.3.bias'])
        self.q1_w, self.q1_b = f(r['mlp_out.0.weight']), f(r['mlp_out.0.bias'])
        self.qln_w, self.qln_b = f(r['mlp_out.1.weight']), f(r['mlp_out.1.bias'])
        self.q2_w, self.q2_b = f(r['mlp_out.3.weight']), f(r['mlp_out.3.bias'])
        self.hidden = None
        self.last_sel = None
        self.init_emb = None

    # a5 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def gnn_embed(self, ob):
        N = ob.N
        row = torch.from_numpy(ob.edge_index[0])
        col = torch.from_numpy(ob.edge_index[1])
        w = torch.from_numpy(ob.edge_w).view(-1, 1)
        deg_in = torch.zeros(N).index_add_(0, col, torch.ones(col.numel())).clamp_(min=1).view(-1, 1)
        x = F.linear(torch.ones(N, 1), self.embed_w, self.embed_b)
        x = _pyg_layer_norm(x)
        for _ in range(4):                                            # configs.py:99 hard-codes 4
            y = _lrelu(F.linear(x, self.snd_w, self.snd_b))
            msg = torch.zeros(N, 16).index_add_(0, col, y[row] * w) / deg_in   # scatter_mean
            inp = _lrelu(F.linear(torch.cat([msg, x], -1), self.rec_w, self.rec_b))
            x = self.gnn_cell(inp, x)
            x = _pyg_layer_norm(x)
        return x                                                     # [N,16] (pre gnn2rnn)

    @torch.no_grad()
    def initialise_embeddings(self, ob, T):
        self.gnn_out = self.gnn_embed(ob)
        self.init_emb = F.linear(self.gnn_out, self.g2r_w, self.g2r_b)   # [N,16]
        self.hidden = torch.zeros(ob.G, T, 1024)
        self.last_sel = torch.zeros(ob.G, T, 32)
        self.T = T

    # a6 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def update_internal_state(self, glob):
        x = torch.cat([self.last_sel, glob], -1)                    # [G,T,34]
        u = _lrelu(F.linear(x, self.msg_w, self.msg_b)).view(-1, 64)
        h = self.cell(u, self.hidden.view(-1, 1024))
        self.hidden = h.view(*self.hidden.shape)
        self.last_sel = None

    # a7 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def node_embeddings(self, node_feat):
        enc = F.linear(node_feat, self.enc_w, self.enc_b)            # [N,T,16]
        return torch.cat([self.init_emb.unsqueeze(1).expand(-1, node_feat.size(1), -1), enc], -1)

    # a8 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def project(self):
        p = torch.tanh(self.hidden)
        p = _lrelu(F.layer_norm(F.linear(p, self.p1_w, self.p1_b), (512,), self.ln1_w, self.ln1_b))
        p = _lrelu(F.layer_norm(F.linear(p, self.p2_w, self.p2_b), (32,), self.ln2_w, self.ln2_b))
        return p                                                     # [G,T,32]

    @torch.no_grad()
    def q_values(self, ob, node_feat):
        batch = torch.from_numpy(ob.batch)
        emb = self.node_embeddings(node_feat)
        P = self.project()
        inp = torch.cat([P[batch], emb], -1)                         # [N,T,64]
        a = F.linear(inp, self.q1_w, self.q1_b)
        a = _lrelu(F.layer_norm(a, (64,), self.qln_w, self.qln_b))
        a = F.linear(a, self.q2_w, self.q2_b)                        # [N,T,1]
        v = F.linear(_lrelu(F.linear(torch.tanh(P), self.v1_w, self.v1_b)), self.v2_w, self.v2_b)  # [G,T,1]
        return (a + v[batch]).squeeze(-1), emb, P

    # a9 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def select(self, ob, q, emb, tau=0.0, gen=None, forced_actions=None):
        """Returns LOCAL actions [G,T] (int64) and caches the chosen embedding.
        compat=True reproduces quirk A: the gather uses the LOCAL index into the FLAT array
        (ecord.py:167,183 -> actors/_base.py:305-308)."""
        G, T = ob.G, q.size(1)
        if forced_actions is not None:
            act_local = torch.as_tensor(forced_actions, dtype=torch.long).clone()
        else:
            act_glob = torch.empty(G, T, dtype=torch.long)
            if tau and tau > 0:
                u = torch.rand(q.shape, generator=gen)
            for g in range(G):
                lo, hi = int(ob.ptr[g]), int(ob.ptr[g + 1])
                if not tau:
                    act_glob[g] = _first_argmax(q[lo:hi]) + lo   # scatter_max tie -> lowest index
                else:
                    x = q[lo:hi] / tau
                    x = x - x.max(0).values
                    z = -(u[lo:hi]).log() / x.exp()
                    act_glob[g] = _first_argmin(z) + lo
            act_local = act_glob - torch.from_numpy(ob.ptr[:-1]).view(-1, 1)
        idx = act_local if self.compat else act_local + torch.from_numpy(ob.ptr[:-1]).view(-1, 1)
        self.last_sel = emb.gather(0, idx.unsqueeze(-1).expand(-1, -1, emb.size(-1))).clone()
        return act_local


def _first_argmax(x):
    m = x.max(0).values
    n = x.size(0)
    pos = torch.arange(n).view(-1, 1).expand_as(x)
    return torch.where(x == m, pos, torch.full_like(pos, n)).min(0).values


def _first_argmin(x):
    m = x.min(0).values
    n = x.size(0)
    pos = torch.arange(n).view(-1, 1).expand_as(x)
    return torch.where(x == m, pos, torch.full_like(pos, n)).min(0).values


# --------------------------------------------------------------------------------------------
# a10: rollout; a11: test_solver
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def rollout(model, ob, num_tradj, num_steps, init_state, tau=0.0, forced_actions=None, use_ref_env=False,
            record=False, seed=None):
    """DQNSolver.rollout (solver.py:196-377) for test mode (no epsilon, no callbacks).
    num_steps<0 -> |k|*n_g per graph and the loop runs to the max (solver.py:306-316).
    forced_actions: optional [S,G,T] local actions (teacher forcing for parity checks).
    record=True additionally returns per-step q / hidden / P / last_sel for parity tests."""
    log = defaultdict(list)
    model.initialise_embeddings(ob, num_tradj)
    env = EnvOracle(ob, num_tradj, init_state, use_ref=use_ref_env)
    log['init_score'] = torch.from_numpy(env.scores.copy())
    if num_steps < 0:
        S = int(abs(num_steps) * ob.num_nodes_list.max())
    else:
        S = int(num_steps)
    gen = None
    if tau and tau > 0:
        gen = torch.Generator().manual_seed(0 if seed is None else seed)
    node_feat, glob = env.observe()
    for t in range(S):
        model.update_internal_state(glob)
        q, emb, P = model.q_values(ob, node_feat)
        fa = None if forced_actions is None else forced_actions[t]
        actions = model.select(ob, q, emb, tau=tau, gen=gen, forced_actions=fa)
        if record:
            log['q'].append(q.clone())
            log['hidden'].append(model.hidden.clone())
            log['P'].append(P.clone())
            log['last_sel'].append(model.last_sel.clone())
        env.step(actions.numpy())
        node_feat, glob = env.observe()
        log['actions'].append(actions.clone())
        log['scores'].append(torch.from_numpy(env.scores.copy()))
    log['best_state'] = env.best_states_batched()
    log['best_scores'] = torch.from_numpy(env.best_scores.copy())
    log['env'] = env
    return log


def summarize(scores_list):
    """test_solver's result semantics (testing.py:201-204) for one rollout batch:
    scores [G,T,S] -> (scores_max[G], scores_mean[G]); the t=0 score is not part of the max."""
    s = torch.stack(scores_list, -1)
    by_tradj = s.max(-1).values
    return by_tradj.max(-1).values, by_tradj.mean(-1)


@torch.no_grad()
def test_solver(model, graphs, num_steps, num_tradj, graphs_bsz, init_state_fn, tau=0.0, use_ref_env=False):
    """testing.py:146-230 (graph chunking only; trajectory chunking concatenates on dim 1 the same way).
    init_state_fn(chunk_idx, N, T) -> int8 [N,T]."""
    smax, smean, init = [], [], []
    t_step = 0.0
    for i in range(0, len(graphs), graphs_bsz):
        ob = build_batch(graphs[i:i + graphs_bsz])
        log = rollout(model, ob, num_tradj, num_steps, init_state_fn(i // graphs_bsz, ob.N, num_tradj), tau=tau,
                      use_ref_env=use_ref_env)
        a, b = summarize(log['scores'])
        smax.append(a)
        smean.append(b)
        init.append(log['init_score'])
    return {'scores_max': torch.cat(smax), 'scores_mean': torch.cat(smean), 'init_score': torch.cat(init, 0)}
