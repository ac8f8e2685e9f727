"""Graph batching without torch_geometric (loaders for the inputs).

Mirrors `GraphBatcher.__call__` -> `Batch(nx, tg, info)` of the reference
(`ecord/environment/data.py:14-17,53-94,114-128`): graphs are relabelled to integers in node order, made
directed, and their edge lists concatenated with cumulative vertex offsets.  Edge ORDER is the reference's
(source ascending, each source's neighbours in adjacency order) because the environment's CSR views
(`tradjectories.py:38-60`) and the summation order of the GNN's scatter_mean depend on it.

Accepted graph formats (what `ecord/graphs/utils.py:15-34` produces, plus a raw form):
  networkx.Graph / DiGraph, dense numpy adjacency [n,n], scipy.sparse matrix, or
  EdgeListGraph(n, src, dst, w) -- directed edge arrays already in reference order.
"""
from collections import namedtuple

import numpy as np
import torch

Batch = namedtuple('Batch', ['nx', 'tg', 'info'])
BatchInfo = namedtuple('BatchInfo', field_names=['num_nodes', 'batch_offset', 'batch_to_flat_idxs'])
EdgeListGraph = namedtuple('EdgeListGraph', ['n', 'src', 'dst', 'w'])


def to_edge_list(g):
    """One graph -> EdgeListGraph with DIRECTED edges (both directions of every undirected edge)."""
    if isinstance(g, EdgeListGraph):
        return g
    if isinstance(g, np.ndarray):
        # nx.from_numpy_array semantics (graphs/utils.py:19-21): vertex i's neighbours ascending
        a = np.asarray(g)
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("dense adjacency must be square")
        sym = (a != 0) | (a.T != 0)
        src, dst = np.nonzero(sym)
        w = np.where(a[src, dst] != 0, a[src, dst], a[dst, src]).astype(np.float32)
        # an undirected networkx graph keeps the LAST value assigned to (i,j); from_numpy_array visits
        # (i,j) row-major, so for asymmetric input the lower-triangle value wins.
        lo = np.where(src > dst, a[src, dst], a[dst, src])
        w = np.where(lo != 0, lo, w).astype(np.float32)
        return EdgeListGraph(a.shape[0], src.astype(np.int64), dst.astype(np.int64), w)
    try:
        import scipy.sparse as sp
        if sp.issparse(g):
            m = sp.coo_matrix(g)
            m = (m + m.T).tocsr() if (abs(m - m.T)).nnz else m.tocsr()
            m.sort_indices()
            src = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
            return EdgeListGraph(m.shape[0], src.astype(np.int64), m.indices.astype(np.int64), m.data.astype(np.float32))
    except ImportError:
        pass
    # networkx: GraphData.from_networkx (data.py:114-128) = convert_node_labels_to_integers + to_directed + G.edges
    index = {u: i for i, u in enumerate(g.nodes())}
    src, dst, w = [], [], []
    directed = g.is_directed()
    has_w = None
    for u, nbrs in g.adjacency():
        iu = index[u]
        for v, d in nbrs.items():
            src.append(iu)
            dst.append(index[v])
            ww = d.get('weight')
            if has_w is None:
                has_w = ww is not None
            w.append(1.0 if ww is None else ww)
    del directed
    return EdgeListGraph(len(index), np.asarray(src, np.int64), np.asarray(dst, np.int64), np.asarray(w, np.float32))


class TgBatch:
    """The fields of the reference's collated `Batch_tg` that the rollout path reads
    (`data.py:53-69`): edge_index i64[2,E], edge_attr f32[E,1], batch i64[N], ptr i64[G+1],
    degree i64[N], x f32[N,1] (ones), num_nodes; plus numpy views for the device upload."""

    def __init__(self, els):
        ns = np.asarray([e.n for e in els], np.int64)
        offs = np.concatenate([[0], np.cumsum(ns)])
        self.num_nodes = int(offs[-1])
        self.num_graphs = len(els)
        self.num_nodes_list = ns
        src = np.concatenate([e.src + o for e, o in zip(els, offs[:-1])]) if els else np.zeros(0, np.int64)
        dst = np.concatenate([e.dst + o for e, o in zip(els, offs[:-1])]) if els else np.zeros(0, np.int64)
        w = np.concatenate([e.w for e in els]).astype(np.float32)
        if np.any(np.diff(src) < 0):
            raise ValueError("edges must be grouped by ascending source vertex (reference edge order)")
        # the env kernels update the neighbours of a flipped vertex in parallel: a self-loop or a repeated (src, dst) pair
        # would make two lanes read-modify-write the same entry (the sequential reference has no such hazard)
        if np.any(src == dst):
            raise ValueError("self-loops are not supported (Max-Cut on simple graphs)")
        if src.size:
            key = src * np.int64(self.num_nodes) + dst
            if np.unique(key).size != key.size:
                raise ValueError("duplicate (src, dst) edges are not supported (Max-Cut on simple graphs)")
        self.np_src, self.np_dst, self.np_w = src, dst, w
        self.np_ptr = offs.astype(np.int64)
        self.np_edge_counts = np.asarray([len(e.w) for e in els], np.int64)
        self.edge_index = torch.from_numpy(np.stack([src, dst]))
        self.edge_attr = torch.from_numpy(w).unsqueeze(1)
        self.batch = torch.from_numpy(np.repeat(np.arange(len(els)), ns))
        self.ptr = torch.from_numpy(self.np_ptr.copy())
        self.degree = torch.from_numpy(np.bincount(src, minlength=self.num_nodes).astype(np.int64))
        self.x = torch.ones(self.num_nodes, 1)
        self.u = None

    @property
    def num_edges(self):
        return self.np_src.shape[0]

    def select(self, graph_ids):
        """A new batch made of the given graphs of this one, in the given order (graph sharding, dist.py)."""
        ptr, epos = self.np_ptr, np.concatenate([[0], np.cumsum(self.np_edge_counts)])
        els = []
        for g in graph_ids:
            e0, e1 = int(epos[g]), int(epos[g + 1])
            els.append(EdgeListGraph(int(ptr[g + 1] - ptr[g]), self.np_src[e0:e1] - ptr[g], self.np_dst[e0:e1] - ptr[g],
                                     self.np_w[e0:e1]))
        return TgBatch(els)

    def vertex_index(self, graph_ids):
        """int64 [n]: the vertices of the given graphs (in that order) as indices into this batch."""
        ptr = self.np_ptr
        return np.concatenate([np.arange(ptr[g], ptr[g + 1]) for g in graph_ids]) if len(graph_ids) else np.zeros(0, np.int64)

    @classmethod
    def from_reference_batch(cls, tg):
        """A batch collated by the REFERENCE's own GraphBatcher (a torch_geometric `Batch`: edge_index i64[2,E] with
        cumulative vertex offsets, edge_attr [E,1], ptr i64[G+1]; environment/data.py:53-69) -> TgBatch.  This is what lets
        `ecord.validate.testing.test_solver` drive the engine with the reference's own batcher."""
        ptr = np.asarray(tg.ptr.cpu().numpy(), np.int64)
        ei = np.asarray(tg.edge_index.cpu().numpy(), np.int64)
        w = np.asarray(tg.edge_attr.detach().cpu().numpy(), np.float32).reshape(-1)
        gid = np.searchsorted(ptr, ei[0], side="right") - 1          # graph of every edge (by its source vertex)
        els = []
        for g in range(len(ptr) - 1):
            m = gid == g
            els.append(EdgeListGraph(int(ptr[g + 1] - ptr[g]), ei[0][m] - ptr[g], ei[1][m] - ptr[g], w[m]))
        return cls(els)


def as_tg_batch(tg):
    """TgBatch as is; a torch_geometric-style batch (the reference's) converted once and cached on the object."""
    if isinstance(tg, TgBatch):
        return tg
    cached = getattr(tg, "_ecord_b200_tg", None)
    if cached is None:
        cached = TgBatch.from_reference_batch(tg)
        try:
            tg._ecord_b200_tg = cached
        except Exception:
            pass
    return cached


def get_batch_info(batch, num_nodes_list):
    """`get_batch_info` of the reference (data.py:85-94)."""
    num_nodes_list = torch.as_tensor(num_nodes_list)
    batch_offset = torch.roll(torch.cumsum(num_nodes_list, 0), 1)
    batch_offset[0] = 0
    batch_to_flat_idxs = (batch, torch.cat([torch.arange(int(n)) for n in num_nodes_list]))
    return BatchInfo(num_nodes_list, batch_offset, batch_to_flat_idxs)


class GraphBatcher:
    """Drop-in for `ecord.environment.data.GraphBatcher` for the canonical configuration (all the
    optional node profiles / Laplacian PEs are off in experiments/train.py:57-60, and are rejected here)."""

    def __init__(self, add_normalised_degree=False, add_degree_profile=False, add_weight_profile=False,
                 norm_features=False, add_laplacian_PEs=False):
        if any([add_normalised_degree, add_degree_profile, add_weight_profile, norm_features, add_laplacian_PEs]):
            raise NotImplementedError("node profiles / Laplacian PEs are outside the canonical ECORD configuration")

    def __call__(self, batch_nx, return_tg_only=False):
        els = [to_edge_list(g) for g in batch_nx]
        tg = TgBatch(els)
        if return_tg_only:
            return Batch(tg=tg, nx=None, info=None)
        return Batch(list(batch_nx), tg, get_batch_info(tg.batch, [e.n for e in els]))
