"""Graph inputs for the rollout path: file loaders and the synthetic sets BASELINE.json's configs name.

Loaders mirror `ecord/graphs/utils.py:9-52` (pickled lists of dense ndarrays / scipy CSR / networkx graphs
with `opts/cuts_<name>.pkl` best-known cuts beside them, and GSet .mtx files) without the networkx-2 calls
that no longer exist (`nx.from_scipy_sparse_matrix`).  Graphs are returned in the formats
`data.to_edge_list` accepts; nothing here needs networkx unless the pickle itself holds networkx graphs.

Generators follow the recipes in SURVEY.md 8d, which restate `ecord/graphs/generators.py:60-62,78-98`:
Erdos-Renyi G(n,p) with +-1 ("SIGNED") or unit weights.  They build EdgeListGraph directly with numpy
(seeded, machine-independent), since the reference's graph files for ER500 / ER2000..ER10000 are not shipped.
"""
import os
import pickle
import re

import numpy as np

from .data import EdgeListGraph


def load_graphs(graph_dir, graph_name, quiet=False):
    with open(os.path.join(graph_dir, graph_name + ".pkl"), 'rb') as f:
        graphs = pickle.load(f)
    try:
        with open(os.path.join(graph_dir, "opts", f"cuts_{graph_name}.pkl"), 'rb') as f:
            sols = pickle.load(f)
    except FileNotFoundError:
        sols = None
    if not quiet:
        print(f'{len(graphs)} target graphs loaded from {graph_dir}/{graph_name}.pkl')
    return list(graphs), sols


def load_gset_graph(fname):
    import scipy.io
    if os.path.splitext(fname)[-1] == '':
        fname += ".mtx"
    g = scipy.io.mmread(fname).tocsr().astype(np.float32)
    graph_dir, graph_name = os.path.split(fname)
    try:
        graph_id = int(re.findall('[0-9]+', graph_name)[0])
        with open(os.path.join(graph_dir, 'cuts.pkl'), 'rb') as f:
            cut = pickle.load(f)[graph_id - 1]
    except Exception:
        cut = None
    return [g], [cut]


def load_graph_from_file(fname, quiet=False):
    if 'gset' in fname:
        return load_gset_graph(fname)
    name = os.path.basename(fname)
    if name.endswith(".pkl"):
        name = name[:-4]
    return load_graphs(os.path.dirname(fname), name, quiet)


def _from_pairs(n, iu, ju, w):
    """undirected pairs (iu<ju) -> directed EdgeListGraph in reference order (source asc, neighbour asc)."""
    src = np.concatenate([iu, ju])
    dst = np.concatenate([ju, iu])
    ww = np.concatenate([w, w])
    order = np.lexsort((dst, src))
    return EdgeListGraph(int(n), src[order].astype(np.int64), dst[order].astype(np.int64), ww[order].astype(np.float32))


def erdos_renyi(n, p, seed, signed=True):
    """G(n,p); weights +-1 (EdgeType.SIGNED) or 1.  Dense sampling for small n, geometric skipping for large."""
    rs = np.random.RandomState(seed)
    if n <= 4096:
        iu, ju = np.triu_indices(n, 1)
        keep = rs.random_sample(iu.shape[0]) < p
        iu, ju = iu[keep], ju[keep]
    else:
        total = n * (n - 1) // 2
        m = rs.binomial(total, p)
        idx = np.unique(rs.randint(0, total, size=int(m * 1.02) + 16, dtype=np.int64))[:m]
        # linear index -> (i,j), i<j, row-major over the strict upper triangle
        i = (n - 2 - np.floor(np.sqrt(-8.0 * idx + 4.0 * n * (n - 1) - 7) / 2.0 - 0.5)).astype(np.int64)
        j = (idx + i + 1 - n * (n - 1) // 2 + (n - i) * ((n - i) - 1) // 2).astype(np.int64)
        iu, ju = i, j
    w = (2.0 * rs.randint(0, 2, iu.shape[0]) - 1.0) if signed else np.ones(iu.shape[0])
    return _from_pairs(n, iu, ju, w)


def barabasi_albert(n, m, seed, signed=True):
    """Preferential attachment with m edges per new vertex (power-law degrees, BASELINE config 3 shape)."""
    rs = np.random.RandomState(seed)
    targets = list(range(m))
    repeated = []
    iu, ju = [], []
    for src in range(m, n):
        for t in set(targets):
            iu.append(min(src, t)); ju.append(max(src, t))
        repeated.extend(set(targets))
        repeated.extend([src] * m)
        targets = set()
        while len(targets) < m:
            targets.add(repeated[rs.randint(0, len(repeated))])
        targets = list(targets)
    iu, ju = np.asarray(iu, np.int64), np.asarray(ju, np.int64)
    w = (2.0 * rs.randint(0, 2, iu.shape[0]) - 1.0) if signed else np.ones(iu.shape[0])
    return _from_pairs(n, iu, ju, w)


def synthetic_set(kind, n, num_graphs, avg_degree=None, p=None, signed=True, seed0=0):
    """BASELINE.json config shapes: ('er', 500, 100, p=0.15) = ER500; ('er', 10000, 10, avg_degree=10) = ER10000;
    ('ba', 500, 50) = BA500 (m=4)."""
    out = []
    for s in range(num_graphs):
        if kind == 'er':
            pp = p if p is not None else avg_degree / (n - 1)
            out.append(erdos_renyi(n, pp, seed0 + s, signed))
        elif kind == 'ba':
            out.append(barabasi_albert(n, 4, seed0 + s, signed))
        else:
            raise ValueError(kind)
    return out
