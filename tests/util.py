"""Shared helpers for the test-suite: golden fixtures and graph reconstruction."""
import os

import numpy as np

from ecord_b200.data import EdgeListGraph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def unpack_graphs(z, prefix=""):
    """Inverse of oracle/gen_golden.py:pack_graphs -> list of EdgeListGraph (reference edge order)."""
    ns, ne = z[prefix + "g_n"], z[prefix + "g_ne"]
    src, dst, w = z[prefix + "g_src"], z[prefix + "g_dst"], z[prefix + "g_w"]
    out, o = [], 0
    for n, e in zip(ns, ne):
        out.append(EdgeListGraph(int(n), src[o:o + e].astype(np.int64), dst[o:o + e].astype(np.int64),
                                 w[o:o + e].astype(np.float32)))
        o += e
    return out


def oracle_batch(graphs):
    """EdgeListGraph list -> oracle.OracleBatch (the oracle accepts tuples directly)."""
    import ecord_oracle as orc
    return orc.build_batch(graphs)


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the 'relative' metric used for the fp32 / bf16 tolerances."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def spread_rel_err(q, q_ref, ptr):
    """Per-(graph, trajectory) error of a Q plane relative to the SPREAD of that unit's Q-values: the arg-max of a unit
    depends on differences within its column, so max|dq| / (max_n q - min_n q) is the figure that says whether an error
    can change the selection (a tensor-wide max-norm can hide it).  q, q_ref: [N, T]; ptr: [G+1] vertex offsets.
    Returns the maximum over all units."""
    q = np.asarray(q, dtype=np.float64)
    q_ref = np.asarray(q_ref, dtype=np.float64)
    worst = 0.0
    for g in range(len(ptr) - 1):
        a, b = q[ptr[g]:ptr[g + 1]], q_ref[ptr[g]:ptr[g + 1]]
        if a.shape[0] < 2:
            continue
        spread = np.maximum(b.max(0) - b.min(0), 1e-30)
        worst = max(worst, float((np.abs(a - b).max(0) / spread).max()))
    return worst
