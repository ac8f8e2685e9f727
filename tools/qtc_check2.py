"""Debug aid: tensor-core Q kernel vs the exact kernel at bench scale."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from bench import make_graphs
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts
from util import rel_err

workload = sys.argv[1] if len(sys.argv) > 1 else "er500"
gemm = sys.argv[2] if len(sys.argv) > 2 else "bf16"
graphs, n, G, T, desc = make_graphs(workload)
sd = random_state_dicts(0)
w = DeviceWeights(sd)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
init = np.random.RandomState(1234).randint(0, 2, (dg.N, T)).astype(np.int8)
a = RolloutState(dg, w, T, gemm=gemm, q_mode="exact", dump_q=True)
b = RolloutState(dg, w, T, gemm=gemm, q_mode="tc", dump_q=True)
for st in (a, b):
    st.gnn_embed(); st.env_init(init)
NS = int(sys.argv[3]) if len(sys.argv) > 3 else 6
for s in range(NS):
    for st in (a, b):
        st.policy_step(); st.q_select()
    qa, qb = a.q_values(), b.q_values()
    torch.cuda.synchronize()
    d = (qb - qa).abs()
    if s < 6 or s % 25 == 0: print(f"step {s}: tc vs exact rel {rel_err(qb.cpu().numpy(), qa.cpu().numpy()):.3e} |q|max {float(qa.abs().max()):.3f} "
          f"argmax diffs {int((a.actions != b.actions).sum())} of {a.M}; rows with err>1e-2: {int((d > 1e-2 * qa.abs().max()).sum())}; nan {int(torch.isnan(qb).sum())}")
    if s == 0:
        bad = (d > 1e-2 * qa.abs().max()).nonzero()
        print("  first bad (vertex, tradj):", bad[:8].tolist())
    b.actions.copy_(a.actions); b.last_sel.copy_(a.last_sel)
    for st in (a, b):
        st.env_step()
print("scores equal:", torch.equal(a.score, b.score))
