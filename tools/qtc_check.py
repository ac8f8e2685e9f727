"""Debug aid: tensor-core Q kernel vs the exact kernel, printed per step (run on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from util import rel_err

graphs = synthetic_set("er", 300, 2, p=0.05) + synthetic_set("er", 129, 1, p=0.1, seed0=5) + synthetic_set("er", 40, 2, p=0.2, seed0=9)
sd = random_state_dicts(21, perturb_norms=True)
w = DeviceWeights(sd)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
T = 37
init = np.random.RandomState(3).randint(0, 2, (dg.N, T)).astype(np.int8)
a = RolloutState(dg, w, T, gemm="fp32", q_mode="exact", dump_q=True)
b = RolloutState(dg, w, T, gemm="fp32", q_mode="tc", dump_q=True)
c = RolloutState(dg, w, T, gemm="fp32", q_mode="fast", dump_q=True)
for st in (a, b, c):
    st.gnn_embed(); st.env_init(init)
for s in range(6):
    for st in (a, b, c):
        st.policy_step(); st.q_select()
    qa, qb, qc = a.q_values(), b.q_values(), c.q_values()
    torch.cuda.synchronize()
    print(f"step {s}: tc vs exact rel {rel_err(qb.cpu().numpy(), qa.cpu().numpy()):.3e}  fast vs exact {rel_err(qc.cpu().numpy(), qa.cpu().numpy()):.3e}  "
          f"|q|max {float(qa.abs().max()):.3f}  argmax diffs tc {int((a.actions != b.actions).sum())} fast {int((a.actions != c.actions).sum())} of {a.M}")
    if s == 0:
        d = (qb - qa).abs()
        i = int(d.argmax())
        print("  worst at vertex", i // T, "tradj", i % T, "exact", float(qa.flatten()[i]), "tc", float(qb.flatten()[i]))
        print("  first row exact", qa[0, :4].tolist(), "tc", qb[0, :4].tolist())
    for st in (b, c):
        st.actions.copy_(a.actions); st.last_sel.copy_(a.last_sel)
    for st in (a, b, c):
        st.env_step()
print("scores equal:", torch.equal(a.score, b.score))
