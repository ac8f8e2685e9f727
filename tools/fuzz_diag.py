"""Debug aid: one case of tests/test_rollout_gpu.py::test_free_running_rollout_matches_oracle_on_random_shapes, verbose."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ecord_oracle as orc
from ecord_b200.data import EdgeListGraph, GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts

case = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ns, T = [([5, 1, 37], 3), ([130], 1), ([64, 200, 17, 129], 5), ([3] * 9, 2), ([260, 40], 7), ([128, 127], 4)][case]
graphs = []
for i, n in enumerate(ns):
    p = [0.5, 0.1, 0.25][i % 3]
    graphs += synthetic_set("er", n, 1, p=p, signed=bool((case + i) % 2), seed0=1000 * case + i) if n > 1 else \
        [EdgeListGraph(1, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))]
compat = case % 2 == 0
sd = random_state_dicts(20 + case, perturb_norms=True)
S = 14
ob = orc.build_batch(graphs)
init = np.random.RandomState(100 + case).randint(0, 2, (ob.N, T)).astype(np.int8)
model = orc.ModelOracle(sd, compat=compat)
ref = orc.rollout(model, ob, T, S, init, tau=0.0, record=True)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, DeviceWeights(sd), T, gemm="fp32", compat=compat, max_steps=S, dump_q=True)
st.gnn_embed(); st.env_init(init)
emb_ref = model.gnn_embed(ob)
print("gnn rel err", float((st.gnn_out.cpu() - emb_ref).abs().max() / emb_ref.abs().max()))
for s in range(S):
    st.policy_step(); st.q_select()
    q = st.q_values().cpu()
    qr = ref["q"][s]
    a = st.actions.view(dg.G, T).cpu().numpy()
    ar = ref["actions"][s].numpy()
    print(f"step {s}: q rel err {float((q - qr).abs().max() / qr.abs().max()):.2e}  hidden err "
          f"{float((st.hidden_state().cpu() - ref['hidden'][s]).abs().max()):.2e}  last_sel err "
          f"{float((st.last_sel[:st.M].view(dg.G, T, 32).cpu() - ref['last_sel'][s]).abs().max()):.2e}  actions equal {np.array_equal(a, ar)}")
    if not np.array_equal(a, ar):
        for g, t in zip(*np.nonzero(a != ar)):
            qq = qr[ob.ptr[g]:ob.ptr[g + 1], t]
            print("   unit", g, t, "gpu", a[g, t], "ref", ar[g, t], "ref q[gpu a]", float(qq[a[g, t]]), "ref q max", float(qq.max()),
                  "gpu q[gpu a]", float(q[ob.ptr[g] + a[g, t], t]), "gpu q[ref a]", float(q[ob.ptr[g] + ar[g, t], t]))
        break
    st.env_step()
