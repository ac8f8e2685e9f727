"""Debug aid: engine vs oracle, teacher-forced, on a fixture's graphs (default config4_er2000): per step the GNN embedding error, the
Q error relative to each unit's Q spread, and whether the engine's own greedy choice equals the oracle's."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import ecord_oracle as orc
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts
from util import load_gold, unpack_graphs

fixture = sys.argv[1] if len(sys.argv) > 1 else "config4_er2000.npz"
gemm = sys.argv[2] if len(sys.argv) > 2 else "fp32"
S = int(sys.argv[3]) if len(sys.argv) > 3 else 6
z = load_gold(fixture)
graphs = unpack_graphs(z)
T = min(int(z["T"]), 8)
sd = random_state_dicts(int(z["weight_seed"]))
ob = orc.build_batch(graphs)
init = np.random.RandomState(int(z["state_seed"])).randint(0, 2, (ob.N, int(z["T"])), dtype=np.int8)[:, :T].copy()
model = orc.ModelOracle(sd)
model.initialise_embeddings(ob, T)
env = orc.EnvOracle(ob, T, init)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, DeviceWeights(sd), T, gemm=gemm, max_steps=S + 1, dump_q=True)
st.gnn_embed(); st.env_init(init); torch.cuda.synchronize()
rel = lambda a, b: float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
print(f"{fixture} [{gemm}] N={ob.N} T={T}: gnn_out rel {rel(st.gnn_out.cpu().numpy(), model.gnn_out.numpy()):.2e}  "
      f"node_emb rel {rel(st.node_emb.cpu().numpy(), model.init_emb.numpy()):.2e}")
nf, gf = env.observe()
for s in range(S):
    model.update_internal_state(gf)
    q, emb, P = model.q_values(ob, nf)
    a = model.select(ob, q, emb)
    st.policy_step(); st.q_select(forced_actions=a.numpy().astype(np.int32).reshape(-1)); torch.cuda.synchronize()
    qg = st.q_values().cpu()
    h = st.hidden[(st.steps_done + 1) & 1, :st.M].view(ob.G, T, 1024).cpu()
    worst, same = 0.0, True
    for g in range(ob.G):
        lo, hi = int(ob.ptr[g]), int(ob.ptr[g + 1])
        spread = (q[lo:hi].max(0).values - q[lo:hi].min(0).values)
        worst = max(worst, float(((qg[lo:hi] - q[lo:hi]).abs().max(0).values / spread).max()))
        same &= bool((orc._first_argmax(qg[lo:hi]) == a[g]).all())
    print(f"step {s}: hidden rel {rel(h.numpy(), model.hidden.numpy()):.2e}  P rel {rel(st.P[:st.M].view(ob.G, T, 32).cpu().numpy(), P.numpy()):.2e}  "
          f"q rel {rel(qg.numpy(), q.numpy()):.2e}  worst |dq| / spread {worst:.2e}  engine argmax == oracle {same}   "
          f"last_sel rel {rel(st.last_sel[:st.M].view(ob.G, T, 32).cpu().numpy(), model.last_sel.numpy()):.2e}")
    st.env_step(); env.step(a.numpy()); nf, gf = env.observe()
    assert np.array_equal(st.scores().cpu().numpy(), env.scores), "scores differ"
