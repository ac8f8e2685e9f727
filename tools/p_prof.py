"""Profiling driver: isolated launches of the policy step (GRU, PROJ, tail kernels) on a bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_graphs
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts

workload = sys.argv[1] if len(sys.argv) > 1 else "er500"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gemm = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
graphs, n, G, T, desc = make_graphs(workload)
sd = random_state_dicts(0)
batch = GraphBatcher()(graphs)
w = DeviceWeights(sd)
dg = DeviceGraph(batch.tg)
st = RolloutState(dg, w, T, gemm=gemm, max_steps=64)
st.gnn_embed()
st.env_init(np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8))
st.run(8)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.policy_step(); torch.cuda.synchronize()
a.record()
for _ in range(reps):
    st.policy_step()
b.record(); b.synchronize()
print(f"{desc} [{gemm}]: policy_step {a.elapsed_time(b) / reps * 1e3:.1f} us per step (3 kernels), rows {st.M}")
