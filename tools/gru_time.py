"""Debug aid: time the GRU tcgen05 kernel alone (ECORD_TC_DEBUG / ECORD_GEMM_PAIR select variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
from bench import make_graphs
from ecord_b200 import _lib
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts
graphs, n, G, T, desc = make_graphs(sys.argv[1] if len(sys.argv) > 1 else "er500")
w = DeviceWeights(random_state_dicts(0))
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, w, T, gemm="bf16", max_steps=64)
st.gnn_embed(); st.env_init(np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8))
st.run(4); torch.cuda.synchronize()
# whole policy step minus the other kernels is not separable through the ABI: time policy_step and report
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): st.policy_step()
torch.cuda.synchronize(); a.record()
for _ in range(20): st.policy_step()
b.record(); b.synchronize()
print(f"ECORD_TC_DEBUG={os.environ.get('ECORD_TC_DEBUG','0')} PAIR={os.environ.get('ECORD_GEMM_PAIR','1')}: policy_step {a.elapsed_time(b)/20*1e3:.1f} us")
