"""Debug aid: time one policy kernel alone (GRU by default) under the diagnostic switches of gemm_tc.cu.

    ECORD_TC_DEBUG=<bits> ECORD_TC_CLUSTERS=<n> python tools/gru_time.py [workload] [gemm] [mask]

bits: 1 no MMA issue, 2 no epilogue math, 8 prologue + teardown only, 16 no collector reuse (see TcParams.debug);
mask: ECORD_POLICY_MASK (1 GRU, 2 projection, 4 tail).  The switches are read once per process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_graphs
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts
graphs, n, G, T, desc = make_graphs(sys.argv[1] if len(sys.argv) > 1 else "er500")
gemm = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
os.environ["ECORD_POLICY_MASK"] = sys.argv[3] if len(sys.argv) > 3 else "1"
w = DeviceWeights(random_state_dicts(0))
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, w, T, gemm=gemm, max_steps=64)
st.gnn_embed(); st.env_init(np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8))
if int(os.environ.get("ECORD_TC_DEBUG", "0")) == 0:
    st.run(4)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): st.policy_step()
torch.cuda.synchronize(); a.record()
for _ in range(20): st.policy_step()
b.record(); b.synchronize()
print(f"{sys.argv[1:]} ECORD_TC_DEBUG={os.environ.get('ECORD_TC_DEBUG','0')} CLUSTERS={os.environ.get('ECORD_TC_CLUSTERS','-')} "
      f"NARROW={os.environ.get("ECORD_GRU_NARROW","auto")}: {a.elapsed_time(b)/20*1e3:.1f} us per launch (mask {os.environ['ECORD_POLICY_MASK']}, warm L2)")
