"""Debug aid: time the CUDA-graph rollout loop with / without the L2 flush between chunks."""
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
q_mode = sys.argv[2] if len(sys.argv) > 2 else "tc"
graphs, n, G, T, desc = make_graphs(workload)
sd = random_state_dicts(0)
w = DeviceWeights(sd)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, w, T, gemm="bf16", q_mode=q_mode, max_steps=2000)
st.gnn_embed()
st.env_init(np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8))
st.run(16)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for do_flush in (False, True, False):
    tot = 0.0
    for _ in range(8):
        if do_flush:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.run(16); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    print(f"{q_mode}: flush={do_flush}: {tot / 128 * 1e3:.1f} us/step")
for nm, fn in (("policy_step", st.policy_step), ("q_select", st.q_select)):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record(); b.synchronize()
    print(f"  {nm}: {a.elapsed_time(b) / 10 * 1e3:.1f} us")
