"""Debug aid: per-kernel durations of the t=0 GNN embedding and the env set-up (CUPTI records via torch.profiler)."""
import os, sys, json, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from bench import make_graphs
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.network import random_state_dicts
workload = sys.argv[1] if len(sys.argv) > 1 else "er500"
graphs, n, G, T, desc = make_graphs(workload)
w = DeviceWeights(random_state_dicts(0))
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, w, T, max_steps=64)
init = np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8)
st.gnn_embed(); st.env_init(init); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st.gnn_embed(); st.env_init(init)
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", f"gnn_trace_{workload}.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
agg = collections.OrderedDict()
for e in sorted(ev, key=lambda e: e["ts"]):
    m = re.search(r'(k_[a-z0-9_]+)', e["name"])
    nm = m.group(1) if m else e["name"][:40]
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for nm, (k, d) in agg.items():
    print(f"{nm:40s} x{k // 3:3d} per call  {d / k:8.1f} us each  {d / 3:8.1f} us per call")
