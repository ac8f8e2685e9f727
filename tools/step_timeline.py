"""Where a rollout step spends its time INSIDE the CUDA-graph plan (programmatic dependent launch and all): kernel start /
end times from CUPTI activity records (torch.profiler), not from serialised launches.

    python tools/step_timeline.py [workload] [gemm] [steps] [trajectories per graph]

Prints, per kernel of the step: mean duration, mean start offset within the step, the gap (or overlap) to the previous kernel's
end, and the step period.  A number taken under the tracer is not a bench value: use it for shares and gaps only."""
import os, sys, json, collections
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
gemm = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
graphs, n, G, T, desc = make_graphs(workload)
if len(sys.argv) > 4:
    T = int(sys.argv[4])
w = DeviceWeights(random_state_dicts(0))
dg = DeviceGraph(GraphBatcher()(graphs).tg)
st = RolloutState(dg, w, T, gemm=gemm, max_steps=steps + 64)
st.gnn_embed()
st.env_init(np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8))
st.run(32)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st.run(steps)
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", f"timeline_{workload}_{gemm}_T{T}.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
import re
def short(nm):
    m = re.search(r"(k_[a-z0-9_]+(<[^>]*>)?)", nm)
    return m.group(1) if m else nm[:28]
# split into steps at every GRU kernel (the first kernel of a step)
first = next(i for i, e in enumerate(ev) if "k_gemm_tc2<0" in e["name"] or "k_sgemm" in e["name"])
ev = ev[first:]
rows = collections.OrderedDict()
step_starts = []
t0 = None
prev_end = None
for e in ev:
    nm = short(e["name"])
    if "k_gemm_tc2<0" in e["name"]:
        t0 = e["ts"]; step_starts.append(t0)
    r = rows.setdefault(nm, dict(n=0, dur=0.0, off=0.0, gap=0.0, end=0.0))
    r["n"] += 1; r["dur"] += e["dur"]; r["off"] += e["ts"] - t0; r["end"] += e["ts"] + e["dur"] - t0
    if prev_end is not None:
        r["gap"] += e["ts"] - prev_end
    prev_end = e["ts"] + e["dur"]
period = (step_starts[-1] - step_starts[0]) / (len(step_starts) - 1)
print(f"{workload} {gemm}: {len(step_starts)} steps traced, step period {period:.1f} us (under the tracer)")
print(f"{'kernel':30s} {'launches':>8s} {'dur us':>8s} {'start':>8s} {'end':>8s} {'gap to prev end':>16s}")
for nm, r in rows.items():
    k = r["n"]
    print(f"{nm:30s} {k:8d} {r['dur'] / k:8.1f} {r['off'] / k:8.1f} {r['end'] / k:8.1f} {r['gap'] / k:16.1f}")
