"""Experiment: k independent RolloutStates (trajectory slices of the same batch) on k streams, timed together.

    python tools/split_time.py er500 2

The kernels of one step are a serial chain; two half-batches on two streams let the block scheduler fill the
drain / single-wave phases of one chain with the other chain's kernels."""
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
graphs, n, G, T, desc = make_graphs(workload)
sd = random_state_dicts(0)
w = DeviceWeights(sd)
dg = DeviceGraph(GraphBatcher()(graphs).tg)
init = np.random.RandomState(1).randint(0, 2, (dg.N, T), dtype=np.int8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in [int(x) for x in sys.argv[2:]] or [1, 2]:
    bounds = [T * i // k for i in range(k + 1)]
    sts = []
    for i in range(k):
        st = RolloutState(dg, w, bounds[i + 1] - bounds[i], gemm="bf16", max_steps=4000)
        st.gnn_embed()
        st.env_init(np.ascontiguousarray(init[:, bounds[i]:bounds[i + 1]]))
        sts.append(st)
    for st in sts:
        st.run(16)
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    for do_flush in (False, True):
        tot = 0.0
        for _ in range(8):
            if do_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for st in sts:
                st.stream.wait_stream(cur)
            for c in range(4):
                for st in sts:
                    with torch.cuda.stream(st.stream):
                        st.run(16)
            for st in sts:
                cur.wait_stream(st.stream)
            b.record(); b.synchronize()
            tot += a.elapsed_time(b)
        print(f"{workload} split={k} flush={do_flush}: {tot / (8 * 64) * 1e3:.1f} us/step "
              f"({G * T / (tot / (8 * 64) * 1e-3) / 1e6:.1f} M steps/s)", flush=True)
    best = torch.cat([st.best_score.view(G, -1) for st in sts], 1)
    print("  mean best", float(best.max(1).values.mean()))
    for st in sts:
        st.close()
    del sts
