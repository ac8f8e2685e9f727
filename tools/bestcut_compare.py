"""Best-cut parity at the real rollout length: fp32 parity mode vs the throughput mode on ER500 / ER200 graphs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ecord_b200.data import GraphBatcher
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from ecord_b200.solver import B200Solver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
G = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = 50
graphs = synthetic_set("er", n, G, p=0.15, signed=True)
batch = GraphBatcher()(graphs)
init = np.random.RandomState(7).randint(0, 2, (batch.tg.num_nodes, T)).astype(np.int8)
sd = random_state_dicts(0)
res = {}
for gemm in ("fp32", "bf16"):
    s = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], gemm=gemm)
    log = s.rollout(batch, num_tradj=T, num_steps=-4, tau=0, use_epsilon=False, init_state=init)
    res[gemm] = log['best_score_by_tradj']
    print(f"{gemm}: {len(log['scores'])} steps, best cut per graph {res[gemm].max(1).values.tolist()}")
a, b = res["fp32"], res["bf16"]
print(f"ER{n} x{G} x{T}: best cut equal on {float((a.max(1).values == b.max(1).values).float().mean()):.2f} of graphs; "
      f"mean best {float(a.max(1).values.mean()):.1f} vs {float(b.max(1).values.mean()):.1f}; "
      f"mean over trajectories {float(a.mean()):.2f} vs {float(b.mean()):.2f}")
