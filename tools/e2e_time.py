"""Debug aid: where the end-to-end B200Solver.rollout wall time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_graphs
from ecord_b200.data import GraphBatcher
from ecord_b200.network import random_state_dicts
from ecord_b200.solver import B200Solver
import ecord_b200.solver as S_
graphs, n, G, T, desc = make_graphs(sys.argv[1] if len(sys.argv) > 1 else "er500")
S = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sd = random_state_dicts(0)
batch = GraphBatcher()(graphs)
init = np.random.RandomState(1).randint(0, 2, (batch.tg.num_nodes, T), dtype=np.int8)
solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], steps_per_graph=16)
import cProfile, pstats
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if rep == 2:
        pr = cProfile.Profile(); pr.enable()
    log = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)
    if rep == 2:
        pr.disable()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sc = torch.stack(log['scores'], -1).max(-1).values.cuda()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: rollout wall {1e3*(t1-t0):.1f} ms, post {1e3*(t2-t1):.1f} ms; t_init_emb {1e3*float(log['t_init_emb'][0]):.2f} t_setup {1e3*float(log['t_setup'][0]):.2f} "
          f"t_rollout {1e3*float(log['t_rollout']):.2f} ms ({1e3*float(log['t_rollout'])/S*1e3:.1f} us/step)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
