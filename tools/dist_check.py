"""Multi-GPU check (torchrun, NCCL): trajectory sharding and graph sharding give the single-GPU rollout's per-trajectory best
scores on every rank, bit for bit.  Prints one line per rank-0 check and the wall time of each sharded call.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from ecord_b200.data import GraphBatcher
from ecord_b200.dist import rollout_graph_sharded, rollout_sharded
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts
from ecord_b200.solver import B200Solver

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sd = random_state_dicts(0)
solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'], device=torch.device("cuda", local))
for name, kind, n, G, T, S, kw in (("ER200 x100", "er", 200, 100, 48, 200, dict(p=0.15, signed=True)),
                                   ("ragged BA", "ba", 0, 0, 24, 120, dict(signed=True))):
    if n:
        graphs = synthetic_set(kind, n, G, seed0=0, **kw)
    else:
        graphs = [g for i, m in enumerate((300, 120, 500, 80, 260, 410, 64, 350, 222, 190, 470, 33)) for g in synthetic_set("ba", m, 1, seed0=i, **kw)]
    batch = GraphBatcher()(graphs)
    init = np.random.RandomState(11).randint(0, 2, (batch.tg.num_nodes, T), dtype=np.int8)
    ref = solver.rollout(batch, num_tradj=T, num_steps=S, tau=0, use_epsilon=False, init_state=init)['best_score_by_tradj']
    for mode, fn in (("trajectories", rollout_sharded), ("graphs", rollout_graph_sharded)):
        fn(solver, batch, T, S, init)                         # warm-up (graph capture, NCCL)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        log = fn(solver, batch, T, S, init)
        dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ok = torch.equal(log['scores_by_tradj'], ref)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            print(f"{name}: {mode:12s} sharding over {world} GPUs == single-GPU rollout on every rank: {all(flags)}   "
                  f"({len(graphs)} graphs x {T} tradj x {S} steps in {dt * 1e3:.1f} ms)", flush=True)
        assert ok
dist.destroy_process_group()
