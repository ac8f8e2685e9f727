"""BASELINE.json configs[4]: throughput sweep over synthetic GSet-shaped sparse graphs (|V| x average degree), one GPU.

    python tools/gset_sweep.py [--steps 200] [--gemm bf16x3] [--out profiles/r02_gset_sweep.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gset_sweep.py   # N GPUs, weak scaling:
                                                                  # every rank runs 128 more trajectories of the same graphs

Graphs: G(n, deg/(n-1)), weights +1 (and one +-1 variant), seeds 0..3 (the real GSet files are not
downloadable here), 128 trajectories per graph, greedy, throughput mode.  Per configuration: device-resident rollout steps/s
(CUDA events around `steps` fused steps after 32 warm-up steps) and the CPU rollout of the reference (compiled Cython step +
torch CPU model, oracle port) on a bounded sample when --cpu is given."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ecord_b200.data import GraphBatcher
from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_gset_sweep.json"))
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--gemm", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
args = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T, NG = 128, 4
sd = random_state_dicts(0)
w = DeviceWeights(sd)
rows = []
for n in (800, 2000, 5000, 10000):
    for deg, signed in ((4, False), (12, False), (40, False), (12, True)):
        graphs = synthetic_set("er", n, NG, avg_degree=deg, signed=signed)
        dg = DeviceGraph(GraphBatcher()(graphs).tg)
        st = RolloutState(dg, w, T, gemm=args.gemm, max_steps=32 + args.steps)
        st.gnn_embed()
        st.env_init(np.random.RandomState(1 + rank).randint(0, 2, (dg.N, T), dtype=np.int8))
        st.run(32)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.run(args.steps); b.record(); b.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:                                   # max over ranks
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        rate = NG * T * world * args.steps / (ms * 1e-3)
        row = {"n": n, "avg_degree": deg, "weights": "+-1" if signed else "+1", "graphs": NG, "tradj": T * world, "gpus": world, "steps": args.steps,
               "us_per_step": ms / args.steps * 1e3, "steps_per_s": rate,
               "vertex_rows_per_s": rate * n, "best_cut_mean": float(st.best_score.view(NG, T).max(1).values.mean())}
        if args.cpu and rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from bench import cpu_rollout_rate
            cr, dt, _ = cpu_rollout_rate(graphs[:1], T, 8)
            row["cpu_steps_per_s"] = cr
        rows.append(row)
        if rank == 0:
            print(f"n={n:6d} deg={deg:3d} {row['weights']:>3s}: {row['us_per_step']:7.1f} us/step  {rate / 1e6:6.2f} M steps/s  "
                  f"{rate * n / 1e9:6.1f} G vertex-rows/s" + (f"  cpu {row['cpu_steps_per_s']:.0f} steps/s" if args.cpu else ""), flush=True)
        st.close()
        del st, dg
if rank == 0:
    with open(args.out, "w") as f:
        json.dump({"what": f"GSet-shaped sweep, {world} x B200, gemm={args.gemm}", "rows": rows}, f, indent=1)
if world > 1:
    dist.destroy_process_group()
