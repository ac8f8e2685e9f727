#!/usr/bin/env python
"""bench.py -- ECORD rollout throughput on B200: rollout steps/sec = graphs x trajectories x steps / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload er500|er10000|ba500|er40] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (GRU update + projection, Q over every vertex, selection, env flip)
over one rollout batch.  Workload at N=1 = BASELINE.json configs[1]: ER500 (100 graphs G(500,0.15), +-1
weights, synthesised as SURVEY.md 8d prescribes because the reference's file is not shipped), 50
trajectories per graph, greedy tau=0; all 100 graphs form one rollout batch (5000 units per step).
N>1: weak scaling -- every rank runs 50 more trajectories of the same graphs (trajectory sharding,
ecord_b200/dist.py); the only collective is the final best-cut reduce, inside the timed region.

Prints ONE JSON line (rank 0).  `value`: device-resident throughput (CUDA events, max over ranks).
`e2e`: the same metric through the public API (B200Solver.rollout) with host inputs/outputs.
`roofline`: the dominant kernel timed alone with CUDA events.  `cpu_baseline`: the oracle port of the
reference rollout on the host cores, bounded sample.  `--impl reference` times that CPU path as the
reference arm (Cython env step from oracle/_ref when present + torch CPU model restated in oracle/).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (kind, n, num_graphs, gen kwargs, T per GPU, description)
    "er500": ("er", 500, 100, dict(p=0.15, signed=True), 50, "ER500: 100 x G(500,0.15) +-1 weights, 50 tradj/graph/GPU, greedy"),
    "er200": ("er", 200, 100, dict(p=0.15, signed=True), 50, "ER200: 100 x G(200,0.15) +-1 weights, 50 tradj/graph/GPU, greedy"),
    "ba500": ("ba", 500, 50, dict(signed=True), 50, "BA500: 50 x BA(500,m=4) +-1 weights, 50 tradj/graph/GPU, greedy"),
    "er40": ("er", 40, 100, dict(p=0.15, signed=True), 50, "ER40: 100 x G(40,0.15) +-1 weights, 50 tradj/graph/GPU, greedy"),
    "er10000": ("er", 10000, 10, dict(avg_degree=10, signed=False), 128, "ER10000: 10 x G(10000, deg 10) unit weights, 128 tradj/graph/GPU"),
    "er2000": ("er", 2000, 10, dict(avg_degree=10, signed=False), 128, "ER2000: 10 x G(2000, deg 10) unit weights, 128 tradj/graph/GPU"),
}
F_UNIT = lambda n: 7772480 + 8416 * n            # literal FLOPs per (g,t,step), SURVEY.md 8d
B_UNIT = lambda n, d: 12 * n + 20 * d + 8268     # compulsory bytes per (g,t,step), SURVEY.md 8d


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is polled
    in-process (pynvml, initialised before the warm-up so that its start-up cost stays outside the timed region);
    `nvidia-smi -lms` in a subprocess is the fallback.  Only samples taken between mark_begin() and mark_end()
    are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.01):
        self.rows, self.idx, self.period = [], gpu_index, period_s
        self.t0 = self.t1 = None
        self.stop_flag = False
        self.mode, self.proc, self.th = None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES if it is set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if self.idx < len(ids) and ids[self.idx].isdigit():
                    phys = int(ids[self.idx])
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.th = threading.Thread(target=self._read_smi, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), mhz, self.max_mhz, pw, [k for k, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]), float(r[3]),
                                  [nm for nm, v in zip(names, r[4:8]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.02)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if self.t0 is None or (self.t0 <= r[0] <= (self.t1 or 1e30))]
        if not rows:                                   # region shorter than one poll: take the nearest sample
            rows = self.rows[-1:]
        sm = [r[1] for r in rows]
        reasons = sorted({x for r in rows for x in r[4]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "power_w_max": max(r[3] for r in rows) if rows else None, "samples": len(rows), "reasons": reasons,
                "source": self.mode}


def make_graphs(workload, seed0=0):
    from ecord_b200.graphs import synthetic_set
    kind, n, G, kw, T, desc = WORKLOADS[workload]
    return synthetic_set(kind, n, G, seed0=seed0, **kw), n, G, T, desc


# ------------------------------------------------------------------------------------------------
def cpu_rollout_rate(graphs_sample, T, steps, use_ref_env=True):
    """Time the CPU restatement of the reference rollout (oracle) on a bounded sample; returns units/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ecord_oracle as orc
    from ecord_b200.network import random_state_dicts
    torch.set_num_threads(os.cpu_count())
    have_ref = False
    if use_ref_env:
        try:
            import build_ref
            have_ref = build_ref.load_step_cy() is not None
        except Exception:
            have_ref = False
    ob = orc.build_batch([tuple(g) for g in graphs_sample])
    model = orc.ModelOracle(random_state_dicts(0))
    init = np.random.RandomState(0).randint(0, 2, (ob.N, T)).astype(np.int8)
    model.initialise_embeddings(ob, T)
    env = orc.EnvOracle(ob, T, init, use_ref=have_ref)
    nf, gf = env.observe()

    def one_step():
        nonlocal nf, gf
        model.update_internal_state(gf)
        q, emb, _ = model.q_values(ob, nf)
        a = model.select(ob, q, emb)
        env.step(a.numpy())
        nf, gf = env.observe()

    one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return ob.G * T * steps / dt, dt, have_ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    graphs, n, G, T, desc = make_graphs(args.workload)
    gs = max(1, min(G, args.cpu_graphs))
    # warm-up on a small sample; then the WHOLE workload (all G graphs in one rollout batch, the b200 arm's config) if K timed
    # steps of it end within ~2 minutes at the measured pace, else the largest sample that does
    rate_w, dt_w, _ = cpu_rollout_rate(graphs[:gs], T, max(args.warmup, 1))
    est_full = dt_w / max(args.warmup, 1) * args.steps * (G / gs)
    gs = G if est_full <= 120.0 else max(1, int(G * 120.0 / est_full))
    sample = graphs[:gs]
    rate, dt, have_ref = cpu_rollout_rate(sample, T, args.steps)
    cores = os.cpu_count()
    sample_s = f"{gs} of {G} graphs x {T} tradj x {args.steps} steps ({'compiled reference Cython step' if have_ref else 'C restatement of the env step'} + torch CPU model, {torch.get_num_threads()} threads)"
    out = {"impl": "reference", "metric": "rollout steps/sec (graphs x tradj x steps)", "value": rate, "unit": "steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "note": "CPU path; per-step batch is the sample below"},
           "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample_s},
           "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from ecord_b200.data import GraphBatcher
    from ecord_b200.dist import reduce_best, shard_bounds
    from ecord_b200.engine import DeviceGraph, DeviceWeights, RolloutState
    from ecord_b200.network import random_state_dicts
    from ecord_b200.solver import B200Solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    graphs, n, G, T, desc = make_graphs(args.workload)
    K, W = args.steps, max(args.warmup, 3)
    if args.tradj:
        T = args.tradj
    strong = args.scaling == "strong"
    # weak: every rank runs T more trajectories of the same graphs; strong: the config's T trajectories are split over the
    # ranks (BASELINE config 4: "128 trajectories, sharded across 8xB200" -> 16 per GPU)
    T_total = T if strong else T * world
    lo, hi = shard_bounds(T_total, world, rank)
    T = hi - lo
    sd = random_state_dicts(0)
    batch = GraphBatcher()(graphs)
    avg_deg = batch.tg.num_edges / batch.tg.num_nodes
    # global initial spins, sliced per rank: results do not depend on the GPU count
    init_all = np.random.RandomState(1234).randint(0, 2, (batch.tg.num_nodes, T_total), dtype=np.int8)
    init = np.ascontiguousarray(init_all[:, lo:hi])

    weights = DeviceWeights(sd, dev)
    dg = DeviceGraph(batch.tg, dev)
    k = args.steps_per_graph
    if K % k and K <= 2 * k:
        # short runs (the driver's --steps 20): ONE CUDA graph of K steps, which is also what B200Solver.rollout builds for a
        # rollout of that length, instead of a full graph plus a short remainder whose cold start (L2 flushed before every
        # graph launch) would be amortised over a handful of steps only
        k = K
    W_req = W
    W = max(W, k)        # at least one untimed launch of the k-step graph: its first launch uploads the instantiated graph
    st = RolloutState(dg, weights, T, gemm=args.gemm, compat=True, max_steps=W + K + k, steps_per_graph=k)
    st.gnn_embed()
    st.env_init(init)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also instantiates the CUDA graphs, NCCL and the clock sampler) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    st.run(W)
    reduce_best(st.best_from_log(W), T_total)          # first collective call sets NCCL up: keep that out of the timing
    barrier()
    launches0 = st.kernel_launches
    # ---- timed region: exactly K steps; L2 flushed before every graph launch (outside the events) ----
    chunks = [k] * (K // k) + ([K % k] if K % k else [])
    evs = []
    barrier()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for c in chunks:
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.run(c); e1.record()
        evs.append((e0, e1))
    # final reduce (the path's only collective) is part of the job
    by_tradj = st.best_from_log(W + K)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    best, by_all = reduce_best(by_tradj, T_total)
    e3.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    ms = sum(a.elapsed_time(b) for a, b in evs) + e2.elapsed_time(e3)
    clocks = sampler.stop() if rank == 0 else None
    launches = st.kernel_launches - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    units = G * T_total * K
    value = units / (ms * 1e-3)

    # ---- roofline: every kernel of the step timed ALONE with CUDA events on the rollout stream (L2 flushed before every
    #      launch), the dominant one reported as `roofline`, the rest under `others`.  Algorithmic work:
    #        GRU   2 x M x (64 + 1024) x 3072 FLOP       PROJ  2 x M x 1024 x 512 FLOP      tail  2 x M x (512 x 32 + 32 x 32 + 32) FLOP
    #        Q     8416 FLOP and 12 B per (vertex, trajectory) row                          env   (20 d + 76) B per unit
    #      literal FLOPs of the reference's fp32 model: the bf16x3 mode executes three tensor-core products for each of them. ----
    roof = None
    if rank == 0:
        import ctypes as C
        reps = 20
        hbm, tf_burst, tf_sus, which = peaks()
        M = G * T
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic_tab = json.load(f).get(args.workload, {})
        except Exception:
            traffic_tab = {}

        def timed(fn, n=reps):
            """Mean duration of one launch, L2 flushed before each.  The events and the kernel are enqueued while a ~150 us
            spin kernel still runs ahead of them on the rollout stream, so that the host's launch path (ctypes call, tensor-map
            encoding) is not inside the event pair; the median-of-n guards against a stray long launch."""
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.fill_(1)
                st.stream.wait_stream(torch.cuda.current_stream())     # the flush has finished before the kernel starts
                with torch.cuda.stream(st.stream):
                    torch.cuda._sleep(300000)
                    a.record(); fn(); b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            mid = ts[len(ts) // 4: len(ts) - len(ts) // 4]          # interquartile mean
            return sum(mid) / len(mid)

        sp = C.c_void_p(st.stream.cuda_stream)

        def policy_part(mask):
            def run():
                os.environ["ECORD_POLICY_MASK"] = str(mask)
                st.lib.ecord_policy_step(C.byref(st.g.c), C.byref(st.env_c), C.byref(st.w.c), C.byref(st.policy_c), int(st.steps_done & 1),
                                         int(st.gemm), int(st.fused), sp)
                os.environ.pop("ECORD_POLICY_MASK", None)
            return run

        def q_part():
            st.lib.ecord_q_select(C.byref(st.g.c), C.byref(st.env_c), C.byref(st.w.c), C.byref(st.policy_c), int(st.steps_done),
                                  C.c_float(st.tau), C.c_uint64(st.seed), int(st.compat), None, int(st.q_mode), C.c_float(st.epsilon), sp)

        def env_part():
            st.lib.ecord_env_step_fused(C.byref(st.g.c), C.byref(st.env_c), C.byref(st.w.c), C.byref(st.policy_c), int(st.steps_done), None, sp)

        tms = {}
        if st.gemm != 0:
            tms["gru"], tms["proj"], tms["tail"] = timed(policy_part(1)), timed(policy_part(2)), timed(policy_part(4))
        else:
            tms["policy_fp32"] = timed(policy_part(7))
        tms["q"] = timed(q_part)
        tms["env"] = timed(env_part)
        st.gnn_embed(); torch.cuda.synchronize()       # (its workspace allocation stays outside the timing)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            st.gnn_embed()
        b.record(); b.synchronize()
        tms["gnn"] = a.elapsed_time(b) / 5
        rows = float(dg.N) * T
        gm = args.gemm
        ns = "3" if gm == "bf16x3" else "1"

        def entry(kernel, bound, ms, work, unit_work, traffic_key=None, note=None, extra=None):
            peak = tf_burst if bound == "tensor" else hbm
            ach = work / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
            e = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                 "frac": ach / peak, "traffic": traffic_tab.get(traffic_key) if traffic_key else None, "ms_per_launch": ms,
                 "launch_share_of_step": ms / (ms_total / K), "algorithmic": unit_work,
                 "peak_source": which + (" (bf16 burst; kernel timed alone)" if bound == "tensor" else " (copy bandwidth)")}
            if note:
                e["note"] = note
            if extra:
                e.update(extra)
            return e

        ms_total = ms
        ents = []
        if "gru" in tms:
            ents.append(entry(f"k_gemm_tc2<0,{ns}> (GRU cell: tcgen05 cta_group::2, {gm})", "tensor", tms["gru"], 2.0 * M * 1088 * 3072,
                              {"flops_per_unit": 2 * 1088 * 3072, "units_per_launch": M, "tensor_products_per_literal_flop": int(ns)},
                              f"k_gemm_tc2<0, {ns}>",
                              "literal fp32-model FLOPs; bf16x3 issues 3 MMAs per literal product (hi.hi + lo.hi + hi.lo), so the tensor pipe "
                              "does 3x this work" if ns == "3" else None,
                              extra={"executed_frac_of_peak": int(ns) * 2.0 * M * 1088 * 3072 / (tms["gru"] * 1e-3) / 1e12 / tf_burst}))
            ents.append(entry(f"k_gemm_tc2<1|2,{ns}> (projection 1024->512)", "tensor", tms["proj"], 2.0 * M * 1024 * 512,
                              {"flops_per_unit": 2 * 1024 * 512, "units_per_launch": M}, f"k_gemm_tc2<2, {ns}>"))
            ents.append(entry("k_tail_tc (LN512, 512->32 on tcgen05, LN32, value head, Q operands)", "tensor", tms["tail"],
                              2.0 * M * (512 * 32 + 32 * 32 + 32 + 32 * 64), {"flops_per_unit": 2 * (512 * 32 + 32 * 32 + 32 + 32 * 64), "units_per_launch": M},
                              "k_tail_tc", "latency chain per CTA, not a throughput kernel"))
        else:
            ents.append(entry("k_sgemm_nt_bias x3 + k_gru_gates + k_policy_tail (fp32 CUDA-core policy step)", "tensor", tms["policy_fp32"],
                              float(F_UNIT(0)) * M, {"flops_per_unit": F_UNIT(0), "units_per_launch": M}))
        qname = {2: "k_q_tc (tcgen05 kind::f16, fp16 hi/lo) + k_select_finish", 1: "k_q_fast + k_select_finish", 0: "k_q_select"}[st.q_mode]
        ents.append(entry(qname, "tensor", tms["q"], 8416.0 * rows, {"flops_per_row": 8416, "bytes_per_row": 12, "rows_per_launch": rows}, "k_q_tc",
                          extra={"hbm_view": {"achieved_gbs": 12.0 * rows / (tms["q"] * 1e-3) / 1e9, "peak_gbs": hbm,
                                              "frac": 12.0 * rows / (tms["q"] * 1e-3) / 1e9 / hbm}}))
        ents.append(entry("k_env_step_fused (flip + neighbour deltas + next msg_in)", "hbm", tms["env"], (20.0 * avg_deg + 76.0) * M,
                          {"bytes_per_unit": 20.0 * avg_deg + 76.0, "units_per_launch": M}, "k_env_step_fused",
                          "latency-bound: four dependent loads (action -> CSR row -> neighbour entries -> write); the DRAM traffic ncu "
                          "counts is sector-granular (32 B per scattered 4-byte access)"))
        # t=0 GNN: compulsory HBM bytes per layer = E x 8 (incoming src + weight) + N x 192 (x, y, state in / out); the E x 64 B of
        # gathers hit the L2-resident y[N,16] and are not HBM traffic
        gnn_bytes = 4.0 * (dg.E * 8 + dg.N * 192)
        ents.append(entry("gnn.cu (t=0 embedding: 4 message-passing layers, once per ROLLOUT, not per step)", "hbm", tms["gnn"], gnn_bytes,
                          {"bytes_per_layer": dg.E * 8 + dg.N * 192, "layers": 4}, "k_gnn_update",
                          "sequential per-vertex accumulation reproduces the reference's summation order; gathers are L2 hits"))
        step_kernels = [e for e in ents if not e["kernel"].startswith("gnn.cu")]
        dom = max(step_kernels, key=lambda e: e["ms_per_launch"])
        roof = dict(dom)
        roof["others"] = [e for e in ents if e is not dom]
        roof["kernel_ms_sum"] = sum(e["ms_per_launch"] for e in step_kernels)
        for e in step_kernels + [roof]:       # the share the ncu launch list of the same command shows (profiles/*_launches_*.txt)
            if not e["kernel"].startswith("gnn.cu"):
                e["share_of_kernel_sum"] = e["ms_per_launch"] / roof["kernel_ms_sum"]
        roof["whole_step"] = {"flops_per_unit": F_UNIT(n), "bytes_per_unit": B_UNIT(n, avg_deg),
                              "tensor_bound_units_s": tf_sus * 1e12 / F_UNIT(n), "hbm_bound_units_s": hbm * 1e9 / B_UNIT(n, avg_deg),
                              "frac_of_binding_bound": (value / world) / min(tf_sus * 1e12 / F_UNIT(n), hbm * 1e9 / B_UNIT(n, avg_deg)),
                              "note": "literal fp32-model FLOPs over the sustained bf16 peak"}
        if gm == "bf16x3":
            # what the tensor pipe executes in this mode: three bf16 products for every literal product of the GRU / projection /
            # tail GEMMs (the price of fp32-grade recurrent state on bf16 tensor cores), the Q head as counted
            f_exec = 3.0 * F_UNIT(0) + 8416.0 * n
            roof["whole_step"]["executed"] = {"flops_per_unit": f_exec, "frac_of_tensor_peak": (value / world) * f_exec / (tf_sus * 1e12),
                                              "note": "bf16x3 executes 3 tensor-core products per literal GEMM product"}
    st.close()

    # ---- e2e: the public API with host buffers (H2D of inputs, D2H of the score log inside the timed region) ----
    solver = B200Solver(gnn=sd['gnn'], rnn=sd['rnn'], gnn2rnn=sd['gnn2rnn'], node_encoder=sd['node_encoder'],
                        device=dev, gemm=args.gemm, steps_per_graph=args.steps_per_graph)
    Ke = K
    # warm-up call of the same shape: the solver keeps the device state of a graph batch (buffers, pinned staging, captured
    # CUDA graphs) and re-initialises it on the next call, which is the steady state of test_solver's trajectory chunks
    pinned = torch.from_numpy(init).pin_memory()
    log = None
    for _ in range(3):                  # (three of them: the page-locked result buffers of a live log plus the next call's)
        log = solver.rollout(batch, num_tradj=T, num_steps=Ke, tau=0, use_epsilon=False, init_state=pinned.numpy())
    e2e_times = []
    for _ in range(3):                  # median of three calls
        barrier()
        t0 = time.perf_counter()
        log = solver.rollout(batch, num_tradj=T, num_steps=Ke, tau=0, use_epsilon=False, init_state=pinned.numpy())
        sc = log['best_score_by_tradj'].to(dev)          # == torch.stack(log['scores'], -1).max(-1).values, reduced on the device
        best_e, _ = reduce_best(sc, T_total)
        best_host = best_e.cpu()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_times.append(float(te.item()))
        e2e_dev = {k: float(torch.as_tensor(log[k]).flatten()[0]) for k in ("t_init_emb", "t_setup", "t_rollout")}
    te = torch.tensor([sorted(e2e_times)[1]], device=dev)
    e2e_val = G * T_total * Ke / float(te.item())
    h2d = (init.nbytes) / Ke
    d2h = (G * T * 4 * Ke + dg.N * T * 4 + G * T * 8 + best_host.numel() * 4) / Ke

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            gs = max(1, min(G, args.cpu_graphs))
            rate, dt, have_ref = cpu_rollout_rate(graphs[:gs], T, args.cpu_steps)
            cb = {"value": rate, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                  "sample": f"{gs} of {G} graphs x {T} tradj x {args.cpu_steps} steps in {dt:.1f}s "
                            f"({'compiled reference Cython step (oracle/_ref)' if have_ref else 'C restatement of the env step'}"
                            f" + torch CPU model, {torch.get_num_threads()} threads)"}
        out = {"metric": "rollout steps/sec (graphs x tradj x steps)", "value": value, "unit": "steps/s", "n_gpus": world,
               "steps": K, "warmup": W_req, "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling,
               "vs_baseline": None,
               "dtype": {"bf16x3": "bf16x3 (tcgen05 on bf16 hi/lo operand pairs, fp32 accumulate: fp32-grade)", "bf16": "bf16",
                         "fp32": "f32"}[args.gemm], "data": "synthetic",
               "config": {"workload": desc, "graphs": G, "tradj_total": T_total, "units_per_step": G * T_total,
                          "parallelism": f"trajectory-sharded x{world}", "gemm": args.gemm,
                          "l2": "flushed (256 MB write) before every CUDA-graph launch of %d steps, outside the timed events" % k,
                          "warmup_steps_run": W,
                          "weights": "random-init canonical ECORD network (seed 0)"},
               "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "note": "B200Solver.rollout incl. GNN, env init, H2D init states, D2H score log and best states; median of 3 calls after three warm-up calls of the same shape (device state of the batch and the page-locked result buffers reused)",
                       "calls_s": e2e_times, "device_s_last_call": e2e_dev},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cb,
               "wall_s_timed_region": t_wall, "best_cut_mean": float(best.mean().item())}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps; default = the config's rollout length 4*|V|, capped at 2000")
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="er500", choices=sorted(WORKLOADS))
    ap.add_argument("--gemm", default="bf16x3", choices=["bf16x3", "bf16", "fp32"],
                    help="bf16x3 (default): tcgen05 GEMMs on bf16 hi/lo operand pairs, fp32-grade, passes the 99 %% best-cut "
                         "tests; bf16: single bf16 products (fast mode, 2e-2 tolerance); fp32: CUDA-core SGEMM")
    ap.add_argument("--steps-per-graph", type=int, default=16)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = T more trajectories per rank (default); strong = the workload's T trajectories split over the ranks")
    ap.add_argument("--tradj", type=int, default=0, help="override the workload's trajectories per graph (per GPU in weak scaling)")
    ap.add_argument("--cpu-graphs", type=int, default=10, help="graphs in the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=24, help="steps of the CPU baseline sample (b200 arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = min(4 * WORKLOADS[args.workload][1], 2000) if args.impl == "b200" else 8
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
