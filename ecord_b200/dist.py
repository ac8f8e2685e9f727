"""Multi-GPU sharding of a rollout: one process per GPU, no data-path collective.

Every (graph, trajectory) unit is independent for all S steps (GRU state per (g,t), env state per
(vertex,t), weights and CSR read-only), so the trajectories of a rollout batch are split across ranks:
rank r runs trajectories [lo_r, hi_r) of the SAME graphs.  All ranks hold the whole batch's CSR and run the
tiny t=0 GNN redundantly, which keeps its batch-wide LayerNorm statistics identical to the single-GPU run
(quirk B).  Initial spins are drawn once for the global [N, T] array and sliced, so results do not depend
on the number of GPUs.  (Graph sharding -- `rollout_graph_sharded`: every rank rolls out all trajectories of a contiguous
range of graphs, for sets of many small graphs -- is the other axis; see graph_shard() for what quirk A asks of it.)
The only communication is at the END of the rollout:
    one all_gather of the per-trajectory best  f32[G, T_local]   (needed for scores_mean, testing.py:201-204);
    the per-graph best cut f32[G] is the maximum of the gathered values
over NCCL (NVLink/NVSwitch) on GPUs, or gloo in the CPU tests.  Payloads are KBs.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous, balanced split of range(total): the first (total % world) ranks get one extra."""
    base, extra = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_best(best_by_tradj_local, total_tradj, group=None):
    """best_by_tradj_local: f32 [G, T_local] on this rank (any device).
    Returns (best_cut f32[G] over all trajectories of all ranks, best_by_tradj f32[G, total_tradj])."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return best_by_tradj_local.max(-1).values, best_by_tradj_local
    world = dist.get_world_size(group)
    G = best_by_tradj_local.shape[0]
    if int(total_tradj) % world == 0:
        # equal shards: gather straight from the caller's tensor, one transposing view, one max -- three device ops around the
        # collective instead of a dozen (the whole reduce sits inside bench.py's timed region)
        t = int(total_tradj) // world
        gathered = torch.empty((world * G, t), dtype=best_by_tradj_local.dtype, device=best_by_tradj_local.device)
        dist.all_gather_into_tensor(gathered, best_by_tradj_local.contiguous(), group=group)      # shards concatenated along dim 0
        by_all = gathered.view(world, G, t).permute(1, 0, 2).reshape(G, world * t)
        return by_all.max(-1).values, by_all
    # ONE collective: ragged split -> pad every shard to the largest, all_gather into one tensor, cut the padding out;
    # the best cut is the maximum of the gathered values (an all_reduce(MAX) on top would be a second ~50 us launch)
    t_max = -(-int(total_tradj) // world)
    pad = torch.full((G, t_max), float("-inf"), dtype=best_by_tradj_local.dtype, device=best_by_tradj_local.device)
    pad[:, :best_by_tradj_local.shape[1]] = best_by_tradj_local
    gathered = torch.empty((world * G, t_max), dtype=pad.dtype, device=pad.device)      # shards concatenated along dim 0
    dist.all_gather_into_tensor(gathered, pad.contiguous(), group=group)
    gathered = gathered.view(world, G, t_max)
    cols = []
    for r in range(world):
        lo, hi = shard_bounds(total_tradj, world, r)
        cols.append(gathered[r, :, :hi - lo])
    by_all = torch.cat(cols, dim=1)
    return by_all.max(-1).values, by_all


def graph_shard(num_nodes_list, world_size, rank, compat=True):
    """Graph sharding: (subset, own) -- `own` = the contiguous, balanced range of graphs whose results this rank reports,
    `subset` = the graphs it has to roll out for that, in order.

    With quirk A on (compat, the reference's behaviour: actors/ecord.py:183-193 gathers the cached embedding of the chosen
    vertex with its LOCAL index used as a batch-wide one) a unit of graph g reads the state of vertex a < n_g of the BATCH,
    i.e. of the leading graphs that cover vertex indices [0, n_g).  Those graphs evolve on their own (their quirk-A reads
    stay inside the same prefix), so a rank reproduces the single-GPU rollout of its graphs exactly by rolling the prefix
    out as well, in front of its own graphs, and dropping the prefix's results.  compat=False needs no prefix."""
    n = [int(x) for x in num_nodes_list]
    lo, hi = shard_bounds(len(n), world_size, rank)
    own = list(range(lo, hi))
    if not own:
        return [], []
    prefix = []
    if compat:
        need, covered, g = max(n[lo:hi]), 0, 0
        while covered < need:
            prefix.append(g)
            covered += n[g]
            g += 1
    subset = prefix + [g for g in own if g not in set(prefix)]
    return subset, own


def gather_graph_shards(local, num_graphs, group=None):
    """local: f32 [G_own, ...] rows of this rank's own graphs (contiguous, balanced split).  Returns [num_graphs, ...] on
    every rank: ONE all_gather of shards padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    g_max = -(-int(num_graphs) // world)
    pad = torch.zeros((g_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * g_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.view((world, g_max) + tuple(local.shape[1:]))
    rows = []
    for r in range(world):
        lo, hi = shard_bounds(num_graphs, world, r)
        rows.append(out[r, :hi - lo])
    return torch.cat(rows, dim=0)


def rollout_graph_sharded(solver, batch, num_tradj, num_steps, init_state, tau=0.0, group=None, world_size=None, rank=None):
    """Run this rank's GRAPH shard of a rollout (all trajectories of a contiguous range of graphs; for many-small-graph sets,
    where a trajectory shard would leave too few rows per GPU) and gather the per-graph results.  `init_state` is the GLOBAL
    int8 [N, num_tradj] array.  world_size / rank override the process group's (single-process checks)."""
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    tg = batch.tg
    n_list = tg.num_nodes_list if hasattr(tg, "num_nodes_list") else (tg.ptr[1:] - tg.ptr[:-1]).tolist()
    G = len(n_list)
    subset, own = graph_shard(n_list, world_size, rank, compat=solver.compat)
    if not own:
        by_tradj = torch.zeros(0, num_tradj, device=solver.device)
        log = {}
    else:
        log = solver.rollout(batch, num_tradj=num_tradj, num_steps=num_steps, tau=tau, use_epsilon=False,
                             init_state=init_state, graph_subset=subset)
        keep = torch.as_tensor([subset.index(g) for g in own])
        by_tradj = log['best_score_by_tradj'][keep].to(solver.device)
        log['graph_subset'], log['graphs_own'] = [subset], [own]
    by_all = gather_graph_shards(by_tradj, G, group)
    log['scores_max'] = by_all.max(-1).values.cpu()
    log['scores_mean'] = by_all.mean(-1).cpu()
    log['scores_by_tradj'] = by_all.cpu()
    return log


def rollout_sharded(solver, batch, num_tradj, num_steps, init_state, tau=0.0, group=None):
    """Run this rank's trajectory shard of a rollout and reduce the results.  `init_state` is the GLOBAL
    int8 [N, num_tradj] array (identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(num_tradj, world, rank)
    log = solver.rollout(batch, num_tradj=num_tradj, num_steps=num_steps, tau=tau, use_epsilon=False,
                         init_state=init_state, traj_slice=(lo, hi))
    if 'best_score_by_tradj' in log:                        # reduced on the device by the engine
        by_tradj = log['best_score_by_tradj'].to(solver.device)
    else:
        by_tradj = torch.stack(log['scores'], -1).max(-1).values.to(solver.device)   # [G, T_local, S] -> [G, T_local]
    best, by_all = reduce_best(by_tradj, num_tradj, group)
    log['scores_max'] = best.cpu()
    log['scores_mean'] = by_all.mean(-1).cpu()
    log['scores_by_tradj'] = by_all.cpu()
    return log
