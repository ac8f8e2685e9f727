"""Multi-GPU sharding of a rollout: one process per GPU, no data-path collective.

Every (graph, trajectory) unit is independent for all S steps (GRU state per (g,t), env state per
(vertex,t), weights and CSR read-only), so the trajectories of a rollout batch are split across ranks:
rank r runs trajectories [lo_r, hi_r) of the SAME graphs.  All ranks hold the whole batch's CSR and run the
tiny t=0 GNN redundantly, which keeps its batch-wide LayerNorm statistics identical to the single-GPU run
(quirk B).  Initial spins are drawn once for the global [N, T] array and sliced, so results do not depend
on the number of GPUs.  The only communication is at the END of the rollout:
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
