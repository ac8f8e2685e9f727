"""Validation driver: the caller of seam B1.

Mirrors `ecord/validate/testing.py`: `TestGraphsConfig` (:16-117), `merge_log` (:120-143) and
`test_solver` (:146-230) with the same chunking (graphs_bsz x tradj_bsz), log merging and result
semantics: scores_max[g] = max over trajectories and steps >= 1 of the logged score, scores_mean[g] =
mean over trajectories of the per-trajectory maximum, apx_* = scores / best-known cut.
"""
import math
import os
import time
from collections import defaultdict
from dataclasses import dataclass
from typing import List, Optional

import torch

from .graphs import load_graph_from_file


@dataclass()
class TestGraphsConfig:
    __test__ = False   # not a pytest class
    graphs: List
    opt_scores: Optional[List]
    num_steps: int
    tradj_per_graph: int
    act_greedy: bool = False
    use_network_initialisation: bool = False
    tau: float = None
    max_time: float = None
    pre_solve_with_greedy: bool = False
    post_solve_with_greedy_from_best: bool = False
    graphs_bsz: int = None
    tradj_bsz: int = None
    fname: str = None
    label: Optional[str] = None

    def __post_init__(self):
        self.num_graphs = len(self.graphs)
        if self.graphs_bsz is None or self.graphs_bsz >= self.num_graphs:
            self._num_graph_batches = 1
            self.graphs_bsz = self.num_graphs
        else:
            self._num_graph_batches = math.ceil(self.num_graphs / self.graphs_bsz)
        if self.tradj_bsz is None or self.tradj_bsz >= self.tradj_per_graph:
            self._num_tradj_batches = 1
            self.tradj_bsz = self.tradj_per_graph
        else:
            self._num_tradj_batches = math.ceil(self.tradj_per_graph / self.tradj_bsz)

    @staticmethod
    def from_file(graph_loc, num_load=None, idx_load=None, num_steps=-1, tradj_per_graph=1, act_greedy=False,
                  use_network_initialisation=False, tau=None, max_time=None, post_solve_with_greedy_from_best=False,
                  pre_solve_with_greedy=False, graphs_bsz=None, tradj_bsz=None, label=None):
        graphs, opt_scores = load_graph_from_file(graph_loc, quiet=True)
        if num_load is not None:
            graphs = graphs[:num_load]
            opt_scores = opt_scores[:num_load] if opt_scores else opt_scores
        elif idx_load is not None:
            if isinstance(idx_load, (list, tuple)) and len(idx_load) == 2:
                sl = slice(idx_load[0], idx_load[1])
                graphs = graphs[sl]
                opt_scores = opt_scores[sl] if opt_scores else opt_scores
            else:
                graphs = [graphs[i] for i in idx_load]
                opt_scores = [opt_scores[i] for i in idx_load] if opt_scores else opt_scores
        if opt_scores:
            opt_scores = torch.tensor(opt_scores)
        return TestGraphsConfig(graphs=graphs, opt_scores=opt_scores, num_steps=num_steps,
                                tradj_per_graph=tradj_per_graph, act_greedy=act_greedy,
                                use_network_initialisation=use_network_initialisation, tau=tau, max_time=max_time,
                                pre_solve_with_greedy=pre_solve_with_greedy,
                                post_solve_with_greedy_from_best=post_solve_with_greedy_from_best,
                                graphs_bsz=graphs_bsz, tradj_bsz=tradj_bsz, fname=graph_loc, label=label)

    def get_num_batches(self):
        return self._num_graph_batches, self._num_tradj_batches

    def get_summary(self):
        s = f"Testing {self.num_graphs} graphs "
        if self.fname is not None:
            s += f"from {self.fname}, "
        return s + f"with {self.graphs_bsz} graphs and {self.tradj_bsz} tradjectories per rollout."

    def get_label(self):
        if self.label is not None:
            return self.label
        if self.fname is not None:
            return os.path.split(self.fname)[-1]
        return "test_graphs"


def merge_log(log_src, log_trg, dim=-1):
    """testing.py:120-143: tensors are stacked on a new last dim (lists) and concatenated on `dim`;
    lists of floats are summed element-wise."""
    for k, v in log_src.items():
        v_trg = log_trg.get(k, None)
        if isinstance(v, (int, float)):
            log_trg[k] = v + (v_trg or 0)
            continue
        if len(v) == 0:
            continue
        if isinstance(v[0], torch.Tensor):
            if not isinstance(v, torch.Tensor):
                v = torch.stack(v, dim=-1)
            if v_trg is None:
                log_trg[k] = v
            elif v_trg.dim() > dim and v.dim() > dim:
                log_trg[k] = torch.cat([v_trg, v], dim=dim)
            else:
                log_trg[k] = torch.cat([v_trg, v], dim=0)
        else:
            v = torch.tensor(v)
            log_trg[k] = v if v_trg is None else v_trg + v
    return log_trg


@torch.no_grad()
def test_solver(solver, graph_batcher, config, log_stats=False, verbose=False, init_state_fn=None):
    """testing.py:146-230.  init_state_fn(graph_chunk, tradj_chunk, N, T) -> int8 [N,T] optionally pins the
    initial spins (the reference draws them from numpy's global RNG inside the environment)."""
    num_g_batches, num_t_batches = config.get_num_batches()
    log = defaultdict()
    was_training = getattr(solver, "is_training", False)
    solver.test()
    tot_solver_time = 0
    for i in range(num_g_batches):
        batch = graph_batcher(config.graphs[i * config.graphs_bsz: min((i + 1) * config.graphs_bsz, config.num_graphs)])
        log_batch = defaultdict()
        tot_tradj = 0
        for j in range(num_t_batches):
            num_tradj = min(config.tradj_bsz, config.tradj_per_graph - tot_tradj)
            if verbose:
                print(f"\t...rollout {i * num_t_batches + j + 1}/{num_g_batches * num_t_batches}", end="...")
            kw = {}
            if init_state_fn is not None:
                kw['init_state'] = init_state_fn(i, j, batch.tg.num_nodes, config.tradj_bsz)
            t = time.time()
            _log = solver.rollout(batch, num_tradj=config.tradj_bsz, num_steps=config.num_steps, log_stats=log_stats,
                                  use_epsilon=False, use_network_initialisation=config.use_network_initialisation,
                                  tau=config.tau, max_time=config.max_time,
                                  pre_solve_with_greedy=config.pre_solve_with_greedy,
                                  post_solve_with_greedy_from_best=config.post_solve_with_greedy_from_best, **kw)
            tot_solver_time += (time.time() - t) - torch.cat(_log['t_setup']).mean()
            if verbose:
                print("done", end="\r")
            log_batch = merge_log(_log, log_batch, dim=1)
            tot_tradj += num_tradj
            solver.actor.reset_state()
        log = merge_log(log_batch, log, dim=0)

    scores_by_tradj, score_idx_by_tradj = log['scores'].max(-1)
    log['scores_max'] = scores_by_tradj.max(-1).values
    log['scores_mean'] = scores_by_tradj.mean(-1)
    log['scores_idx_mean'] = score_idx_by_tradj.float().mean(-1)
    log['init_score_mean'] = log['init_score'].mean(-1)
    log['tot_solver_time'] = tot_solver_time
    log['num_graphs'] = config.num_graphs
    log['num_tradj'] = config.tradj_per_graph
    if config.opt_scores is not None and len(config.opt_scores) > 0:
        opt = torch.as_tensor(config.opt_scores)
        log['opt'] = opt
        log['apx_max'] = log['scores_max'] / opt
        log['apx_mean'] = log['scores_mean'] / opt
        log['init_apx_score_mean'] = log['init_score_mean'] / opt
    else:
        log['apx_max'] = log['apx_mean'] = log['init_apx_score_mean'] = None
    if was_training:
        solver.train()
    return log
