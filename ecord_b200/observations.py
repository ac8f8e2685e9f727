"""Observation channel names -- the API constants of the reference's environment
(`ecord/environment/observations.py:4-44`); values are kept identical so pickled configs and
`node_features=[...]` / `global_features=[...]` solver arguments carry over unchanged."""
from enum import Enum


class NodeObservation(Enum):
    STATE = 0
    PEEK = 1
    PEEK_NORM = 2
    PEEK_SCALE = 3
    DEGREE_NORM = 4
    STEPS_STATIC = 5
    STEPS_STATIC_NORM = 6
    STEPS_STATIC_CLIP = 7
    STATE_BEST = 8
    LAPLACIAN_PE = 9
    ID = 10


class GlobalObservation(Enum):
    SCORE_FROM_BEST = 0
    SCORE_FROM_BEST_NORM = 1
    SCORE_FROM_BEST_SCALE = 12
    NUM_GREEDY_ACTIONS = 2
    NUM_GREEDY_ACTIONS_NORM = 3
    NUM_GREEDY_ACTIONS_CLIP = 4
    STEPS_SINCE_BEST = 5
    STEPS_SINCE_BEST_NORM = 6
    STEPS_SINCE_BEST_CLIP = 7
    STEPS = 8
    STEPS_NORM = 9
    PEEK_MAX = 10
    PEEK_MAX_NORM = 11
    PEEK_MAX_SCALE = 14
    PEEK_MEAN = 15
    PEEK_MEAN_NORM = 16


# The one observation set the fused rollout implements: the canonical ECORD configuration
# (experiments/train.py:62-70).  PEEK_MAX_NORM is never configured by the reference (quirk C,
# tradjectories.py:294) so the second global feature is identically zero.
CANONICAL_NODE_FEATURES = [NodeObservation.STATE, NodeObservation.PEEK_NORM, NodeObservation.STEPS_STATIC_CLIP]
CANONICAL_GLOB_FEATURES = [GlobalObservation.SCORE_FROM_BEST_NORM, GlobalObservation.PEEK_MAX_NORM]


def _names(feats):
    return [getattr(f, "name", str(f)) for f in feats]


def check_canonical(node_features, global_features):
    """Enum identity differs between this module and a pickled reference enum, so compare by name."""
    if _names(node_features) != _names(CANONICAL_NODE_FEATURES) or _names(global_features) != _names(CANONICAL_GLOB_FEATURES):
        raise NotImplementedError(
            "the B200 rollout engine implements the canonical ECORD observation set "
            f"{_names(CANONICAL_NODE_FEATURES)} / {_names(CANONICAL_GLOB_FEATURES)}; got "
            f"{_names(node_features)} / {_names(global_features)}")


# ---- what a training-time rollout hands to its callbacks -------------------------------------
# Field-for-field mirrors of the reference's EnvObservation (environment/environment.py:12-38) and ActorState
# (solvers/actors/_base.py:24-47), so that a replay-buffer callback written for DQNSolver.rollout reads them unchanged.
from dataclasses import dataclass
from typing import Optional

import torch


def _to(obj, names, device):
    for n in names:
        v = getattr(obj, n)
        if torch.is_tensor(v):
            setattr(obj, n, v.to(device))
    return obj


@dataclass
class EnvObservation:
    node_features: torch.Tensor              # [num_nodes, num_tradj, 3] normalised (STATE, PEEK_NORM, STEPS_STATIC_CLIP)
    glob_features: torch.Tensor              # [num_graphs, num_tradj, 2] (SCORE_FROM_BEST_NORM, 0)
    edge_index: Optional[torch.Tensor]       # None: the canonical actor does not read the graph structure per step
    edge_attr: Optional[torch.Tensor]
    batch_idx: torch.Tensor                  # [num_nodes]
    batch_ptr: torch.Tensor                  # [num_graphs + 1]
    degree: Optional[torch.Tensor]
    num_tradj: int

    def to(self, device):
        return _to(self, ("node_features", "glob_features", "edge_index", "edge_attr", "batch_idx", "batch_ptr", "degree"), device)


@dataclass
class ActorState:
    hidden_embeddings: torch.Tensor          # [num_graphs, num_tradj, 1024] GRU state BEFORE the step
    init_node_embeddings: torch.Tensor       # [num_nodes, num_tradj, 16] gnn2rnn(gnn(x)), the same for every trajectory
    last_selected_node_embeddings: Optional[torch.Tensor] = None    # [num_graphs, num_tradj, 32]
    cell_embeddings: Optional[torch.Tensor] = None
    graph_embeddings: Optional[torch.Tensor] = None

    def to(self, *args, **kwargs):
        for n in ("hidden_embeddings", "init_node_embeddings", "last_selected_node_embeddings", "cell_embeddings", "graph_embeddings"):
            v = getattr(self, n)
            if v is not None:
                setattr(self, n, v.to(*args, **kwargs))
        return self
