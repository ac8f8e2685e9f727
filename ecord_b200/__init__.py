"""ecord_b200 -- B200-native rollout engine for ECORD (Max-Cut RL solver), behind the reference's
solver / validate API.  See DESIGN.md for the hot path and its boundary, INTEGRATION.md for the
reference-side switch."""
from .data import Batch, EdgeListGraph, GraphBatcher            # noqa: F401
from .observations import GlobalObservation, NodeObservation    # noqa: F401
from .network import random_state_dicts                          # noqa: F401


def __getattr__(name):
    # solver / testing import the CUDA library lazily so that `import ecord_b200` works on a build box
    if name in ("B200Solver", "DQNSolver"):
        from . import solver
        return getattr(solver, name)
    if name in ("TestGraphsConfig", "test_solver", "merge_log"):
        from . import testing
        return getattr(testing, name)
    raise AttributeError(name)
