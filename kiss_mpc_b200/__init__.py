"""kiss_mpc_b200 -- B200-native batched replacement for the per-step NLP solve of rtarun1/kiss-mpc
(mpc/optimizer.py MotionPlanner.solve -> CasADi/IPOPT).  See DESIGN.md."""
from ._lib import KmpcError  # noqa: F401
from .mapping import map_to_circles, read_pgm  # noqa: F401
from .model import Model  # noqa: F401
from .planner import (BatchedMotionPlanner, MotionPlanner, PlannerConfig, RankGather, ShardedMotionPlanner, SolveResult,  # noqa: F401
                      STATUS_NAMES, gather_results, shard_range)

__all__ = ["BatchedMotionPlanner", "Model", "MotionPlanner", "PlannerConfig", "RankGather", "ShardedMotionPlanner", "SolveResult",
           "STATUS_NAMES", "KmpcError", "gather_results", "shard_range", "map_to_circles", "read_pgm"]
