"""mapf_rl_b200 — B200-native batched MAPF environment hot path (step + observe, BFS heuristic maps,
PER sum-tree + TD) behind the reference's `Environment` / `SumTree` / `LocalBuffer` interface.

Host side is Python/PyTorch (device memory, streams); all compute is hand-written sm_100a CUDA in
libmapf_b200.so, called through the C ABI in include/mapf_b200.h.  There is no CPU fallback.
"""
from . import config  # noqa: F401

__all__ = ["config", "BatchedEnvironment", "Environment", "SumTree", "LocalBuffer", "PrioritizedReplayTree", "ReplayStore"]


def __getattr__(name):
    if name == "BatchedEnvironment":
        from .batched import BatchedEnvironment
        return BatchedEnvironment
    if name == "Environment":
        from .environment import Environment
        return Environment
    if name in ("SumTree", "LocalBuffer", "PrioritizedReplayTree"):
        from . import buffer
        return getattr(buffer, name)
    if name == "ReplayStore":
        from .replay import ReplayStore
        return ReplayStore
    raise AttributeError(name)
