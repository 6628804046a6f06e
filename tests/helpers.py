"""Shared test helpers: golden fixture access and action-stream generators."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


_inst_cache = {}


def instances(num_agents):
    """-> maps uint8[200,40,40], agents uint8[200,N,2], goals uint8[200,N,2] (copy of test{N}_40_0.3.pkl)."""
    if "z" not in _inst_cache:
        z = golden("instances_40_0.3.npz")
        maps = np.unpackbits(z["maps_packed"], axis=1)[:, :1600].reshape(200, 40, 40).astype(np.uint8)
        _inst_cache["z"] = (z, maps)
    z, maps = _inst_cache["z"]
    return maps, z[f"agents{num_agents}"], z[f"goals{num_agents}"]


def greedy_actions(obs, rng, eps=0.1):
    """navi-greedy stream G for a batch: obs bool/uint8 [B,N,6,9,9] -> uint8 [B,N]."""
    B, N = obs.shape[:2]
    bits = obs[:, :, 2:6, 4, 4].astype(bool)                       # [B,N,4]
    score = rng.random((B, N, 4)) * bits                           # random tie-break among set bits
    act = np.where(bits.any(-1), 1 + score.argmax(-1), 0)
    explore = rng.random((B, N)) < eps
    act = np.where(explore, rng.integers(0, 5, size=(B, N)), act)
    return act.astype(np.uint8)


def random_instance(rng, L, N, density):
    """Unconstrained random instance (starts and goals distinct free cells, any component)."""
    while True:
        m = (rng.random((L, L)) < density).astype(np.uint8)
        free = np.argwhere(m == 0)
        if len(free) >= N:
            break
    a = free[rng.permutation(len(free))[:N]]
    g = free[rng.permutation(len(free))[:N]]
    return m, a.astype(np.uint8), g.astype(np.uint8)
