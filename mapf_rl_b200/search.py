"""Host mirror of the reference's search.py (CBS expert, search.py:58-442) and of test.py's `create_test` (test.py:23-79)
over the C ABI (`mapf_cbs_solve`, `mapf_cbs_solve_batch`: host C++, include/mapf_b200.h).

    find_path(env)                 -> list of per-step action lists, or None        (search.py:396-442)
    compute_heuristics(map, goal)  -> {(x, y): distance}                            (search.py:24-55)
    solve / solve_batch            -> the same on arrays, with cost / node counts
    create_test(...)               -> dict(maps, agents, goals, opt_steps, opt_mean_steps): `test_num` instances the expert
                                      solves, drawn on the GPU (BatchedEnvironment.reset), heuristics from the BFS kernel

The expert's sum of costs equals the reference's (CBS is optimal); its paths are one of several optimal sets -- the reference
picks conflicts with random.choice and stops on wall-clock time, this search is deterministic."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native, config

DIST_UNREACHABLE = 2147483647


def _u8(a, shape):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int64)
    assert a.shape == shape, (a.shape, shape)
    if a.size and (a.min() < 0 or a.max() > 255):
        raise IndexError("coordinate outside the map")
    return np.ascontiguousarray(a.astype(np.uint8))


def solve_batch(maps, starts, goals, dist=None, max_steps: int = config.max_steps, time_limit_s: float = 5.0,
                node_limit: int = 1 << 20, max_T: Optional[int] = None, threads: int = 0):
    """n instances of one geometry.  maps [n, L, L] (non-zero = obstacle), starts / goals [n, N, 2], dist optional int32
    [n, N, L, L] (BatchedEnvironment.heuristic_distances()).  Returns (actions uint8 [n, max_T, N], T int32 [n] (-1: unsolved),
    cost int32 [n], expanded int64 [n])."""
    maps = np.asarray(maps)
    n, L = maps.shape[0], maps.shape[1]
    assert maps.shape == (n, L, L)
    m8 = np.ascontiguousarray((maps != 0).astype(np.uint8))
    N = np.asarray(starts).shape[1]
    s8, g8 = _u8(starts, (n, N, 2)), _u8(goals, (n, N, 2))
    if (s8 >= L).any() or (g8 >= L).any():
        raise IndexError("coordinate outside the map")
    max_T = int(max_T if max_T is not None else max_steps)
    dptr = None
    if dist is not None:
        dist = np.ascontiguousarray(np.asarray(dist), dtype=np.int32)
        assert dist.shape == (n, N, L, L)
        dptr = dist.ctypes.data_as(C.c_void_p)
    actions = np.zeros((n, max_T, N), dtype=np.uint8)
    T = np.full(n, -1, dtype=np.int32)
    cost = np.full(n, -1, dtype=np.int32)
    expanded = np.zeros(n, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _native.check(_native.lib().mapf_cbs_solve_batch(n, p(m8), L, N, p(s8), p(g8), dptr, int(max_steps), int(time_limit_s * 1000),
                                                     int(node_limit), p(actions), max_T, p(T), p(cost), p(expanded), int(threads)))
    return actions, T, cost, expanded


def solve(my_map, starts, goals, dist=None, **kw):
    """One instance: (actions uint8 [T, N] or None, cost, expanded)."""
    a, T, cost, ex = solve_batch(np.asarray(my_map)[None], np.asarray(starts)[None], np.asarray(goals)[None],
                                 None if dist is None else np.asarray(dist)[None], **kw)
    if T[0] < 0:
        return None, -1, int(ex[0])
    return a[0, :T[0]], int(cost[0]), int(ex[0])


def compute_heuristics(my_map, goal):
    """search.compute_heuristics (search.py:24-55): {(x, y): cost} for the cells that reach `goal` -- host BFS, for callers
    that want the reference's dict; the batched path takes its distances from the GPU (`heuristic_distances`)."""
    m = np.asarray(my_map) != 0
    L0, L1 = m.shape
    dist = {tuple(int(v) for v in goal): 0}
    frontier = [tuple(int(v) for v in goal)]
    while frontier:
        nxt = []
        for (x, y) in frontier:
            for dx, dy in ((0, -1), (1, 0), (0, 1), (-1, 0)):
                c = (x + dx, y + dy)
                if 0 <= c[0] < L0 and 0 <= c[1] < L1 and not m[c] and c not in dist:
                    dist[c] = dist[(x, y)] + 1
                    nxt.append(c)
        frontier = nxt
    return dist


def find_path(env, time_limit_s: float = 5.0, node_limit: int = 1 << 20):
    """search.find_path (search.py:396-442) for a drop-in `Environment` (or anything with map / agents_pos / goals_pos /
    num_agents): the expert's action script as a list of per-step lists of ints (a list of ints for one agent), or None."""
    acts, _, _ = solve(np.asarray(env.map), np.asarray(env.agents_pos), np.asarray(env.goals_pos), time_limit_s=time_limit_s,
                       node_limit=node_limit)
    if acts is None:
        return None
    if env.num_agents == 1:
        return [int(a[0]) for a in acts]
    return [[int(v) for v in a] for a in acts]


def create_test(num_agents: int, map_length: int, test_num: int = 200, density: Optional[float] = None, seed: int = 0,
                time_limit_s: float = 5.0, node_limit: int = 1 << 16, device=None, threads: int = 0, batch: int = 0):
    """test.create_test (test.py:23-79) for one (num_agents, map_length): `test_num` instances the expert solves, with the
    length of its script as `opt_steps`.  Instances are drawn in batches on the GPU by the device-side generator (the
    reference's distribution, environment.py:100-138), their per-agent distance maps come from the BFS kernel and feed the
    low-level search; unsolved instances are dropped and replaced from the next batch (the reference resets and retries,
    test.py:51-56)."""
    from .batched import BatchedEnvironment
    B = batch or max(64, min(1024, 2 * test_num))
    env = BatchedEnvironment(B, num_agents, map_length, device=device)
    tests = {"maps": [], "agents": [], "goals": [], "opt_steps": []}
    offset = 0
    while len(tests["maps"]) < test_num:
        env.reset(seed=seed, env_offset=offset, density=density)
        env.check()
        offset += B
        maps = env.map.cpu().numpy()
        pos = env.agents_pos.cpu().numpy()
        goals = env.goals_pos.cpu().numpy()
        dist = env.heuristic_distances().cpu().numpy()
        _, T, _, _ = solve_batch(maps, pos, goals, dist=dist, time_limit_s=time_limit_s, node_limit=node_limit, threads=threads)
        for k in np.flatnonzero(T >= 0):
            if len(tests["maps"]) == test_num:
                break
            tests["maps"].append(maps[k].astype(np.int64))
            tests["agents"].append(pos[k].astype(np.int64))
            tests["goals"].append(goals[k].astype(np.int64))
            tests["opt_steps"].append(int(T[k]))
    env.close()
    tests["opt_mean_steps"] = sum(tests["opt_steps"]) / len(tests["opt_steps"])
    return tests
