"""CBS expert / solvable-instance generator (search.py:58-442, test.py:23-79).

Parity.  Textbook CBS determines the SUM OF COSTS of its solution uniquely; the reference's variant (disjoint splitting that
re-plans only the constrained agent, conflict and agent drawn with random.choice, search.py:249,316,340-372) does not: on 9 of
the 24 fixture instances its cost depends on the seed of the `random` module (e.g. 94 ... 102), the smallest value being the
optimum whenever it finds it (tests/golden/cbs.npz keeps six seeds per instance; made by tests/golden/make_golden_cbs.py from
the live reference).  The host C++ search is textbook CBS, so it must (1) equal the reference's cost wherever the six seeds
agree, (2) never exceed the reference's best cost anywhere, (3) stay above the sum of the individual shortest paths, and (4)
return a valid solution: replayed through the oracle's Environment.step no agent ever collides and the episode finishes on the
script's last step."""
import os
import random

import numpy as np
import pytest

from cbs_cases import CBS_CASES, cbs_instance
from oracle import oracle, ref_loader

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cbs.npz"))


def _replay(m, starts, goals, acts):
    """Every step of the script through the oracle environment: no collision reward, done exactly at the end."""
    env = oracle.OracleEnv()
    env.load(m, starts, goals)
    done = bool(np.array_equal(starts, goals))
    for t, a in enumerate(acts):
        assert not done, "the script goes on after the episode has finished"
        (_, pos), rew, done, _ = env.step(np.asarray(a, dtype=np.uint8))
        if not done:
            assert min(rew) > -0.5 + 1e-6, (t, rew)  # config.reward_fn['collision'] = -0.5
    assert done and np.array_equal(np.asarray(env.agents_pos), goals)


def _cost(starts, acts):
    """sum over agents of the last step at which the agent is not yet parked on its goal for good (search.py:17-21)."""
    acts = np.asarray(acts)
    T, N = acts.shape
    cost = 0
    for a in range(N):
        moving = np.flatnonzero(acts[:, a] != 0)
        cost += int(moving[-1]) + 1 if len(moving) else 0
    return cost


@pytest.mark.parametrize("k", range(len(CBS_CASES)))
def test_cost_vs_reference_and_script_is_valid(k):
    from mapf_rl_b200 import search
    L, N, density = CBS_CASES[k]
    m, starts, goals = cbs_instance(k, L, N, density)
    acts, cost, expanded = search.solve(m, starts, goals, time_limit_s=0, node_limit=1 << 18)
    assert acts is not None
    ref = GOLD["cost"][k]
    assert (ref >= 0).all()
    assert cost <= int(ref.min()), (cost, ref.tolist())
    if ref.min() == ref.max():
        assert cost == int(ref[0]), (cost, ref.tolist())
    # the script's own cost: an agent's path ends with its last move (waits at the goal are free), but a wait that is followed
    # by a move counts -- so the replayed cost can only be <= the reported one when paths end with explicit waits
    assert _cost(starts, acts) <= cost
    _replay(m, starts, goals, acts)
    # makespans differ between optimal solutions, but never below the longest individual shortest path
    h = [search.compute_heuristics(m, tuple(g))[tuple(s)] for s, g in zip(starts, goals)]
    assert len(acts) >= max(h) and cost >= sum(h)


def test_find_path_mirror_and_batch_threads():
    from mapf_rl_b200 import search

    class Env:  # what search.find_path reads (search.py:398-402,421)
        pass
    L, N, density = CBS_CASES[3]
    m, starts, goals = cbs_instance(3, L, N, density)
    env = Env()
    env.map, env.agents_pos, env.goals_pos, env.num_agents = m, starts, goals, N
    actions = search.find_path(env)
    assert isinstance(actions, list) and isinstance(actions[0], list) and len(actions[0]) == N
    assert all(isinstance(v, int) and 0 <= v <= 4 for row in actions for v in row)
    _replay(m, starts, goals, actions)
    # one agent: a flat list of ints (search.py:437-438)
    env.agents_pos, env.goals_pos, env.num_agents = starts[:1], goals[:1], 1
    flat = search.find_path(env)
    assert isinstance(flat[0], int) and len(flat) == search.compute_heuristics(m, tuple(goals[0]))[tuple(starts[0])]
    # batch on several threads == one by one
    idx = [k for k, c in enumerate(CBS_CASES) if c[0] == 12 and c[1] == 6] * 3 + [7, 7]
    inst = [cbs_instance(7, 12, 6, 0.2)] * len(idx)
    maps, ss, gg = (np.stack(x) for x in zip(*inst))
    a4, T4, c4, _ = search.solve_batch(maps, ss, gg, time_limit_s=0, threads=4)
    a1, T1, c1, _ = search.solve_batch(maps, ss, gg, time_limit_s=0, threads=1)
    assert np.array_equal(a4, a1) and np.array_equal(T4, T1) and np.array_equal(c4, c1) and (c4 == int(GOLD["cost"][7][0])).all()


def test_unsolvable_and_limits():
    from mapf_rl_b200 import search
    # two agents that must swap in a corridor one cell wide: no solution; the node budget ends the search
    m = np.ones((3, 5), dtype=np.int64)
    m = np.ones((5, 5), dtype=np.int64)
    m[2, :] = 0
    acts, cost, expanded = search.solve(m, [[2, 0], [2, 4]], [[2, 4], [2, 0]], time_limit_s=0, node_limit=200, max_steps=16)
    assert acts is None and cost == -1 and expanded >= 200
    # goal unreachable for the low level (search.py:309 asserts; we report "no solution")
    m2 = np.zeros((5, 5), dtype=np.int64)
    m2[:, 2] = 1
    acts, _, _ = search.solve(m2, [[0, 0]], [[0, 4]], time_limit_s=0)
    assert acts is None
    # start on an obstacle
    with pytest.raises(Exception):
        search.solve(m2, [[0, 2]], [[0, 0]])
    # an agent already on its goal with nobody in the way: empty script
    acts, cost, _ = search.solve(np.zeros((4, 4), dtype=np.int64), [[1, 1]], [[1, 1]])
    assert acts is not None and len(acts) == 0 and cost == 0


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_cost_vs_live_reference_fresh_instances():
    """Instances that are NOT in the fixture, solved by the live reference right here."""
    from mapf_rl_b200 import search
    ref = ref_loader.load_module("search")
    n = 0
    for k in range(100, 112):
        L, N, density = [(7, 4, 0.1), (9, 5, 0.2), (11, 6, 0.15)][k % 3]
        m, starts, goals = cbs_instance(k, L, N, density)
        random.seed(k)
        paths = ref.CBSSolver(m.copy(), [tuple(int(v) for v in s) for s in starts], [tuple(int(v) for v in g) for g in goals]).find_solution()
        if paths is None:
            continue
        acts, cost, _ = search.solve(m, starts, goals, time_limit_s=0, node_limit=1 << 18)
        lower = sum(ref.compute_heuristics(m, tuple(int(v) for v in g))[tuple(int(v) for v in s)] for s, g in zip(starts, goals))
        assert lower <= cost <= ref.get_sum_of_cost(paths), k   # (equal unless this seed sent the reference astray)
        # ... and the heuristic table is the reference's (search.py:24-55)
        assert search.compute_heuristics(m, tuple(goals[0])) == {tuple(int(v) for v in c): int(d) for c, d in
                                                                 ref.compute_heuristics(m, tuple(int(v) for v in goals[0])).items()}
        n += 1
    assert n >= 8


@pytest.mark.gpu
def test_create_test_on_gpu():
    """test.create_test (test.py:23-79): instances from the device generator, heuristics from the BFS kernel, solved on host
    threads; every kept instance's script replays to `done` through the batched CUDA environment in exactly opt_steps."""
    import torch
    from mapf_rl_b200 import BatchedEnvironment, search
    tests = search.create_test(6, 12, test_num=24, density=0.2, seed=5, batch=64, time_limit_s=0, node_limit=1 << 14)
    assert len(tests["maps"]) == 24 and len(tests["opt_steps"]) == 24
    assert abs(tests["opt_mean_steps"] - np.mean(tests["opt_steps"])) < 1e-9
    maps, pos, goals = np.stack(tests["maps"]), np.stack(tests["agents"]), np.stack(tests["goals"])
    # distances from the GPU and from the host search give the same scripts
    env = BatchedEnvironment(24, 6, 12)
    env.load(maps, pos, goals)
    dist = env.heuristic_distances().cpu().numpy()
    a_gpu, T_gpu, c_gpu, _ = search.solve_batch(maps, pos, goals, dist=dist, time_limit_s=0)
    a_host, T_host, c_host, _ = search.solve_batch(maps, pos, goals, time_limit_s=0)
    assert np.array_equal(T_gpu, T_host) and np.array_equal(c_gpu, c_host) and np.array_equal(a_gpu, a_host)
    assert np.array_equal(T_gpu, np.asarray(tests["opt_steps"]))
    # replay: padded with stays (agents parked on their goals), every environment reports done exactly at its own last step
    Tmax = int(T_gpu.max())
    first_done = np.full(24, -1)
    for t in range(Tmax):
        obs, rew, done = env.step(torch.as_tensor(a_gpu[:, t]).cuda())
        d = done.cpu().numpy()
        r = rew.cpu().numpy()
        assert (r[d == 0] > -0.5 + 1e-6).all(), t
        first_done[(first_done < 0) & (d != 0)] = t + 1
    assert np.array_equal(first_done, T_gpu)
    env.check()
