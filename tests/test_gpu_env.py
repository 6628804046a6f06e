"""GPU parity: the CUDA environment path (through the C ABI) against the golden vectors from the live
reference and against the C oracle on identical seeded inputs.  Bit-exact: positions, rewards (fp32),
done, observation bytes, heuristic maps, distances."""
import hashlib

import numpy as np
import pytest

from helpers import golden, greedy_actions, instances, random_instance
from oracle import oracle

pytestmark = pytest.mark.gpu


def sha8(b):
    return np.frombuffer(hashlib.sha256(b).digest()[:8], dtype=np.uint64)[0]


def make_env(B, N, L):
    from mapf_rl_b200 import BatchedEnvironment
    return BatchedEnvironment(B, N, L)


def test_native_library_loaded():
    import torch
    from mapf_rl_b200 import _native, build
    assert torch.cuda.is_available()
    _native.lib()
    with open("/proc/self/maps") as f:
        assert build.LIB_PATH in f.read()


def test_crafted_cases():
    z = golden("crafted.npz")
    for name in z["names"]:
        m, a, g = z[f"{name}_map"], z[f"{name}_agents"], z[f"{name}_goals"]
        env = make_env(1, a.shape[0], m.shape[0])
        env.load(m[None], a[None], g[None])
        obs, rew, done = env.step(z[f"{name}_actions"][None])
        assert np.array_equal(env.agents_pos[0].cpu().numpy(), z[f"{name}_pos"]), name
        assert np.array_equal(rew[0].cpu().numpy(), z[f"{name}_rewards"]), name
        assert int(done[0]) == z[f"{name}_done"], name
        assert np.array_equal(obs[0].cpu().numpy(), z[f"{name}_obs"]), name
        obs, rew, done = env.step(np.zeros((1, a.shape[0]), dtype=np.uint8))
        assert np.array_equal(rew[0].cpu().numpy(), z[f"{name}_rewards2"]), name
        assert int(done[0]) == z[f"{name}_done2"]
        assert int(env.steps[0]) - 1 == z[f"{name}_info2"]
        env.check()


@pytest.mark.parametrize("N", [16, 32, 64])
@pytest.mark.parametrize("stream", ["U", "G"])
def test_golden_traces(N, stream):
    z = golden("traces.npz")
    maps, agents, goals = instances(N)
    pre = f"n{N}_{stream}_"
    ks = z[pre + "instances"]
    env = make_env(len(ks), N, 40)
    env.load(maps[ks], agents[ks], goals[ks])
    obs, pos = env.observe()
    obs, pos = obs.cpu().numpy(), pos.cpu().numpy()
    for q in range(len(ks)):
        assert np.array_equal(pos[q], z[pre + "pos"][q, 0])
        assert sha8(obs[q].tobytes()) == z[pre + "obs_sha8"][q, 0]
    T = z[pre + "actions"].shape[1]
    for s in range(T):
        obs, rew, done = env.step(np.ascontiguousarray(z[pre + "actions"][:, s]))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        pos = env.agents_pos.cpu().numpy()
        for q in range(len(ks)):
            assert np.array_equal(pos[q], z[pre + "pos"][q, s + 1]), (q, s)
            assert np.array_equal(rew[q], z[pre + "rewards"][q, s]), (q, s)
            assert done[q] == z[pre + "done"][q, s]
            assert sha8(obs[q].tobytes()) == z[pre + "obs_sha8"][q, s + 1], (q, s)
    assert np.array_equal(env.steps.cpu().numpy(), np.full(len(ks), T))
    env.check()


@pytest.mark.parametrize("N", [16, 32, 64])
def test_navi_all_pkl_instances(N):
    z = golden("navi.npz")
    maps, agents, goals = instances(N)
    env = make_env(200, N, 40)
    env.load(maps, agents, goals)
    nv = env.navi_map.cpu().numpy()
    for k in range(200):
        got = np.frombuffer(hashlib.sha256(nv[k].tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(got, z[f"navi{N}_sha256"][k]), k
    assert np.array_equal(env.map.cpu().numpy(), maps)
    assert np.array_equal(env.goals_pos.cpu().numpy(), goals)


def test_distances_vs_compute_heuristics():
    z = golden("navi.npz")
    maps, agents, goals = instances(32)
    env = make_env(200, 32, 40)
    env.load(maps, agents, goals)
    dist = env.heuristic_distances(env_ids=[0, 199]).cpu().numpy()
    assert np.array_equal(dist[0], z["dist32_0"])
    assert np.array_equal(dist[1], z["dist32_199"])
    # recomputing must leave the heuristic maps unchanged
    nv = env.navi_map.cpu().numpy()
    got = np.frombuffer(hashlib.sha256(nv[199].tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(got, z["navi32_sha256"][199])


@pytest.mark.parametrize("N", [16, 32, 64])
@pytest.mark.parametrize("stream", ["U", "G"])
def test_full_pkl_rollout_vs_oracle(N, stream):
    """All 200 instances x 256 steps (config.max_steps), every step compared with the C oracle."""
    maps, agents, goals = instances(N)
    B, T = 200, 256
    env = make_env(B, N, 40)
    env.load(maps, agents, goals)
    ora = []
    for k in range(B):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        ora.append(o)
    rng = np.random.default_rng(2024 + N)
    obs, pos = env.observe()
    obs = obs.cpu().numpy()
    for k in range(B):
        oo, op = ora[k].observe()
        assert np.array_equal(oo.astype(np.uint8), obs[k])
    collisions = 0
    for s in range(T):
        acts = rng.integers(0, 5, size=(B, N)).astype(np.uint8) if stream == "U" else greedy_actions(obs, rng)
        g_obs, g_rew, g_done = env.step(acts)
        obs, g_rew, g_done = g_obs.cpu().numpy(), g_rew.cpu().numpy(), g_done.cpu().numpy()
        g_pos = env.agents_pos.cpu().numpy()
        for k in range(B):
            (o_obs, o_pos), o_rew, o_done, info = ora[k].step(acts[k])
            assert np.array_equal(o_pos, g_pos[k]), (k, s)
            assert np.array_equal(np.asarray(o_rew, dtype=np.float32), g_rew[k]), (k, s)
            assert int(o_done) == g_done[k], (k, s)
            assert np.array_equal(o_obs.astype(np.uint8), obs[k]), (k, s)
        collisions += int((g_rew == -0.5).sum())
        # invariant of environment.py:424-428
        cells = g_pos[..., 0].astype(np.int32) * 40 + g_pos[..., 1]
        assert all(len(np.unique(c)) == N for c in cells)
    assert collisions > 0
    env.check()


@pytest.mark.parametrize("L,N,density,occupancy", [
    (3, 5, 0.0, None), (4, 9, 0.1, None), (6, 30, 0.0, None), (8, 33, 0.05, None), (12, 64, 0.1, None),
    (10, 7, 0.3, None), (20, 6, 0.3, None), (33, 37, 0.25, None), (56, 40, 0.3, None), (57, 12, 0.3, None),
    (80, 64, 0.3, None), (88, 100, 0.2, None), (89, 8, 0.3, None), (120, 128, 0.1, None), (7, 45, 0.0, None),
    (28, 10, 0.2, None), (72, 33, 0.3, None),
])
def test_random_grids_vs_oracle(L, N, density, occupancy):
    """Generic sizes: every RW (row words) / K (agents per lane) template, unaligned observation blocks,
    dense boards (heavy swap / vertex / chain conflicts) and all-same-direction streams."""
    rng = np.random.default_rng(L * 1000 + N)
    B, T = 24, 40
    ms, as_, gs = [], [], []
    for _ in range(B):
        m, a, g = random_instance(rng, L, N, density)
        ms.append(m), as_.append(a), gs.append(g)
    ms, as_, gs = np.stack(ms), np.stack(as_), np.stack(gs)
    if B > 2:
        gs[1] = as_[1]  # everyone starts on its goal: exercises finish / stay_on_goal
    env = make_env(B, N, L)
    env.load(ms, as_, gs)
    nv = env.navi_map.cpu().numpy()
    ora = []
    for k in range(B):
        o = oracle.OracleEnv()
        o.load(ms[k], as_[k], gs[k])
        assert np.array_equal(o.navi_map, nv[k]), k
        ora.append(o)
    dist = env.heuristic_distances().cpu().numpy()
    for k in range(B):
        assert np.array_equal(ora[k].dist_map, dist[k]), k
    obs, pos = env.observe()
    obs = obs.cpu().numpy()
    for k in range(B):
        assert np.array_equal(ora[k].observe()[0].astype(np.uint8), obs[k])
    for s in range(T):
        mode = s % 4
        if mode == 0:
            acts = rng.integers(0, 5, size=(B, N))
        elif mode == 1:
            acts = np.broadcast_to(rng.integers(1, 5, size=(B, 1)), (B, N))
        elif mode == 2:
            acts = greedy_actions(obs, rng, eps=0.2)
        else:
            acts = rng.integers(1, 5, size=(B, N))
        acts = np.ascontiguousarray(acts, dtype=np.uint8)
        if s == T - 1:
            acts[:] = 0
        g_obs, g_rew, g_done = env.step(acts)
        obs, g_rew, g_done = g_obs.cpu().numpy(), g_rew.cpu().numpy(), g_done.cpu().numpy()
        g_pos = env.agents_pos.cpu().numpy()
        for k in range(B):
            (o_obs, o_pos), o_rew, o_done, info = ora[k].step(acts[k])
            assert np.array_equal(o_pos, g_pos[k]), (k, s)
            assert np.array_equal(np.asarray(o_rew, dtype=np.float32), g_rew[k]), (k, s)
            assert int(o_done) == g_done[k], (k, s)
            assert np.array_equal(o_obs.astype(np.uint8), obs[k]), (k, s)
    env.check()


def test_observe_corners_and_borders():
    L, N = 9, 6
    m = np.zeros((L, L), dtype=np.uint8)
    m[4, 4] = 1
    a = np.array([[0, 0], [0, L - 1], [L - 1, 0], [L - 1, L - 1], [0, 1], [4, 0]], dtype=np.uint8)
    g = np.array([[8, 8], [8, 0], [0, 8], [0, 0], [5, 5], [4, 8]], dtype=np.uint8)
    env = make_env(1, N, L)
    env.load(m[None], a[None], g[None])
    obs, pos = env.observe()
    obs = obs[0].cpu().numpy()
    o = oracle.OracleEnv()
    o.load(m, a, g)
    assert np.array_equal(o.observe()[0].astype(np.uint8), obs)
    assert obs[0, 1].sum() == 1 and obs[0, 1, 8, 8] == 1      # the obstacle at (4,4) seen from (0,0); outside = 0
    assert obs[0, 0, 4, 5] == 1 and obs[0, 0, 4, 4] == 0      # neighbour visible, own centre cleared
    assert obs[0, :, :4, :].sum() == 0 and obs[0, :, :, :4].sum() == 0   # everything outside the map is 0


def test_invalid_action_latches_assertion():
    m, a, g = random_instance(np.random.default_rng(0), 8, 4, 0.1)
    env = make_env(1, 4, 8)
    env.load(m[None], a[None], g[None])
    env.step(np.array([[0, 7, 1, 2]], dtype=np.uint8))
    with pytest.raises(AssertionError):
        env.check()
    env.check()  # cleared


def test_subset_reload_and_obs_into_replay_slot():
    import torch
    maps, agents, goals = instances(32)
    env = make_env(8, 32, 40)
    env.load(maps[:8], agents[:8], goals[:8])
    acts = np.random.default_rng(0).integers(0, 5, size=(8, 32)).astype(np.uint8)
    replay = torch.zeros((3, 8, 32, 6, 9, 9), dtype=torch.uint8, device=env.device)
    env.step(acts, out_obs=replay[1])
    env.load(maps[100:102], agents[100:102], goals[100:102], env_ids=[5, 2])
    assert np.array_equal(env.steps.cpu().numpy(), [1, 1, 0, 1, 1, 0, 1, 1])
    obs, pos = env.observe()
    o = oracle.OracleEnv()
    o.load(maps[101], agents[101], goals[101])
    assert np.array_equal(o.observe()[0].astype(np.uint8), obs[2].cpu().numpy())
    o.load(maps[3], agents[3], goals[3])
    (oo, _), _, _, _ = o.step(acts[3])
    assert np.array_equal(oo.astype(np.uint8), replay[1, 3].cpu().numpy())
    assert replay[0].sum() == 0 and replay[2].sum() == 0


def test_step_host_entry_point():
    maps, agents, goals = instances(16)
    env = make_env(4, 16, 40)
    env.load(maps[:4], agents[:4], goals[:4])
    acts = np.random.default_rng(1).integers(0, 5, size=(4, 16)).astype(np.uint8)
    obs, rew, done, steps = env.step_host(acts, want_obs=True)
    for k in range(4):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        (oo, op), orw, od, _ = o.step(acts[k])
        assert np.array_equal(oo.astype(np.uint8), obs[k])
        assert np.array_equal(np.asarray(orw, dtype=np.float32), rew[k])
        assert int(od) == done[k] and steps[k] == 1


@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_every_mode(mode, pinned):
    """Both forms of the fp32 host-buffer step (step kernel, then the observe kernel next to the result copies: issued call by
    call, and replayed from a CUDA graph) with page-locked and pageable caller buffers:
    several consecutive steps (graph replay with changing observation targets), bit-exact against the oracle."""
    import ctypes as C
    import torch
    from mapf_rl_b200 import _native
    lib = _native.lib()
    prev = lib.mapf_debug_step_host_mode(-1)
    try:
        assert lib.mapf_debug_step_host_mode(mode) == mode
        maps, agents, goals = instances(32)
        B, N = 70, 32
        env = make_env(B, N, 40)
        env.load(maps[:B], agents[:B], goals[:B])
        ora = []
        for k in range(B):
            o = oracle.OracleEnv()
            o.load(maps[k], agents[k], goals[k])
            ora.append(o)
        mk = (lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=True).numpy()) if pinned else \
             (lambda shape, dt: torch.zeros(shape, dtype=dt).numpy())
        acts, rew = mk((B, N), torch.uint8), mk((B, N), torch.float32)
        done, steps = mk((B,), torch.uint8), mk((B,), torch.int32)
        ring = torch.zeros((3, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda:0")
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rng = np.random.default_rng(mode)
        for s in range(7):
            acts[:] = rng.integers(0, 5, size=(B, N))
            slot = ring[s % 3]
            _native.check(lib.mapf_env_step_host(env._h, vp(acts), None, vp(rew), vp(done), vp(steps),
                                                 C.c_void_p(slot.data_ptr()), env._stream()))
            obs = slot.cpu().numpy()
            for k in range(B):
                (oo, op), orw, od, _ = ora[k].step(acts[k])
                assert np.array_equal(oo.astype(np.uint8), obs[k]), (mode, s, k)
                assert np.array_equal(np.asarray(orw, dtype=np.float32), rew[k]), (mode, s, k)
                assert int(od) == done[k] and steps[k] == s + 1
        env.check()
    finally:
        lib.mapf_debug_step_host_mode(prev)


def test_step_host_results_are_fresh_when_the_call_returns():
    """Many back-to-back mapf_env_step_host calls (graph replay over rotating page-locked action buffers and observation
    slots) at a size where the result copies take tens of microseconds: the host rewards / done / steps read right after
    each call are those of THAT step (a twin stepped on the device), the device observation that of the last."""
    import torch
    B, N, L, T = 4096, 32, 40, 120
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=17, density=0.3)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
    acts_host = acts.cpu().pin_memory()
    exp = []
    for t in range(T):
        o, r, d = twin.step(acts[t])
        exp.append((r.cpu().numpy().copy(), d.cpu().numpy().copy()))
    ring = torch.zeros((2, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
    for t in range(T):
        _, rew, done, steps = env.step_host(acts_host[t], device_obs=ring[t % 2])
        assert np.array_equal(rew, exp[t][0]), t
        assert np.array_equal(done, exp[t][1]), t
        assert int(steps.min()) == t + 1 and int(steps.max()) == t + 1
    assert torch.equal(ring[(T - 1) % 2], o)          # stream-ordered read of the last device observation
    assert torch.equal(env.agents_pos, twin.agents_pos)
    env.check()


def test_step_host_pageable_caller_buffers():
    """mapf_env_step_host with ordinary (pageable) numpy buffers goes through the handle's pinned staging area."""
    import ctypes as C
    from mapf_rl_b200 import _native
    maps, agents, goals = instances(32)
    env = make_env(3, 32, 40)
    env.load(maps[:3], agents[:3], goals[:3])
    acts = np.random.default_rng(5).integers(0, 5, size=(3, 32)).astype(np.uint8)
    rew = np.zeros((3, 32), dtype=np.float32)
    done = np.zeros(3, dtype=np.uint8)
    steps = np.zeros(3, dtype=np.int32)
    obs = np.zeros((3, 32, 6, 9, 9), dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _native.check(env._lib.mapf_env_step_host(env._h, vp(acts), vp(obs), vp(rew), vp(done), vp(steps), None, env._stream()))
    for k in range(3):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        (oo, op), orw, od, _ = o.step(acts[k])
        assert np.array_equal(oo.astype(np.uint8), obs[k])
        assert np.array_equal(np.asarray(orw, dtype=np.float32), rew[k])
        assert int(od) == done[k] and steps[k] == 1


def test_full_size_batch_properties():
    """BASELINE configs[1] size (8192 envs x 32 agents): size-independent properties + sampled oracle parity.
    The batch tiles the 200 pkl instances, so replicas of an instance driven by the same actions must agree
    bit for bit wherever they sit in the batch."""
    import torch
    maps, agents, goals = instances(32)
    B, N, T = 8192, 32, 24
    rep = np.arange(B) % 200
    env = make_env(B, N, 40)
    env.load(maps[rep], agents[rep], goals[rep])
    rng = np.random.default_rng(99)
    sample = rng.choice(B, size=48, replace=False)
    sidx = torch.as_tensor(sample, device="cuda:0")
    ora = {}
    for k in sample:
        o = oracle.OracleEnv()
        o.load(maps[rep[k]], agents[rep[k]], goals[rep[k]])
        ora[k] = o
    base_acts = rng.integers(0, 5, size=(T, 200, N)).astype(np.uint8)
    for s in range(T):
        acts = base_acts[s][rep]
        obs, rew, done = env.step(acts)
        pos = env.agents_pos
        # replicas agree (no cross-env interference, no dependence on the slot / CTA / warp an env lands in)
        o4 = obs.view(B, -1)
        assert torch.equal(o4[200:400], o4[:200]) and torch.equal(o4[8000:8192], o4[:192])
        assert torch.equal(rew[4000:4200], rew[:200]) and torch.equal(pos[8000:8192], pos[:192])
        # bool bytes only; the centre of channel 0 is cleared; agent cells unique (environment.py:424-428)
        assert int(obs.max().item()) <= 1
        assert int(obs[:, :, 0, 4, 4].sum().item()) == 0
        cells = pos[..., 0].to(torch.int64) * 40 + pos[..., 1].to(torch.int64)
        assert int((torch.sort(cells, dim=1).values.diff(dim=1) == 0).sum().item()) == 0
        # channel 0 counts ordered pairs of agents within each other's 9x9 window
        dx = (pos[:, :, None, 0].to(torch.int32) - pos[:, None, :, 0].to(torch.int32)).abs()
        dy = (pos[:, :, None, 1].to(torch.int32) - pos[:, None, :, 1].to(torch.int32)).abs()
        near = ((dx <= 4) & (dy <= 4)).sum((1, 2)) - N
        assert torch.equal(obs[:, :, 0].to(torch.int64).sum((1, 2, 3)), near)
        # observe() is idempotent and equals what step() returned
        if s % 8 == 0:
            again, p2 = env.observe()
            assert torch.equal(again, obs) and torch.equal(p2, pos)
        obs_h, rew_h, pos_h, done_h = (t[sidx].cpu().numpy() for t in (obs, rew, pos, done))
        for q, k in enumerate(sample):
            (oo, op), orw, od, _ = ora[k].step(acts[k])
            assert np.array_equal(oo.astype(np.uint8), obs_h[q]) and np.array_equal(op, pos_h[q])
            assert np.array_equal(np.asarray(orw, dtype=np.float32), rew_h[q]) and int(od) == done_h[q]
    env.check()
