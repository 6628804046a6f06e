"""GPU parity of mapf_env_rollout (T scripted steps as independent sub-batch chains on internal streams): the same
golden traces from the live reference, the C oracle on seeded inputs, and a twin handle stepped one launch at a
time.  Bit-exact for every chain count, ragged sub-batches and cyclic output rings."""
import hashlib

import numpy as np
import pytest

from helpers import golden, instances, random_instance
from oracle import oracle

pytestmark = pytest.mark.gpu


def sha8(b):
    return np.frombuffer(hashlib.sha256(b).digest()[:8], dtype=np.uint64)[0]


def make_env(B, N, L):
    from mapf_rl_b200 import BatchedEnvironment
    return BatchedEnvironment(B, N, L)


@pytest.mark.parametrize("N,chains", [(16, 0), (32, 3), (64, 2)])
@pytest.mark.parametrize("stream", ["U", "G"])
def test_golden_traces_through_rollout(N, stream, chains):
    """The traces recorded from the live reference (tests/golden/make_golden.py), all steps in ONE rollout call."""
    import torch
    z = golden("traces.npz")
    maps, agents, goals = instances(N)
    pre = f"n{N}_{stream}_"
    ks = z[pre + "instances"]
    env = make_env(len(ks), N, 40)
    env.load(maps[ks], agents[ks], goals[ks])
    acts = torch.as_tensor(np.ascontiguousarray(z[pre + "actions"].transpose(1, 0, 2))).cuda()  # [T,B,N]
    T = acts.shape[0]
    obs, rew, done, steps = env.rollout(acts, chains=chains)
    obs, rew, done, steps = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), steps.cpu().numpy()
    for s in range(T):
        for q in range(len(ks)):
            assert np.array_equal(rew[s, q], z[pre + "rewards"][q, s]), (q, s)
            assert done[s, q] == z[pre + "done"][q, s]
            assert sha8(obs[s, q].tobytes()) == z[pre + "obs_sha8"][q, s + 1], (q, s)
        assert np.array_equal(steps[s], np.full(len(ks), s + 1))
    assert np.array_equal(env.agents_pos.cpu().numpy(), z[pre + "pos"][:, T])
    env.check()


@pytest.mark.parametrize("B,N,L,chains", [(37, 20, 24, 1), (37, 20, 24, 2), (37, 20, 24, 5), (37, 20, 24, 16),
                                          (3, 8, 12, 4), (130, 40, 40, 7), (64, 96, 56, 3)])
def test_rollout_vs_oracle_and_twin(B, N, L, chains):
    """Ragged sub-batches (B not a multiple of chains or of 4), K > 1 agent slots, cyclic rings shorter than T."""
    import torch
    rng = np.random.default_rng(B * 1000 + N + chains)
    insts = [random_instance(rng, L, N, 0.25) for _ in range(B)]
    maps = np.stack([i[0] for i in insts])
    agents = np.stack([i[1] for i in insts])
    goals = np.stack([i[2] for i in insts])
    T, A, R, S = 11, 4, 3, 5
    acts = rng.integers(0, 5, size=(A, B, N)).astype(np.uint8)
    env, twin = make_env(B, N, L), make_env(B, N, L)
    env.load(maps, agents, goals)
    twin.load(maps, agents, goals)
    dev = env.device
    obs_ring = torch.zeros((R, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    rew_ring = torch.zeros((S, B, N), dtype=torch.float32, device=dev)
    done_ring = torch.zeros((S, B), dtype=torch.uint8, device=dev)
    steps_ring = torch.zeros((S, B), dtype=torch.int32, device=dev)
    d_acts = torch.as_tensor(acts).to(dev)
    env.rollout(d_acts, num_steps=T, out_obs=obs_ring, out_rewards=rew_ring, out_done=done_ring, out_steps=steps_ring,
                chains=chains)
    # twin: one launch per step, keeping what the rings must hold at the end
    exp_obs, exp_rew, exp_done = {}, {}, {}
    n_or = min(B, 6)
    ora = []
    for k in range(n_or):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        ora.append(o)
    for t in range(T):
        o, r, d = twin.step(d_acts[t % A])
        exp_obs[t % R], exp_rew[t % S], exp_done[t % S] = o.clone(), r.clone(), d.clone()
        on, rn, dn = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for k in range(n_or):
            (oo, _), orw, od, _ = ora[k].step(acts[t % A, k])
            assert np.array_equal(oo.astype(np.uint8), on[k]), (t, k)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), rn[k]), (t, k)
            assert bool(od) == bool(dn[k])
    for s in range(R):
        assert torch.equal(obs_ring[s], exp_obs[s]), s
    for s in range(S):
        assert torch.equal(rew_ring[s], exp_rew[s]), s
        assert torch.equal(done_ring[s], exp_done[s]), s
        last_t = max(t for t in range(T) if t % S == s)
        assert torch.equal(steps_ring[s], torch.full((B,), last_t + 1, dtype=torch.int32, device=dev))
    assert torch.equal(env.agents_pos, twin.agents_pos)
    assert torch.equal(env.steps, twin.steps)
    env.check()
    twin.check()


def test_rollout_then_step_and_reset_interleave():
    """A rollout joins back into the caller's stream: a plain step / masked reset / observe queued right after it
    sees its results, and a second rollout continues from them."""
    import torch
    B, N, L = 96, 32, 40
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=5, density=0.3)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    acts = torch.randint(0, 5, (6, B, N), generator=g, device="cuda", dtype=torch.uint8)
    env.rollout(acts[:3], chains=4)
    for t in range(3):
        twin.step(acts[t])
    o1, r1, d1 = env.step(acts[3])
    o2, r2, d2 = twin.step(acts[3])
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    mask = (torch.arange(B, device="cuda") % 3 == 0).to(torch.uint8)
    env.reset(mask=mask, seed=9, density=0.3)
    twin.reset(mask=mask, seed=9, density=0.3)
    ob, rw, dn, st = env.rollout(acts[4:], chains=3)
    for t in (4, 5):
        o2, r2, d2 = twin.step(acts[t])
        assert torch.equal(ob[t - 4], o2) and torch.equal(rw[t - 4], r2) and torch.equal(dn[t - 4], d2)
    assert torch.equal(env.observe()[0], twin.observe()[0])
    assert torch.equal(env.steps, twin.steps)
    env.check()


def test_rollout_full_size_equals_single_launch_steps():
    """BASELINE configs[1] size (8192 x 32, 40 x 40): 12 steps as 4 and as 8 chains equal 12 whole-batch launches."""
    import torch
    B, N, L, T = 8192, 32, 40, 12
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
    ref = make_env(B, N, L)
    ref.reset(seed=11, density=0.3)
    exp = [tuple(x.clone() for x in ref.step(acts[t])) for t in range(T)]
    for chains in (4, 8):
        env = make_env(B, N, L)
        env.reset(seed=11, density=0.3)
        obs, rew, done, steps = env.rollout(acts, chains=chains)
        for t in range(T):
            assert torch.equal(obs[t], exp[t][0]), (chains, t)
            assert torch.equal(rew[t], exp[t][1]) and torch.equal(done[t], exp[t][2])
        assert torch.equal(env.agents_pos, ref.agents_pos)
        assert int(steps[-1].min()) == T and int(steps[-1].max()) == T
        env.check()
        env.close()
        del obs


@pytest.mark.parametrize("chains", [0, 3])
def test_rollout_long_enough_for_graph_replay(chains):
    """T >= 4 slot periods on >= 2048 envs: whole periods are replayed from per-chain captured graphs (twice, so the
    second call hits the cache), the tail is launched directly; different rings afterwards rebuild the graphs."""
    import torch
    B, N, L = 2050, 8, 16
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=21, density=0.2)
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    A, R, S, T = 3, 2, 1, 29     # period lcm(3, 2, 1) = 6: 4 periods replayed + 5 direct steps
    acts = torch.randint(0, 5, (A, B, N), generator=g, device="cuda", dtype=torch.uint8)
    rings = [tuple(torch.zeros(sh, dtype=dt, device="cuda") for sh, dt in
                   (((R, B, N, 6, 9, 9), torch.uint8), ((S, B, N), torch.float32), ((S, B), torch.uint8), ((S, B), torch.int32)))
             for _ in range(2)]
    t_glob = 0
    for rnd, ring in enumerate((rings[0], rings[0], rings[1])):
        obs, rew, done, steps = ring
        env.rollout(acts, num_steps=T, out_obs=obs, out_rewards=rew, out_done=done, out_steps=steps, chains=chains)
        exp = {}
        for t in range(T):
            o, r, d = twin.step(acts[t % A])
            exp[t % R] = o.clone()
            t_glob += 1
        for s in range(R):
            assert torch.equal(obs[s], exp[s]), (rnd, s)
        assert torch.equal(rew[0], r) and torch.equal(done[0], d)
        assert torch.equal(env.agents_pos, twin.agents_pos)
        assert torch.equal(env.steps, twin.steps)
    env.check()


@pytest.fixture
def rollout_tuning():
    """Sets the knobs of the persistent rollout kernel for one test and restores the defaults afterwards."""
    from mapf_rl_b200 import _native
    lib = _native.lib()
    yield lambda persistent=-1, warps_per_sm=-1, chunk=-1, store_mode=-1, stagger_ns=-1: \
        lib.mapf_debug_rollout_tuning(persistent, warps_per_sm, chunk, store_mode, stagger_ns)
    lib.mapf_debug_rollout_tuning(1, 0, 0, 0, 1000)


def _rollout_vs_twin(env, twin, acts, T, A, R, S, want_codes=False, oracle_envs=(0, 1)):
    """env.rollout (one call) against `twin` stepped launch by launch and, for a few environments, the oracle."""
    import torch
    B, N = env.num_envs, env.num_agents
    obs = torch.zeros((R, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
    rew = torch.zeros((S, B, N), dtype=torch.float32, device="cuda")
    codes = torch.full((S, B, N), 255, dtype=torch.uint8, device="cuda") if want_codes else None
    done = torch.zeros((S, B), dtype=torch.uint8, device="cuda")
    steps = torch.zeros((S, B), dtype=torch.int32, device="cuda")
    maps, pos0, goals = env.map.cpu().numpy(), env.agents_pos.cpu().numpy(), env.goals_pos.cpu().numpy()
    steps0 = env.steps.clone()
    env.rollout(acts, num_steps=T, out_obs=obs, out_rewards=rew, out_done=done, out_steps=steps, out_codes=codes)
    ora = []
    for k in sorted(set(min(k, B - 1) for k in oracle_envs) | {B - 1}):
        o = oracle.OracleEnv()
        o.load(maps[k], pos0[k], goals[k])
        ora.append((k, o))
    exp_obs, exp_rew, exp_done = {}, {}, {}
    a_host = acts.cpu().numpy()
    for t in range(T):
        o, r, d = twin.step(acts[t % A])
        exp_obs[t % R], exp_rew[t % S], exp_done[t % S] = o.clone(), r.clone(), d.clone()
        for k, oe in ora:
            (oo, _), orw, od, _ = oe.step(a_host[t % A, k])
            assert np.array_equal(oo.astype(np.uint8), o[k].cpu().numpy()), (t, k)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), r[k].cpu().numpy()), (t, k)
    for s in range(R):
        bad = (obs[s] != exp_obs[s]).flatten(1).any(1).nonzero().flatten().tolist()
        assert not bad, (s, bad[:8], len(bad))
    table = torch.as_tensor(env.reward_table, device="cuda")
    for s in range(S):
        assert torch.equal(rew[s], exp_rew[s]) and torch.equal(done[s], exp_done[s]), s
        if want_codes:
            assert int(codes[s].max()) <= 4 and torch.equal(table[codes[s].long()], exp_rew[s]), s
        last_t = max(t for t in range(T) if t % S == s)
        assert torch.equal(steps[s], steps0 + (last_t + 1))
    assert torch.equal(env.agents_pos, twin.agents_pos) and torch.equal(env.steps, twin.steps)


@pytest.mark.parametrize("B,N,L,warps,chunk,store", [
    (2051, 20, 30, 0, 0, 0),     # one item per environment, direct stores, blocks not 16-byte aligned
    (2051, 32, 40, 0, 4, 0),     # five chunks per environment: hand-over through the progress words
    (2051, 32, 40, 8, 7, 1),     # bulk (TMA) stores from the staging block, ragged last chunk
    (333, 24, 56, 0, 3, 1),      # bulk stores, N % 8 == 0, largest map of the two-word row class
    (300, 7, 25, 4, 5, 1),       # bulk requested but the block is not 16-byte aligned: direct stores serve it
    (700, 64, 40, 0, 6, 0),      # two agents per lane (C3 geometry)
    (700, 48, 40, 0, 0, 1),      # two agents per lane, second slot half empty, bulk stores
    (260, 64, 80, 0, 5, 0),      # three-word rows, two agents per lane (C4 geometry)
    (260, 40, 64, 0, 0, 1),      # three-word rows, bulk
    (150, 16, 100, 0, 4, 0),     # four-word rows
    (90, 6, 16, 0, 2, 1),        # one-word rows (config.py defaults are 20x20)
    (5, 3, 8, 0, 3, 0),          # fewer environments than warps in a CTA pair
])
def test_rollout_persistent_kernel(rollout_tuning, B, N, L, warps, chunk, store):
    """chains = 0: ONE launch of the persistent kernel (work items = environment x chunk of steps, claimed time-major;
    navi tiles cached in shared memory).  Every row class, one and two agents per lane, both store forms, chunked and
    unchunked, cyclic rings shorter than T; bit-exact against a twin stepped launch by launch and against the oracle."""
    import torch
    T, A, R, S = 19, 3, 2, 2
    rollout_tuning(1, warps, chunk, store)
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=31 + N, density=0.25)
    assert env.rollout_plan(T, A, R, S)[0] == 0          # the persistent kernel takes it
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    acts = torch.randint(0, 5, (A, B, N), generator=g, device="cuda", dtype=torch.uint8)
    _rollout_vs_twin(env, twin, acts, T, A, R, S, want_codes=(chunk % 2 == 1))
    # a plain step right after it sees the rollout's state; a second rollout re-uses the (re-armed) scheduler words;
    # the chained form gives the same results
    o1, r1, d1 = env.step(acts[0])
    o2, r2, d2 = twin.step(acts[0])
    assert torch.equal(o1, o2) and torch.equal(r1, r2)
    _rollout_vs_twin(env, twin, acts, 5, A, R, S)
    rollout_tuning(0)
    assert env.rollout_plan(T, A, R, S)[0] > 0
    _rollout_vs_twin(env, twin, acts, 4, A, R, S)
    env.check()
    twin.check()


def test_rollout_persistent_graph_replay(rollout_tuning):
    """The launch re-arms its own scheduler words, so it can be captured once and replayed."""
    import torch
    B, N, L, T = 600, 16, 24, 6
    rollout_tuning(1, 0, 2, 0)
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=3, density=0.2)
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
    rings = [torch.zeros(sh, dtype=dt, device="cuda") for sh, dt in
             (((T, B, N, 6, 9, 9), torch.uint8), ((T, B, N), torch.float32), ((T, B), torch.uint8), ((T, B), torch.int32))]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.rollout(acts, out_obs=rings[0], out_rewards=rings[1], out_done=rings[2], out_steps=rings[3])
    side.synchronize()
    for t in range(T):
        twin.step(acts[t])
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        env.rollout(acts, out_obs=rings[0], out_rewards=rings[1], out_done=rings[2], out_steps=rings[3])
    for it in range(3):
        acts.copy_(torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8))
        graph.replay()
        for t in range(T):
            o, r, d = twin.step(acts[t])
            assert torch.equal(rings[0][t], o) and torch.equal(rings[1][t], r), (it, t)
        assert torch.equal(env.agents_pos, twin.agents_pos) and torch.equal(env.steps, twin.steps)
    env.check()


@pytest.mark.parametrize("store", [0, 1])
def test_rollout_persistent_full_size(rollout_tuning, store):
    """BASELINE configs[1] size through the default path (persistent kernel): 16 steps equal 16 whole-batch launches."""
    import torch
    B, N, L, T = 8192, 32, 40, 16
    rollout_tuning(1, 0, 0, store)
    g = torch.Generator(device="cuda")
    g.manual_seed(13)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
    ref = make_env(B, N, L)
    ref.reset(seed=19, density=0.3)
    env = make_env(B, N, L)
    env.reset(seed=19, density=0.3)
    assert env.rollout_plan(T, T, T, T)[0] == 0
    obs, rew, done, steps = env.rollout(acts)
    for t in range(T):
        o, r, d = ref.step(acts[t])
        assert torch.equal(obs[t], o), t
        assert torch.equal(rew[t], r) and torch.equal(done[t], d)
    assert torch.equal(env.agents_pos, ref.agents_pos)
    assert int(steps[-1].min()) == T and int(steps[-1].max()) == T
    env.check()


@pytest.mark.parametrize("B,N,L,name", [(8192, 64, 40, "C3"), (4096, 64, 80, "C4")])
def test_rollout_full_size_c3_c4_vs_oracle(rollout_tuning, B, N, L, name):
    """BASELINE configs[2] / [3] geometry at full size through the persistent kernel, navi-greedy actions (the
    congestion-heavy stream): 48 sampled environments, every step, against the oracle (heuristic maps included), and the
    invariants of the rest (bool bytes, own centre clear, one agent per cell)."""
    import torch
    torch.manual_seed(7)   # the action stream below draws from torch's default CUDA generator
    T = 10
    rollout_tuning(1, 0, 4 if name == "C4" else 0, 0)
    env = make_env(B, N, L)
    env.reset(seed=41, density=0.3)
    env.check()
    sample = np.unique(np.r_[0, 1, B - 1, np.random.default_rng(5).integers(0, B, size=45)])
    maps, pos0, goals = (x.cpu().numpy() for x in (env.map, env.agents_pos, env.goals_pos))
    navi = env.navi_map[torch.as_tensor(sample, device="cuda")].cpu().numpy()
    ora = []
    for q, k in enumerate(sample):
        o = oracle.OracleEnv()
        o.load(maps[k], pos0[k], goals[k])
        assert np.array_equal(o.navi_map, navi[q]), (name, k)
        ora.append(o)
    # navi-greedy with epsilon 0.1: follow a set direction bit of the agent's own cell, else stay
    rng = np.random.default_rng(7)
    obs0, _ = env.observe()
    acts = torch.empty((T, B, N), dtype=torch.uint8, device="cuda")
    twin = make_env(B, N, L)
    twin.reset(seed=41, density=0.3)
    cur = obs0
    for t in range(T):                                               # actions from the twin's observations (centre of ch 2..5)
        bits = cur[:, :, 2:6, 4, 4].float() + torch.rand((B, N, 4), device="cuda") * 0.5
        greedy = torch.where(cur[:, :, 2:6, 4, 4].any(-1), bits.argmax(-1) + 1, torch.zeros((B, N), dtype=torch.long, device="cuda"))
        eps = torch.rand((B, N), device="cuda") < 0.1
        a = torch.where(eps, torch.randint(0, 5, (B, N), device="cuda"), greedy).to(torch.uint8)
        acts[t] = a
        cur, _, _ = twin.step(a)
    obs, rew, done, steps = env.rollout(acts)
    a_host = acts.cpu().numpy()
    idx = torch.as_tensor(sample, device="cuda")
    n_coll = 0
    for t in range(T):
        ob, rw, dn = obs[t][idx].cpu().numpy(), rew[t][idx].cpu().numpy(), done[t][idx].cpu().numpy()
        for q, k in enumerate(sample):
            (oo, op), orw, od, _ = ora[q].step(a_host[t, k])
            assert np.array_equal(oo.astype(np.uint8), ob[q]), (name, t, k)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), rw[q]) and int(od) == dn[q], (name, t, k)
        n_coll += int((rw == -0.5).sum())
        assert int(obs[t].max()) <= 1 and int(obs[t][:, :, 0, 4, 4].max()) == 0
    assert n_coll > 50, "the greedy stream must exercise the conflict rounds"
    assert torch.equal(env.agents_pos, twin.agents_pos)
    pos = env.agents_pos.long()
    cell = pos[..., 0] * L + pos[..., 1]
    assert int((cell.sort(dim=1).values.diff(dim=1) == 0).sum()) == 0   # environment.py:424-428
    env.check()


@pytest.mark.parametrize("B,N,L,cap,chunk,store", [(300, 1, 6, 9, 0, 0), (257, 2, 7, 6, 3, 1), (600, 8, 12, 5, 4, 0),
                                                     (96, 32, 40, 4, 2, 1), (40, 64, 40, 3, 5, 0), (24, 40, 64, 4, 0, 0)])
@pytest.mark.parametrize("pregen,tasks", [(2, 0), (3, 1), (0, 0), (0, 1)])
def test_rollout_autoreset_vs_twin(rollout_tuning, B, N, L, cap, chunk, store, pregen, tasks):
    """Episode handling inside the launch (worker.py:390,422-428): a step that finds its environment finished -- all agents
    on their goals after the previous step (tiny boards reach that within a few random steps), or `cap` steps taken --
    re-generates the slot and emits the first observation.  The twin does the same through the public pieces:
    reset(mask, env_offset = base + n * stride) + observe, else step.
    pregen = 2: the first episode end of a slot in a launch adopts an instance staged by the dedicated generator / BFS kernels
    BEFORE the launch (double-buffered heuristic maps), later ones re-generate inside the kernel -- generator by the slot's
    warp, the per-agent searches as tasks any warp takes; 3: staged BESIDE the launch on a second stream (adopted if ready in
    time); 0: always inside."""
    import torch
    from mapf_rl_b200 import _native
    T, A = 23, 5
    seed, base, stride, density = 77, 1000, 4096, 0.2
    rollout_tuning(1, 0, chunk, store)
    _native.lib().mapf_debug_rollout_pregen(pregen)
    _native.lib().mapf_debug_rollout_tasks(tasks)   # in-launch searches as tasks any warp takes / on the slot's own warp
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=seed, env_offset=base, density=density)
    env.set_autoreset(cap, seed=seed, env_offset=base, stride=stride, density=density)
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    acts = torch.randint(0, 5, (A, B, N), generator=g, device="cuda", dtype=torch.uint8)
    codes = torch.zeros((T, B, N), dtype=torch.uint8, device="cuda")
    if chunk % 2 == 0:
        obs, rew, done, steps = env.rollout(acts, num_steps=T, out_codes=codes)
    else:
        # two calls: the second one starts from flipped heuristic-map buffers and staged flags of the first
        T1 = 11
        parts = [env.rollout(acts, num_steps=T1, out_codes=codes[:T1]),
                 env.rollout(torch.roll(acts, shifts=-(T1 % A), dims=0), num_steps=T - T1, out_codes=codes[T1:])]
        obs, rew, done, steps = (torch.cat([a.clone(), b]) for a, b in zip(*parts))
    _native.lib().mapf_debug_rollout_pregen(1)
    _native.lib().mapf_debug_rollout_tasks(0)
    episode = np.zeros(B, dtype=np.int64)
    fin = np.zeros(B, dtype=bool)
    n_resets = n_done = 0
    for t in range(T):
        o_step, r_step, d_step = (x.clone() for x in twin.step(acts[t % A]))   # wrong for the slots that reset; fixed below
        st = twin.steps.clone()
        if fin.any():
            # undo the step for finished slots is not possible: redo them from a re-generated instance instead
            ids = np.flatnonzero(fin)
            for n in np.unique(episode[ids] + 1):
                mask = np.zeros(B, dtype=np.uint8)
                mask[ids[episode[ids] + 1 == n]] = 1
                twin.reset(mask=mask, seed=seed, env_offset=base + int(n) * stride, density=density)
            episode[ids] += 1
            n_resets += len(ids)
            o_new, _ = twin.observe()
            m = torch.as_tensor(fin, device="cuda")
            o_step[m] = o_new[m]
            r_step[m] = 0
            d_step[m] = 0
            st = twin.steps.clone()
        assert torch.equal(obs[t], o_step), (t, (obs[t] != o_step).flatten(1).any(1).nonzero().flatten().tolist()[:8])
        assert torch.equal(rew[t], r_step) and torch.equal(done[t], d_step), t
        assert torch.equal(steps[t], st), t
        m = torch.as_tensor(fin, device="cuda")
        assert bool((codes[t][m] == 5).all()) and bool((codes[t][~m] <= 4).all())
        n_done += int(d_step.sum())
        fin = (d_step.cpu().numpy() != 0) | (st.cpu().numpy() >= cap)
    assert n_resets >= B * (T // (cap + 1)) - B, n_resets
    if N <= 2:
        assert n_done > 0, "tiny boards must exercise the done-triggered reset"
    assert torch.equal(env.agents_pos, twin.agents_pos) and torch.equal(env.goals_pos, twin.goals_pos)
    assert torch.equal(env.map, twin.map) and torch.equal(env.navi_map, twin.navi_map)
    env.check()
    twin.check()
    # switched off again: plain stepping, finished environments keep stepping like the reference's
    env.set_autoreset(0)
    o1 = env.rollout(acts, num_steps=2)[0]
    twin.step(acts[0])
    o2, _, _ = twin.step(acts[1])
    assert torch.equal(o1[1], o2)
    # ... and through the single-step kernels (they read the live heuristic-map buffer too)
    assert torch.equal(env.step(acts[2])[0], twin.step(acts[2])[0])
    assert torch.equal(env.observe()[0], twin.observe()[0])


def test_rollout_bad_arguments():
    import ctypes as C
    import torch
    from mapf_rl_b200 import _native
    env = make_env(8, 4, 10)
    env.reset(seed=0, density=0.2)
    acts = torch.zeros((2, 8, 4), dtype=torch.uint8, device="cuda")
    obs = torch.empty((2, 8, 4, 6, 9, 9), dtype=torch.uint8, device="cuda")
    rew = torch.empty((2, 8, 4), dtype=torch.float32, device="cuda")
    done = torch.empty((2, 8), dtype=torch.uint8, device="cuda")
    lib, vp = _native.lib(), C.c_void_p
    ok = (vp(acts.data_ptr()), 2, vp(obs.data_ptr()), 2, vp(rew.data_ptr()), vp(done.data_ptr()), None, 2)
    assert lib.mapf_env_rollout(env._h, 0, *ok, 0, None) == 0          # T = 0: nothing to do
    assert lib.mapf_env_rollout(env._h, 2, *ok, 17, None) == _native.MAPF_EINVAL
    assert lib.mapf_env_rollout(env._h, -1, *ok, 0, None) == _native.MAPF_EINVAL
    assert lib.mapf_env_rollout(env._h, 2, None, 2, *ok[2:], 0, None) == _native.MAPF_EINVAL
    assert lib.mapf_env_rollout(env._h, 2, ok[0], 0, *ok[2:], 0, None) == _native.MAPF_EINVAL
    # an out-of-range action latches like in step (environment.py:289-290)
    acts[1, 3, 2] = 7
    env.rollout(acts, chains=2)
    with pytest.raises(AssertionError):
        env.check()


@pytest.mark.gpu
@pytest.mark.parametrize("pregen", [2, 3, 0])
def test_rollout_autoreset_graph_replay(rollout_tuning, pregen):
    """A rollout with episode handling captured into a CUDA graph (due list, pre-generation kernels and the rollout kernel on the
    capturing stream; the 'beside' form falls back to in-place re-generation while capturing) and replayed: every replay equals
    an eager twin with the same episode settings."""
    import torch
    from mapf_rl_b200 import _native
    B, N, L, T, cap = 192, 8, 12, 7, 5
    rollout_tuning(1, 0, 0, 0)
    _native.lib().mapf_debug_rollout_pregen(pregen)
    try:
        env, twin = make_env(B, N, L), make_env(B, N, L)
        for e in (env, twin):
            e.reset(seed=21, env_offset=0, density=0.2)
            e.set_autoreset(cap, seed=21, env_offset=B, stride=B, density=0.2)
        g = torch.Generator(device="cuda")
        g.manual_seed(5)
        acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
        rings = [torch.zeros(sh, dtype=dt, device="cuda") for sh, dt in
                 (((T, B, N, 6, 9, 9), torch.uint8), ((T, B, N), torch.float32), ((T, B), torch.uint8), ((T, B), torch.int32))]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside the capture (lazy allocations, module loading)
            env.rollout(acts, out_obs=rings[0], out_rewards=rings[1], out_done=rings[2], out_steps=rings[3])
        side.synchronize()
        twin.rollout(acts)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            env.rollout(acts, out_obs=rings[0], out_rewards=rings[1], out_done=rings[2], out_steps=rings[3])
        for it in range(4):
            acts.copy_(torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8))
            graph.replay()
            o, r, d, s = twin.rollout(acts)
            assert torch.equal(rings[0], o) and torch.equal(rings[1], r) and torch.equal(rings[2], d) and torch.equal(rings[3], s), it
            assert torch.equal(env.agents_pos, twin.agents_pos) and torch.equal(env.goals_pos, twin.goals_pos)
            assert torch.equal(env.navi_map, twin.navi_map)
        assert int(env.episode_counts().sum()) > B      # episodes did end inside the replays
        env.check()
        twin.check()
    finally:
        _native.lib().mapf_debug_rollout_pregen(1)
