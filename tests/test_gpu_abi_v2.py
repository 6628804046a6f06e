"""GPU parity of the ABI-v2 additions: reward codes, the host-buffer step that publishes its results zero-copy
(mapf_env_step_host_codes), input validation at load / set_state, the device-side 'unique' check, the fused learner
cycle (mapf_per_cycle), sum-tree index bounds and the n-step actor TD — each against the C oracle / the existing
single-purpose entry points."""
import ctypes as C

import numpy as np
import pytest

from helpers import instances, random_instance
from oracle import oracle

pytestmark = pytest.mark.gpu


def make_env(B, N, L):
    from mapf_rl_b200 import BatchedEnvironment
    return BatchedEnvironment(B, N, L)


def test_step_reward_codes_match_rewards():
    import torch
    maps, agents, goals = instances(32)
    B, N = 100, 32
    env = make_env(B, N, 40)
    env.load(maps[:B], agents[:B], goals[:B])
    rng = np.random.default_rng(3)
    table = torch.as_tensor(env.reward_table, device="cuda")
    codes = torch.full((B, N), 255, dtype=torch.uint8, device="cuda")
    seen = set()
    for s in range(12):
        acts = torch.as_tensor(rng.integers(0, 5, size=(B, N)).astype(np.uint8)).cuda()
        _, rew, _ = env.step(acts, out_codes=codes)
        assert int(codes.max()) <= 4
        assert torch.equal(table[codes.long()], rew)
        seen |= set(codes.unique().tolist())
    assert {0, 2, 3} <= seen     # move, stay off goal, collision all occur under uniform actions
    env.check()


@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_codes_vs_oracle(pinned):
    """mapf_env_step_host_codes straight through the C ABI: page-locked and pageable caller buffers, several consecutive
    steps over rotating observation slots; codes / done / steps are final when the call returns, the observation is
    stream-ordered."""
    import torch
    from mapf_rl_b200 import _native
    lib = _native.lib()
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
    acts, codes = mk((B, N), torch.uint8), mk((B, N), torch.uint8)
    done, steps = mk((B,), torch.uint8), mk((B,), torch.int32)
    ring = torch.zeros((3, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda:0")
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    table = env.reward_table
    rng = np.random.default_rng(9)
    for s in range(9):
        acts[:] = rng.integers(0, 5, size=(B, N))
        slot = ring[s % 3]
        codes[:] = 255
        _native.check(lib.mapf_env_step_host_codes(env._h, vp(acts), vp(codes), vp(done), vp(steps), C.c_void_p(slot.data_ptr()),
                                                   env._stream()))
        c, d, st = codes.copy(), done.copy(), steps.copy()      # final on return, before any synchronisation
        obs = slot.cpu().numpy()                                # stream-ordered after the kernel
        for k in range(B):
            (oo, _), orw, od, _ = ora[k].step(acts[k])
            assert np.array_equal(oo.astype(np.uint8), obs[k]), (s, k)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), table[c[k]]), (s, k)
            assert int(od) == d[k] and st[k] == s + 1
    env.check()


@pytest.mark.parametrize("L,N", [(80, 64), (72, 20), (120, 128), (20, 7)])
def test_step_host_codes_large_maps_vs_oracle(L, N):
    """The host-buffer step on maps wider than 56 cells (step-only kernel with the ranked occupant lookup + observe kernel
    on the position snapshot), dense enough for swaps, vertex conflicts and chains; every step against the oracle."""
    import torch
    from mapf_rl_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(L + N)
    B = 9
    insts = [random_instance(rng, L, N, 0.2) for _ in range(B)]
    maps, agents, goals = (np.stack([i[j] for i in insts]) for j in range(3))
    # crowd the agents of half the environments into one corner so that conflicts are frequent
    for k in range(0, B, 2):
        free = np.argwhere(maps[k][:12, :12] == 0)
        if len(free) >= N:
            agents[k] = free[rng.permutation(len(free))[:N]]
    env = make_env(B, N, L)
    env.load(maps, agents, goals)
    ora = []
    for k in range(B):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        ora.append(o)
    mk = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=True).numpy()
    acts, codes = mk((B, N), torch.uint8), mk((B, N), torch.uint8)
    done, steps = mk((B,), torch.uint8), mk((B,), torch.int32)
    ring = torch.zeros((2, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda:0")
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    table = env.reward_table
    collisions = 0
    for s in range(12):
        acts[:] = rng.integers(0, 5, size=(B, N))
        slot = ring[s % 2]
        _native.check(lib.mapf_env_step_host_codes(env._h, vp(acts), vp(codes), vp(done), vp(steps), C.c_void_p(slot.data_ptr()),
                                                   env._stream()))
        c, d, st = codes.copy(), done.copy(), steps.copy()
        obs = slot.cpu().numpy()
        for k in range(B):
            (oo, _), orw, od, _ = ora[k].step(acts[k])
            assert np.array_equal(oo.astype(np.uint8), obs[k]), (s, k)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), table[c[k]]), (s, k)
            assert int(od) == d[k] and st[k] == s + 1
        collisions += int((c == 3).sum())
    assert collisions > 0
    env.check()


def test_step_host_codes_full_size_back_to_back():
    """Many back-to-back calls at BASELINE configs[1] size over rotating page-locked action buffers: every call's host
    results are those of THAT step (a twin stepped on the device), although the call returns before its kernel ends."""
    import torch
    B, N, L, T = 8192, 32, 40, 60
    env, twin = make_env(B, N, L), make_env(B, N, L)
    for e in (env, twin):
        e.reset(seed=17, density=0.3)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.uint8)
    acts_host = acts.cpu().pin_memory()
    table = torch.as_tensor(env.reward_table, device="cuda")
    exp = []
    codes_dev = torch.empty((B, N), dtype=torch.uint8, device="cuda")
    for t in range(T):
        o, r, d = twin.step(acts[t], out_codes=codes_dev)
        exp.append((codes_dev.cpu().numpy().copy(), d.cpu().numpy().copy()))
    ring = torch.zeros((2, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
    for t in range(T):
        codes, done, steps = env.step_host_codes(acts_host[t], device_obs=ring[t % 2])
        assert np.array_equal(codes, exp[t][0]), t
        assert np.array_equal(done, exp[t][1]), t
        assert int(steps.min()) == t + 1 and int(steps.max()) == t + 1
    assert torch.equal(ring[(T - 1) % 2], o)
    assert torch.equal(env.agents_pos, twin.agents_pos)
    env.check()


def test_load_validation():
    """Coordinates outside the map raise IndexError (host check in the Python mirror, device check behind the C ABI); two
    agents on one cell raise RuntimeError('unique') (environment.py:424-428); a slot id outside the batch IndexError."""
    import torch
    from mapf_rl_b200 import _native
    rng = np.random.default_rng(0)
    L, N, B = 12, 5, 6
    insts = [random_instance(rng, L, N, 0.2) for _ in range(B)]
    maps, agents, goals = (np.stack([i[j] for i in insts]) for j in range(3))
    env = make_env(B, N, L)
    bad = agents.copy()
    bad[2, 3, 1] = L
    with pytest.raises(IndexError):
        env.load(maps, bad, goals)
    bad = agents.astype(np.int64)
    bad[0, 0, 0] = 256 + 3          # would wrap to a valid coordinate in a uint8 cast
    with pytest.raises(IndexError):
        env.load(maps, bad, goals)
    neg = goals.astype(np.int64)
    neg[1, 1, 0] = -1
    with pytest.raises(IndexError):
        env.load(maps, agents, neg)
    # behind the C ABI (no Python range check): the kernel clamps and latches
    lib = _native.lib()
    m = torch.as_tensor(maps).cuda()
    a = torch.as_tensor(agents.astype(np.uint8)).cuda()
    g = torch.as_tensor(goals.astype(np.uint8)).cuda()
    a_bad = a.clone()
    a_bad[4, 2, 0] = 200
    vp = lambda t: C.c_void_p(t.data_ptr())
    assert lib.mapf_env_load(env._h, None, B, vp(m), vp(a_bad), vp(g), env._stream()) == 0
    with pytest.raises(IndexError):
        env.check()
    acts = torch.zeros((B, N), dtype=torch.uint8, device="cuda")
    env.step(acts)                   # the clamped state is steppable: no fault
    torch.cuda.synchronize()
    ids = torch.as_tensor([0, 1, 2, 3, 4, B], dtype=torch.int32).cuda()
    assert lib.mapf_env_load(env._h, vp(ids), B, vp(m), vp(a), vp(g), env._stream()) == 0
    with pytest.raises(IndexError):
        env.check()
    dup = agents.copy()
    dup[3, 1] = dup[3, 4]
    with pytest.raises(RuntimeError, match="unique"):
        env.load(maps, dup, goals)
    with pytest.raises(RuntimeError, match="unique"):
        env.load(maps, agents, goals)
        env.set_state(agents_pos=dup)
    env.load(maps, agents, goals)    # a clean load afterwards works
    env.step(acts)
    env.check()


@pytest.mark.parametrize("L,N", [(10, 40), (72, 40), (100, 70)])
def test_device_unique_check(L, N):
    """environment.py:424-428 on the device: with set_checks(check_unique=True) a step from a state with two agents on one
    cell latches 'unique'; a correct step never does; the drop-in Environment enables it.  Maps wider than 56 cells take
    the ranked occupant lookup (one bit per cell instead of the byte grid)."""
    import torch
    from mapf_rl_b200 import _native
    rng = np.random.default_rng(2)
    B = 7
    insts = [random_instance(rng, L, N, 0.1) for _ in range(B)]
    maps, agents, goals = (np.stack([i[j] for i in insts]) for j in range(3))
    env = make_env(B, N, L)
    env.load(maps, agents, goals)
    env.set_checks(check_unique=True)
    for s in range(8):
        env.step(rng.integers(0, 5, size=(B, N)).astype(np.uint8))
    env.check()
    # inject a duplicate behind the validation (raw pointer copy into the arena is not exposed: use the C ABI set_state and
    # swallow its latch)
    pos = env.agents_pos.clone()
    pos[5, 33] = pos[5, 2]
    lib = _native.lib()
    assert lib.mapf_env_set_state(env._h, C.c_void_p(pos.data_ptr()), None, env._stream()) == 0
    with pytest.raises(RuntimeError, match="unique"):
        env.check()                  # set_state's own validation
    env.step(np.zeros((B, N), dtype=np.uint8))
    with pytest.raises(RuntimeError, match="unique"):
        env.check()                  # ... and the step kernel's
    env.set_checks(check_unique=False)
    env.step(np.zeros((B, N), dtype=np.uint8))
    env.check()


def _learner_batch(rng, n, capacity):
    import torch
    q = lambda: torch.as_tensor(rng.normal(size=(n, 5)).astype(np.float32)).cuda()
    return dict(q_online=q(), q_target_next=q(),
                action=torch.as_tensor(rng.integers(0, 5, size=n)).cuda(),
                reward=torch.as_tensor(rng.choice([-0.075, -0.5, 0.0, 3.0], size=n).astype(np.float16).astype(np.float32)).cuda(),
                done=torch.as_tensor((rng.random(n) < 0.1).astype(np.float32)).cuda(),
                steps=torch.as_tensor(rng.integers(1, 3, size=n).astype(np.float32)).cuda(),
                idx=torch.as_tensor(rng.integers(0, capacity, size=n)).cuda())


def test_per_cycle_equals_td_update_then_sample():
    """mapf_per_cycle (ONE launch: priorities of the last batch in, next batch out) against mapf_per_td_update followed by
    mapf_per_sample on a twin tree, and against the oracle tree: bit-equal heap, identical samples, weights to 1e-6."""
    import torch
    from mapf_rl_b200 import SumTree
    cap, n, slot_steps = 1 << 14, 192, 256
    rng = np.random.default_rng(4)
    a, b, ref = SumTree(cap, device="cuda:0"), SumTree(cap, device="cuda:0"), oracle.OracleSumTree(cap)
    idx0 = rng.permutation(cap)[:6000].astype(np.int64)
    pr0 = rng.random(6000) + 1e-3
    for t in (a, b):
        t.update_device(idx0.copy(), pr0)
    ref.batch_update(idx0.copy(), pr0)
    old_ptr, ptr = 3, 9
    for rnd in range(6):
        batch = _learner_batch(rng, n, cap)
        if rnd == 2:
            batch["idx"][5] = batch["idx"][100]          # duplicate leaf: the later batch position wins (buffer.py:97)
        u = torch.as_tensor(rng.random(n)).cuda()
        out = a.cycle(update=batch, sample_size=n, uniforms=u, beta=0.4, old_ptr=old_ptr, ptr=ptr, slot_steps=slot_steps)
        td, pr = b.td_update(batch["q_online"], batch["q_target_next"], batch["action"], batch["reward"], batch["done"],
                             batch["steps"], batch["idx"], old_ptr=old_ptr, ptr=ptr, slot_steps=slot_steps)
        bi, bp, bw = b.sample_device(n, u, beta=0.4)
        assert torch.equal(out["td"], td) and torch.equal(out["prio"], pr)
        assert torch.equal(a.tree, b.tree)
        assert torch.equal(out["idx"], bi) and torch.equal(out["sample_prio"], bp) and torch.equal(out["weights"], bw)
        # oracle: stale mask + prio ** alpha in fp64 of the fp32 priority, then batch_update.  The device evaluates the power
        # with CUDA's pow (not correctly rounded: <= 2 ulp from numpy's), so leaves agree to 1e-15 relative, not bitwise;
        # the ancestors are then checked as exact sums of the DEVICE's own leaves (root == sum over levels, buffer.py:105)
        ix = batch["idx"].cpu().numpy()
        keep = (ix < old_ptr * slot_steps) | (ix >= ptr * slot_steps)
        leaf = np.power(pr.cpu().numpy().astype(np.float64), 0.6)
        ref.batch_update(ix[keep].copy(), leaf[keep])
        tree = a.tree.cpu().numpy()
        assert np.allclose(ref.tree, tree, rtol=1e-14, atol=0)
        chk = oracle.OracleSumTree(cap)
        leaves = np.arange(cap, dtype=np.int64)
        chk.batch_update(leaves.copy(), tree[cap - 1:].copy())
        assert np.array_equal(chk.tree, tree)                      # every ancestor == left + right, level by level
        ri, rp = chk.batch_sample(n, u.cpu().numpy())
        assert np.array_equal(ri, out["idx"].cpu().numpy()) and np.array_equal(rp, out["sample_prio"].cpu().numpy())
        w = (rp / rp.min()) ** -0.4
        assert np.allclose(out["weights"].cpu().numpy(), w, rtol=1e-6)
    # either half alone
    only_s = a.cycle(sample_size=32, uniforms=torch.as_tensor(rng.random(32)).cuda())
    assert only_s["td"] is None and only_s["idx"].numel() == 32 and only_s["weights"] is None
    only_u = a.cycle(update=_learner_batch(rng, 64, cap))
    assert only_u["idx"] is None and only_u["td"].numel() == 64
    a.check()


def test_sumtree_index_bounds():
    """A leaf index outside [0, capacity): numpy raises IndexError; the kernels skip the entry (no out-of-bounds write) and
    latch MAPF_EINDEX, the numpy-API mirror raises before launching."""
    import torch
    from mapf_rl_b200 import SumTree
    cap = 1 << 10
    t, ref = SumTree(cap, device="cuda:0"), oracle.OracleSumTree(cap)
    idx = np.asarray([5, 17, cap, 900, -3, 17], dtype=np.int64)
    pr = np.asarray([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    t.update_device(idx, pr)
    with pytest.raises(IndexError):
        t.check()
    good = (idx >= 0) & (idx < cap)
    ref.batch_update(idx[good].copy(), pr[good])
    assert np.array_equal(t.tree.cpu().numpy(), ref.tree)
    t.check()
    with pytest.raises(IndexError):
        t.batch_update(idx.copy(), pr)
    big = torch.as_tensor(np.r_[np.arange(5000) % cap, cap + 7].astype(np.int64)).cuda()   # multi-CTA path
    t.update_device(big, torch.ones(5001, dtype=torch.float64, device="cuda"))
    with pytest.raises(IndexError):
        t.check()


@pytest.mark.parametrize("n,gamma", [(2, 0.99), (3, 0.99), (5, 0.9), (1, 0.99)])
def test_actor_td_n_step(n, gamma):
    """LocalBuffer.finish TD for config.forward_steps = n (buffer.py:174-175) against the oracle's restatement; n = 2,
    gamma = 0.99 is the reference's configuration (pinned to the live reference in tests/golden/per.npz)."""
    import torch
    from mapf_rl_b200.buffer import actor_td_errors
    rng = np.random.default_rng(n)
    E, cap = 9, 64
    size = rng.integers(1, cap + 1, size=E).astype(np.int32)
    size[0], size[1] = cap, 1
    rew = rng.choice([-0.075, -0.5, 0.0, 3.0], size=(E, cap)).astype(np.float16).astype(np.float32)
    q = rng.normal(size=(E, cap, 5)).astype(np.float32)
    act = rng.integers(0, 5, size=(E, cap)).astype(np.uint8)
    td = actor_td_errors(rew, q, act, size, capacity=cap, forward_steps=n, gamma=gamma).cpu().numpy()
    for e in range(E):
        exp = oracle.actor_td_n(rew[e, :size[e]].astype(np.float64), q[e, :size[e]], act[e, :size[e]], cap, n, gamma)
        assert np.array_equal(td[e], exp), e
        if n == 2 and gamma == 0.99:
            assert np.array_equal(exp, oracle.actor_td(rew[e, :size[e]].astype(np.float64), q[e, :size[e]], act[e, :size[e]], cap))
