"""Replay store + sampled-window gather (GlobalBuffer.add / sample_batch / update_priorities, worker.py:68-203).

CPU (dev container): OracleReplay against the LIVE reference GlobalBuffer (worker.py imported with a no-op ray
stub), same episodes, same uniforms -> every element of the sampled tuple identical.
GPU: ReplayStore (CUDA gather) against OracleReplay, bit for bit."""
import numpy as np
import pytest

from oracle import oracle, ref_loader
from replay_cases import BATCH, CAPACITY, EPISODES, make_episode


def drive(store, sample, rounds=3):
    """Common scenario: add 4 episodes, sample, add 2 more (wrap: slots 0/1 overwritten), sample, update priorities
    with a stale-slot window, sample again.  Returns the list of sampled tuples."""
    rng = np.random.default_rng(42)
    eps = [make_episode(rng, k, n, size, done) for k, (n, size, done) in enumerate(EPISODES)]
    outs = []
    store.add(eps[:4])
    outs.append(sample(store, 100))
    old_ptr = outs[-1][10]
    store.add(eps[4:])
    outs.append(sample(store, 101))
    # priorities for the first sample arrive late: slots 0 and 1 were overwritten meanwhile (worker.py:192-201)
    idx = np.array(outs[0][8], dtype=np.int64)
    pr = (np.random.default_rng(5).random(idx.shape[0]) + 0.01).astype(np.float16)
    store.update_priorities(idx, pr, old_ptr)
    outs.append(sample(store, 102))
    return outs


def to_np(x):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(x)


def assert_same(a, b):
    names = ["obs", "action", "reward", "done", "steps", "bt_steps", "hidden", "comm_mask", "idxes", "weights", "ptr"]
    for k, name in enumerate(names):
        x, y = to_np(a[k]), to_np(b[k])
        assert x.shape == y.shape, (name, x.shape, y.shape)
        assert np.array_equal(x.astype(np.float64), y.astype(np.float64)), name


def oracle_sample(store, seed):
    return store.sample_batch(BATCH, np.random.RandomState(seed).random_sample(BATCH))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_oracle_replay_vs_live_global_buffer():
    worker = ref_loader.load_worker()

    def ref_sample(store, seed):
        np.random.seed(seed)                      # SumTree.batch_sample draws from the global stream (buffer.py:60)
        return store.sample_batch(BATCH)

    ref_outs = drive(worker.GlobalBuffer(CAPACITY), ref_sample)
    ora = oracle.OracleReplay(CAPACITY)
    ora_outs = drive(ora, oracle_sample)
    for a, b in zip(ref_outs, ora_outs):
        assert_same(a, b)
    assert len({int(t) for o in ora_outs for t in o[5]}) > 5          # short and full burn-in windows both hit


def test_oracle_replay_golden():
    """Same scenario against the vectors recorded from the live reference (tests/golden/replay.npz)."""
    from helpers import golden
    z = golden("replay.npz")
    outs = drive(oracle.OracleReplay(CAPACITY), oracle_sample)
    for r, o in enumerate(outs):
        assert np.array_equal(np.packbits(o[0].astype(bool).reshape(-1)), z[f"obs_packed_{r}"])
        assert np.array_equal(o[1], z[f"action_{r}"]) and np.array_equal(o[2], z[f"reward_{r}"])
        assert np.array_equal(o[3], z[f"done_{r}"]) and np.array_equal(o[4], z[f"steps_{r}"])
        assert np.array_equal(o[5], z[f"bt_steps_{r}"]) and np.array_equal(o[6], z[f"hidden_{r}"])
        assert np.array_equal(np.packbits(o[7].reshape(-1)), z[f"comm_packed_{r}"])
        assert np.array_equal(o[8], z[f"idxes_{r}"]) and np.array_equal(o[9], z[f"weights_{r}"]) and o[10] == z[f"ptr_{r}"]


@pytest.mark.gpu
def test_replay_store_vs_oracle():
    from mapf_rl_b200 import ReplayStore

    def gpu_sample(store, seed):
        return store.sample_batch(BATCH, np.random.RandomState(seed).random_sample(BATCH))

    gpu_outs = drive(ReplayStore(CAPACITY, device="cuda:0"), gpu_sample)
    ora_outs = drive(oracle.OracleReplay(CAPACITY), oracle_sample)
    for a, b in zip(gpu_outs, ora_outs):
        assert_same(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("n,cap", [(8, 4), (32, 2), (1, 8), (5, 4)])
def test_replay_gather_agent_counts(n, cap):
    """Both conversion paths (16-byte when N % 8 == 0, 4-byte otherwise), random episodes, reference batch size 192."""
    from mapf_rl_b200 import ReplayStore
    rng = np.random.default_rng(n * 10 + cap)
    eps = [make_episode(rng, k, int(rng.integers(1, n + 1)), int(rng.integers(1, 257)), bool(rng.integers(0, 2)))
           for k in range(cap + 1)]
    gpu, ora = ReplayStore(cap, max_num_agents=n, device="cuda:0"), oracle.OracleReplay(cap, max_num_agents=n)
    gpu.add(eps), ora.add(eps)
    u = rng.random(192)
    assert_same(gpu.sample_batch(192, u), ora.sample_batch(192, u))
    assert gpu.size == ora.size and gpu.ptr == ora.ptr


@pytest.mark.gpu
def test_step_kernel_writes_into_store_rows():
    """Observations land in the replay store straight from the step kernel (no copy)."""
    import torch
    from helpers import instances
    from mapf_rl_b200 import BatchedEnvironment, ReplayStore
    maps, agents, goals = instances(16)
    store = ReplayStore(2, max_num_agents=16, device="cuda:0")
    env = BatchedEnvironment(1, 16, 40, device="cuda:0")
    env.load(maps[:1], agents[:1], goals[:1])
    o = oracle.OracleEnv()
    o.load(maps[0], agents[0], goals[0])
    env.observe(out_obs=store.obs_rows(1, 0, 1))
    rng = np.random.default_rng(3)
    want = [o.observe()[0]]
    for t in range(5):
        a = rng.integers(0, 5, size=(1, 16)).astype(np.uint8)
        env.step(a, out_obs=store.obs_rows(1, t + 1, 1))
        want.append(o.step(a[0])[0][0])
    got = store.obs_buf[257:257 + 6].cpu().numpy()
    assert np.array_equal(got, np.stack(want).astype(np.uint8))
    assert int(store.obs_buf[:257].sum().item()) == 0
