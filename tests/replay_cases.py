"""Shared builder of replay episodes for the ReplayStore / OracleReplay / live GlobalBuffer tests."""
import numpy as np


def make_episode(rng, actor_id, num_agents, size, done, latent=256, cap=256):
    """A tuple shaped like LocalBuffer.finish's return value (buffer.py:179; layout worker.py:72)."""
    obs = rng.random((size + 1, num_agents, 6, 9, 9)) < 0.2
    act = rng.integers(0, 5, size=size).astype(np.uint8)
    rew = rng.choice([-0.075, -0.5, 0.0, 3.0], size=size).astype(np.float16)
    hid = rng.standard_normal((size, num_agents, latent)).astype(np.float16)
    td = np.zeros(cap, dtype=np.float64)
    td[:size] = rng.random(size) * 2 + 1e-3
    comm = rng.random((size + 1, num_agents, num_agents)) < 0.3
    return (actor_id, num_agents, 10, obs, act, rew, hid, td, bool(done), size, comm)


# (num_agents, size, done) of the episodes added in order; capacity 4 => the 5th and 6th overwrite slots 0 and 1
EPISODES = [(6, 5, True), (3, 16, False), (6, 17, True), (2, 40, False), (4, 256, False), (6, 1, True)]
CAPACITY = 4
BATCH = 48
