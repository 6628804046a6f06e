"""GPU: the drop-in `Environment` class keeps the reference's API surface and types
(environment.py:74-508) and survives ray-free restatements of its two callers."""
import numpy as np
import pytest

from helpers import instances
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_signature_and_types():
    from mapf_rl_b200 import Environment, config
    env = Environment()
    assert env.num_agents == config.num_agents and env.map_size == (config.map_length, config.map_length)
    assert env.map.shape == (20, 20) and env.agents_pos.shape == (6, 2) and env.goals_pos.dtype == np.int64
    assert env.steps == 0 and env.obs_radius == 4 and env.reward_fn == config.reward_fn
    # a fresh instance: starts and goals are 12 distinct cells (environment.py:118-137 removes every drawn cell from its
    # partition); checked BEFORE the step -- the instance is unseeded, and a step may put an agent on a goal
    assert len({tuple(p) for p in env.agents_pos} | {tuple(g) for g in env.goals_pos}) == 12
    obs, pos = env.observe()
    assert obs.shape == (6, 6, 9, 9) and obs.dtype == np.bool_ and pos.dtype == np.int64
    (obs, pos), rewards, done, info = env.step([0, 1, 2, 3, 4, 0])
    assert isinstance(rewards, list) and len(rewards) == 6 and isinstance(done, bool) and info == {'step': 0}
    assert env.steps == 1
    assert env.navi_map.shape == (6, 4, 28, 28) and env.navi_map.dtype == np.bool_
    with pytest.raises(AssertionError):
        env.step([0, 1, 2, 3, 4, 5])
    with pytest.raises(AssertionError):
        env.step([0, 1])


def test_adaptive_and_reset():
    from mapf_rl_b200 import Environment
    env = Environment(adaptive=True)
    assert env.num_agents == 1 and env.map_size == (10, 10)
    obs, pos = env.reset([(2, 10), (3, 15)])
    assert env.num_agents in (2, 3) and obs.shape[0] == env.num_agents and env.map.dtype == np.float32
    env2 = Environment(num_agents=4, map_length=12)
    obs, pos = env2.reset(num_agents=8, map_length=16)
    assert obs.shape == (8, 6, 9, 9) and env2.map_size == (16, 16) and env2.steps == 0


def test_load_and_eval_loop_matches_oracle():
    """test.py:105-131 inner loop with a fixed pseudo-policy, on pkl instances, vs the oracle."""
    from mapf_rl_b200 import Environment
    maps, agents, goals = instances(16)
    env = Environment()
    rng = np.random.default_rng(5)
    for k in (3, 77):
        env.load(maps[k].astype(np.float32), agents[k].astype(np.int64), goals[k].astype(np.int64))
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        assert np.array_equal(env.navi_map[:, :, 4:-4, 4:-4], o.navi_map.astype(bool))
        done = False
        while not done and env.steps < 40:
            obs_pos = env.observe()
            actions = rng.integers(0, 5, size=env.num_agents).tolist()
            (obs, pos), r, done, info = env.step(actions)
            (oobs, opos), orr, odone, oinfo = o.step(actions)
            assert np.array_equal(obs, oobs) and np.array_equal(pos, opos) and r == orr and done == odone and info == oinfo
            assert [type(x) for x in r] == [type(x) for x in orr]
        assert env.steps == 40 or done


def test_actor_loop_restatement():
    """worker.py:368-414 without ray / the network: env.step -> LocalBuffer.add -> finish at max_steps."""
    from mapf_rl_b200 import Environment, LocalBuffer, SumTree, config
    env = Environment(num_agents=3, map_length=10)
    obs_pos = env.reset(num_agents=3, map_length=10)
    lb = LocalBuffer(0, env.num_agents, env.map_size[0], obs_pos[0], size=32)
    rng = np.random.default_rng(0)
    done = False
    while not done and env.steps < 32:
        q = rng.normal(size=(env.num_agents, 5)).astype(np.float32)
        actions = q.argmax(1).tolist()
        next_obs_pos, r, done, _ = env.step(actions)
        lb.add(q[0], actions[0], r[0], next_obs_pos[0], np.zeros((env.num_agents, config.latent_dim)), np.zeros((3, 3)))
    res = lb.finish(None if done else q[0], None if done else np.zeros((3, 3)))
    assert res[9] == lb.size and res[7].shape == (32,)
    tree = SumTree(64)
    tree.batch_update(np.arange(0, 32), res[7] ** config.prioritized_replay_alpha)  # worker.py:94
    assert tree.sum() > 0
