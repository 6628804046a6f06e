"""Drop-in replacement of the reference's `environment.Environment` (environment.py:74-508).

Same constructor, `reset`, `load`, `step`, `observe`, attributes and return types, so the reference's
actors (worker.py:355-428) and evaluation loop (test.py:82-145) run unchanged against it.  All compute
(BFS heuristic maps, conflict resolution, rewards, observation) happens in the CUDA kernels of
libmapf_b200.so on a batch of one; this class only converts between the reference's numpy types and
device tensors.  It exists for compatibility — throughput comes from `BatchedEnvironment`.
"""
from __future__ import annotations

import random
from typing import List

import numpy as np

from . import config
from .batched import BatchedEnvironment
from .instances import generate_instance

# environment.py:12
action_list = np.array([[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1]], dtype=int)


class Environment:
    def __init__(self, adaptive=False, map_length: int = config.map_length, num_agents: int = config.num_agents,
                 obs_radius: int = config.obs_radius, reward_fn: dict = config.reward_fn, *,
                 obstacle_density=None, device=None, seed=None):
        # environment.py:91-97
        self.adaptive = adaptive
        if adaptive:
            self.num_agents = config.init_set[0]
            self.map_size = (config.init_set[1], config.init_set[1])
        else:
            self.num_agents = num_agents
            self.map_size = (map_length, map_length)
        self.obs_radius = obs_radius
        self.reward_fn = reward_fn
        self._device = device
        self._fixed_density = obstacle_density
        self._rng = np.random.default_rng(seed if seed is not None else random.getrandbits(63))
        self._envs = {}   # (num_agents, map_length) -> BatchedEnvironment(1, ...)
        self._env = None
        self.imgs = []
        self._generate()  # environment.py:100-143 builds an instance immediately
        self.steps = 0

    # -- internals -----------------------------------------------------------------------------
    def _backend(self) -> BatchedEnvironment:
        key = (int(self.num_agents), int(self.map_size[0]))
        env = self._envs.get(key)
        if env is None:
            env = BatchedEnvironment(1, key[0], key[1], device=self._device, obs_radius=self.obs_radius,
                                     reward_fn=self.reward_fn)
            env.set_checks(check_unique=True)   # environment.py:424-428, on the device
            self._envs[key] = env
        self._env = env
        return env

    def _generate(self):
        m, a, g, self._last_density = generate_instance(self._rng, self.map_size[0], self.num_agents, self._fixed_density,
                                                        return_density=True)
        self.obstacle_density = float(self._last_density)   # the SAMPLED density, as environment.py:100,156 store it
        self._install(m, a, g)

    def _install(self, m, agents, goals):
        self.map = np.copy(m)
        self.agents_pos = np.asarray(agents, dtype=np.int64).copy()
        self.goals_pos = np.asarray(goals, dtype=np.int64).copy()
        env = self._backend()
        env.load(np.asarray(m)[None], self.agents_pos[None], self.goals_pos[None])   # range-checked, not wrapped
        self._navi_cache = None

    # -- reference API -------------------------------------------------------------------------
    def reset(self, level=None, num_agents=None, map_length=None):  # environment.py:146-196
        if self.adaptive:
            rand = random.choice(level)
            self.num_agents = rand[0]
            self.map_size = (rand[1], rand[1])
        elif num_agents is not None:
            self.num_agents = num_agents
            self.map_size = (map_length, map_length)
        self._generate()
        self.map = self.map.astype(np.float32)  # environment.py:157
        self.steps = 0
        return self.observe()

    def load(self, map: np.ndarray, agents_pos: np.ndarray, goals_pos: np.ndarray):  # environment.py:198-215
        self.num_agents = agents_pos.shape[0]
        self.map_size = (map.shape[0], map.shape[1])
        assert self.map_size[0] == self.map_size[1], "square maps only (environment.py:322)"
        self._install(map, agents_pos, goals_pos)
        self.steps = 0
        self.imgs = []

    @property
    def navi_map(self) -> np.ndarray:
        """bool[N,4,L+2r,L+2r] exactly as environment.py:253-276 leaves it."""
        if self._navi_cache is None:
            nv = self._env.navi_map[0].cpu().numpy().astype(bool)
            r = self.obs_radius
            self._navi_cache = np.pad(nv, ((0, 0), (0, 0), (r, r), (r, r)))
        return self._navi_cache

    def get_navi_map(self):  # environment.py:217-276 — recompute on the device
        import ctypes as C
        from . import _native
        env = self._env
        _native.check(env._lib.mapf_env_bfs_navi(env._h, None, 1, None, env._stream()))
        self._navi_cache = None

    def step(self, actions: List[int]):  # environment.py:278-430
        assert len(actions) == self.num_agents, 'actions number' + str(actions)
        assert all([action_idx < 5 and action_idx >= 0 for action_idx in actions]), 'action index out of range'
        a = np.asarray(actions, dtype=np.uint8).reshape(1, self.num_agents)
        obs, rewards, done, _ = self._env.step_host(a, want_obs=True)
        self.steps += 1
        self.agents_pos = self._env.agents_pos[0].cpu().numpy().astype(np.int64)
        done = bool(done[0])
        # the reference returns python numbers: ints for stay_on_goal (0) and finish (3), floats otherwise
        table = {np.float32(v): v for v in self.reward_fn.values()}
        rewards = [table.get(r, float(r)) for r in rewards[0]]
        info = {'step': self.steps - 1}
        self._env.check()   # environment.py:424-428: RuntimeError('unique') latched by the step kernel
        return (obs[0].astype(bool), self.agents_pos), rewards, done, info

    def observe(self):  # environment.py:433-467
        obs, pos = self._env.observe()
        self.agents_pos = pos[0].cpu().numpy().astype(np.int64)
        return obs[0].cpu().numpy().astype(bool), self.agents_pos

    def render(self):  # environment.py:469-500: matplotlib visualisation, out of scope
        pass

    def close(self, save=False):  # environment.py:502-508
        pass
