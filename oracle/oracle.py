"""TEST INFRASTRUCTURE — ctypes front-end of the C oracle (oracle/mapf_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package `mapf_rl_b200` never does.

`OracleEnv` mirrors the reference `Environment` surface used on the hot path
(environment.py:198-215 load, :278-430 step, :433-467 observe, :217-276 get_navi_map) so parity
tests read like the reference's callers.  `OracleSumTree` mirrors buffer.py:16-105.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmapf_oracle.so")
_lib = None

# config.py:8-12, in the repo-wide order move, stay_on_goal, stay_off_goal, collision, finish
REWARD_FN = dict(move=-0.075, stay_on_goal=0, stay_off_goal=-0.075, collision=-0.5, finish=3)
REWARD_ORDER = ("move", "stay_on_goal", "stay_off_goal", "collision", "finish")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mapf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.mo_step.restype = C.c_int
        L.mo_step.argtypes = [C.c_int, C.c_int, u8p, i32p, i32p, u8p, f64p, f64p]
        L.mo_navi.restype = None
        L.mo_navi.argtypes = [C.c_int, C.c_int, u8p, i32p, i32p, u8p]
        L.mo_observe.restype = None
        L.mo_observe.argtypes = [C.c_int, C.c_int, C.c_int, u8p, i32p, u8p, u8p]
        L.mo_rollout.restype = C.c_long
        L.mo_rollout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, i32p, i32p, u8p, u8p, f64p,
                                 f32p, u8p, u8p]
        L.mo_tree_batch_update.restype = None
        L.mo_tree_batch_update.argtypes = [f64p, C.c_int64, C.c_int, i64p, f64p, C.c_int64]
        L.mo_tree_batch_sample.restype = None
        L.mo_tree_batch_sample.argtypes = [f64p, C.c_int64, C.c_int, f64p, C.c_int64, i64p, f64p]
        L.mo_actor_td.restype = None
        L.mo_actor_td.argtypes = [C.c_int, C.c_int, f64p, f32p, u8p, f64p]
        L.mo_actor_td_n.restype = None
        L.mo_actor_td_n.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, f64p, f32p, u8p, f64p]
        L.mo_learner_td.restype = None
        L.mo_learner_td.argtypes = [C.c_int64, f32p, f32p, i64p, f32p, f32p, f32p, f32p, f32p]
        L.mo_navi_batch.restype = None
        L.mo_navi_batch.argtypes = [C.c_int, C.c_int, C.c_int, u8p, i32p, u8p]
        L.mo_set_threads.restype = C.c_int
        L.mo_set_threads.argtypes = [C.c_int]
        L.mo_comm_mask.restype = None
        L.mo_comm_mask.argtypes = [C.c_int, i32p, C.c_int, C.c_int, u8p]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def reward_vector(reward_fn=None) -> np.ndarray:
    rf = REWARD_FN if reward_fn is None else reward_fn
    return np.asarray([rf[k] for k in REWARD_ORDER], dtype=np.float64)


def navi(map_: np.ndarray, goals: np.ndarray):
    """-> (dist int32[N,L,L], navi uint8[N,4,L,L] unpadded).  environment.py:217-274."""
    m = np.ascontiguousarray(np.asarray(map_) != 0, dtype=np.uint8)
    L = m.shape[0]
    g = np.ascontiguousarray(goals, dtype=np.int32)
    N = g.shape[0]
    dist = np.empty((N, L, L), dtype=np.int32)
    nv = np.empty((N, 4, L, L), dtype=np.uint8)
    lib().mo_navi(L, N, _p(m, C.c_uint8), _p(g, C.c_int32), _p(dist, C.c_int32), _p(nv, C.c_uint8))
    return dist, nv


def navi_batch(maps: np.ndarray, goals: np.ndarray, threads: int = 0) -> np.ndarray:
    """maps [B,L,L], goals [B,N,2] -> navi uint8[B,N,4,L,L] (mo_navi per environment, OpenMP over the batch)."""
    m = np.ascontiguousarray(np.asarray(maps) != 0, dtype=np.uint8)
    g = np.ascontiguousarray(goals, dtype=np.int32)
    B, L, N = m.shape[0], m.shape[1], g.shape[1]
    nv = np.empty((B, N, 4, L, L), dtype=np.uint8)
    if threads:
        lib().mo_set_threads(int(threads))
    lib().mo_navi_batch(B, L, N, _p(m, C.c_uint8), _p(g, C.c_int32), _p(nv, C.c_uint8))
    return nv


class OracleEnv:
    """Single environment with the reference's load/step/observe contract."""

    def __init__(self, obs_radius: int = 4, reward_fn=None):
        self.obs_radius = obs_radius
        self.reward_fn = dict(REWARD_FN if reward_fn is None else reward_fn)
        self._rf = reward_vector(self.reward_fn)

    def load(self, map_, agents_pos, goals_pos):  # environment.py:198-215
        self.map = np.ascontiguousarray(np.asarray(map_) != 0, dtype=np.uint8)
        self.agents_pos = np.ascontiguousarray(agents_pos, dtype=np.int32).copy()
        self.goals_pos = np.ascontiguousarray(goals_pos, dtype=np.int32).copy()
        self.num_agents = self.agents_pos.shape[0]
        self.map_size = (self.map.shape[0], self.map.shape[1])
        self.steps = 0
        self.dist_map, self.navi_map = navi(self.map, self.goals_pos)

    def step(self, actions):  # environment.py:278-430
        a = np.asarray(actions)
        assert a.shape[0] == self.num_agents, "actions number"
        assert np.all((a >= 0) & (a < 5)), "action index out of range"
        a8 = np.ascontiguousarray(a, dtype=np.uint8)
        rew = np.empty(self.num_agents, dtype=np.float64)
        pos = self.agents_pos.copy()
        d = lib().mo_step(self.map_size[0], self.num_agents, _p(self.map, C.c_uint8), _p(pos, C.c_int32),
                          _p(self.goals_pos, C.c_int32), _p(a8, C.c_uint8), _p(self._rf, C.c_double),
                          _p(rew, C.c_double))
        if d == -2:
            raise RuntimeError("unique")
        assert d >= 0
        self.agents_pos = pos
        self.steps += 1
        return self.observe(), rew.tolist(), bool(d), {"step": self.steps - 1}

    def observe(self):  # environment.py:433-467
        F = 2 * self.obs_radius + 1
        obs = np.empty((self.num_agents, 6, F, F), dtype=np.uint8)
        lib().mo_observe(self.map_size[0], self.num_agents, self.obs_radius, _p(self.map, C.c_uint8),
                         _p(self.agents_pos, C.c_int32), _p(self.navi_map, C.c_uint8), _p(obs, C.c_uint8))
        return obs.astype(bool), self.agents_pos.astype(np.int64)


def rollout(maps, pos, goals, navi_maps, actions, reward_fn=None, obs_radius=4, threads=1,
            want_rewards=True, obs_out=None, done_out=None, rewards_out=None):
    """Lockstep batch: maps u8[B,L,L], pos/goals i32[B,N,2] (pos updated in place), navi u8[B,N,4,L,L],
    actions u8[T,B,N].  Returns (rewards f32[T,B,N] | None, done u8[T,B], obs u8[B,N,6,F,F] of last step).
    `threads` host threads share the environments (OpenMP inside mo_rollout).  Caller-owned `obs_out` /
    `done_out` / `rewards_out` of those shapes are written in place instead of allocating per call (a timed
    loop passes them: np.zeros of the 64 MB observation block costs as much as a step of 4096 envs)."""
    B, L = maps.shape[0], maps.shape[1]
    N = pos.shape[1]
    T = actions.shape[0]
    F = 2 * obs_radius + 1
    rf = reward_vector(reward_fn)
    rewards = None
    if want_rewards:
        rewards = rewards_out if rewards_out is not None else np.zeros((T, B, N), dtype=np.float32)
    done = done_out if done_out is not None else np.zeros((T, B), dtype=np.uint8)
    obs = obs_out if obs_out is not None else np.zeros((B, N, 6, F, F), dtype=np.uint8)
    assert maps.dtype == np.uint8 and pos.dtype == np.int32 and goals.dtype == np.int32
    assert navi_maps.dtype == np.uint8 and actions.dtype == np.uint8
    assert obs.dtype == np.uint8 and obs.shape == (B, N, 6, F, F) and done.dtype == np.uint8 and done.shape == (T, B)
    assert rewards is None or (rewards.dtype == np.float32 and rewards.shape == (T, B, N))
    for a in (maps, pos, goals, navi_maps, actions, obs, done):
        assert a.flags.c_contiguous

    def run(lo, hi):
        nb = hi - lo
        if nb <= 0:
            return
        if lo == 0 and hi == B:   # whole batch: every buffer is used in place
            lib().mo_rollout(B, L, N, obs_radius, T, _p(maps, C.c_uint8), _p(pos, C.c_int32), _p(goals, C.c_int32),
                             _p(navi_maps, C.c_uint8), _p(actions, C.c_uint8), _p(rf, C.c_double),
                             _p(rewards, C.c_float) if want_rewards else None, _p(done, C.c_uint8), _p(obs, C.c_uint8))
            return
        # per-slice contiguous views: actions/rewards/done are [T,B,...] so slice-copy them
        act = np.ascontiguousarray(actions[:, lo:hi])
        rw = np.zeros((T, nb, N), dtype=np.float32) if want_rewards else None
        dn = np.zeros((T, nb), dtype=np.uint8)
        lib().mo_rollout(nb, L, N, obs_radius, T, _p(maps[lo:hi], C.c_uint8), _p(pos[lo:hi], C.c_int32),
                         _p(goals[lo:hi], C.c_int32), _p(navi_maps[lo:hi], C.c_uint8), _p(act, C.c_uint8),
                         _p(rf, C.c_double), _p(rw, C.c_float) if want_rewards else None, _p(dn, C.c_uint8),
                         _p(obs[lo:hi], C.c_uint8))
        if want_rewards:
            rewards[:, lo:hi] = rw
        done[:, lo:hi] = dn

    got = lib().mo_set_threads(max(1, int(threads)))
    if got >= threads or threads <= 1:
        run(0, B)           # one call, OpenMP splits the environments over the host threads
    else:                   # library built without OpenMP: python threads, ctypes releases the GIL
        bounds = np.linspace(0, B, threads + 1).astype(int)
        ts = [threading.Thread(target=run, args=(int(bounds[k]), int(bounds[k + 1]))) for k in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    return rewards, done, obs


class OracleSumTree:
    """buffer.py:16-105 with caller-supplied uniforms for batch_sample."""

    def __init__(self, capacity: int):
        layer = 1
        while 2 ** (layer - 1) < capacity:
            layer += 1
        assert 2 ** (layer - 1) == capacity, "buffer size only support power of 2 size"
        self.layer = layer
        self.capacity = capacity
        self.tree = np.zeros(2 ** layer - 1, dtype=np.float64)

    def batch_sample(self, batch_size: int, uniforms: np.ndarray):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        idx = np.empty(batch_size, dtype=np.int64)
        pr = np.empty(batch_size, dtype=np.float64)
        lib().mo_tree_batch_sample(_p(self.tree, C.c_double), self.capacity, self.layer, _p(u, C.c_double),
                                   batch_size, _p(idx, C.c_int64), _p(pr, C.c_double))
        return idx, pr

    def batch_update(self, idxes: np.ndarray, priorities: np.ndarray):
        assert idxes.dtype == np.int64 and idxes.flags.c_contiguous
        p = np.ascontiguousarray(priorities, dtype=np.float64)
        lib().mo_tree_batch_update(_p(self.tree, C.c_double), self.capacity, self.layer, _p(idxes, C.c_int64),
                                   _p(p, C.c_double), idxes.shape[0])


def actor_td(rew_fp16: np.ndarray, q: np.ndarray, act: np.ndarray, capacity: int = 256) -> np.ndarray:
    """buffer.py:170-177.  rew_fp16: float16[size]; q: float32[size(+1),5]; act: uint8[size]."""
    size = rew_fp16.shape[0]
    r = np.asarray(rew_fp16, dtype=np.float16).astype(np.float64)
    qq = np.ascontiguousarray(q[:size], dtype=np.float32)
    a = np.ascontiguousarray(act, dtype=np.uint8)
    td = np.empty(capacity, dtype=np.float64)
    lib().mo_actor_td(size, capacity, _p(r, C.c_double), _p(qq, C.c_float), _p(a, C.c_uint8), _p(td, C.c_double))
    return td


def actor_td_n(rew_fp16: np.ndarray, q: np.ndarray, act: np.ndarray, capacity: int, forward_steps: int, gamma: float) -> np.ndarray:
    """buffer.py:170-177 for config.forward_steps = n and discount gamma (the reference: n = 2, 0.99)."""
    size = int(len(rew_fp16))
    r = np.ascontiguousarray(rew_fp16, dtype=np.float64)
    qq = np.ascontiguousarray(q, dtype=np.float32)
    a = np.ascontiguousarray(act, dtype=np.uint8)
    td = np.empty(capacity, dtype=np.float64)
    lib().mo_actor_td_n(size, capacity, int(forward_steps), float(gamma), _p(r, C.c_double), _p(qq, C.c_float), _p(a, C.c_uint8),
                        _p(td, C.c_double))
    return td


def learner_td(q_online, q_target_next, action, reward, done, steps):
    """worker.py:300-308 in fp32 -> (td f32[n], priority f32[n])."""
    n = q_online.shape[0]
    qo = np.ascontiguousarray(q_online, dtype=np.float32)
    qt = np.ascontiguousarray(q_target_next, dtype=np.float32)
    a = np.ascontiguousarray(action, dtype=np.int64).reshape(-1)
    r = np.ascontiguousarray(reward, dtype=np.float32).reshape(-1)
    d = np.ascontiguousarray(done, dtype=np.float32).reshape(-1)
    s = np.ascontiguousarray(steps, dtype=np.float32).reshape(-1)
    td = np.empty(n, dtype=np.float32)
    pr = np.empty(n, dtype=np.float32)
    lib().mo_learner_td(n, _p(qo, C.c_float), _p(qt, C.c_float), _p(a, C.c_int64), _p(r, C.c_float),
                        _p(d, C.c_float), _p(s, C.c_float), _p(td, C.c_float), _p(pr, C.c_float))
    return td, pr


def comm_mask(pos: np.ndarray, max_comm_agents: int = 3, obs_radius: int = 4) -> np.ndarray:
    """model.py:196-208 for one environment: pos int[N,2] -> uint8[N,N] (ties -> lower agent id)."""
    p = np.ascontiguousarray(pos, dtype=np.int32)
    N = p.shape[0]
    out = np.empty((N, N), dtype=np.uint8)
    lib().mo_comm_mask(N, _p(p, C.c_int32), min(max_comm_agents, N), obs_radius, _p(out, C.c_uint8))
    return out


class OracleReplay:
    """numpy restatement of GlobalBuffer's storage / sampling (worker.py:21-203), with the sum tree of this
    module and caller-supplied uniforms.  Constants default to config.py:29-30,51,47,65."""

    def __init__(self, capacity, alpha=0.6, beta=0.4, max_num_agents=6, max_steps=256, bt_steps=16, forward_steps=2,
                 latent_dim=256):
        self.capacity, self.alpha, self.beta = capacity, alpha, beta
        self.n, self.S, self.bt, self.fwd, self.latent = max_num_agents, max_steps, bt_steps, forward_steps, latent_dim
        self.size = 0
        self.ptr = 0
        self.priority_tree = OracleSumTree(capacity * max_steps)                                   # :27
        n, S = self.n, self.S
        self.obs_buf = np.zeros(((S + 1) * capacity, n, 6, 9, 9), dtype=bool)                      # :36
        self.act_buf = np.zeros((S * capacity), dtype=np.uint8)
        self.rew_buf = np.zeros((S * capacity), dtype=np.float16)
        self.hid_buf = np.zeros((S * capacity, n, latent_dim), dtype=np.float16)
        self.done_buf = np.zeros(capacity, dtype=bool)
        self.size_buf = np.zeros(capacity, dtype=np.uint64)
        self.comm_mask = np.zeros(((S + 1) * capacity, n, n), dtype=bool)                          # :42

    def add(self, buffer_list):  # :68-104
        S = self.S
        for buffer in buffer_list:
            idxes = np.arange(self.ptr * S, (self.ptr + 1) * S, dtype=np.int64)
            start_idx = self.ptr * S
            self.size -= int(self.size_buf[self.ptr])
            self.size += buffer[9]
            self.priority_tree.batch_update(idxes, np.asarray(buffer[7], dtype=np.float64) ** self.alpha)
            self.obs_buf[start_idx + self.ptr:start_idx + self.ptr + buffer[9] + 1, :buffer[1]] = buffer[3]
            self.act_buf[start_idx:start_idx + buffer[9]] = buffer[4]
            self.rew_buf[start_idx:start_idx + buffer[9]] = buffer[5]
            self.hid_buf[start_idx:start_idx + buffer[9], :buffer[1]] = buffer[6]
            self.done_buf[self.ptr] = buffer[8]
            self.size_buf[self.ptr] = buffer[9]
            self.comm_mask[start_idx + self.ptr:start_idx + self.ptr + buffer[9] + 1, :buffer[1], :buffer[1]] = buffer[10]
            self.ptr = (self.ptr + 1) % self.capacity

    def sample_batch(self, batch_size, uniforms):  # :106-184
        S, bt, fwd = self.S, self.bt, self.fwd
        b_obs, b_action, b_reward, b_done, b_steps, b_bt_steps, b_comm_mask, b_hidden = [], [], [], [], [], [], [], []
        idxes, priorities = self.priority_tree.batch_sample(batch_size, uniforms)
        global_idxes = idxes // S
        local_idxes = idxes % S
        for idx, global_idx, local_idx in zip(idxes, global_idxes, local_idxes):
            idx, global_idx, local_idx = int(idx), int(global_idx), int(local_idx)
            size = int(self.size_buf[global_idx])
            assert local_idx < size                                                               # :120
            steps = int(min(fwd, size - local_idx))                                                # :122
            if local_idx < bt - 1:                                                                  # :124-127
                obs = self.obs_buf[global_idx * (S + 1):idx + global_idx + 1 + steps]
                comm = self.comm_mask[global_idx * (S + 1):idx + global_idx + 1 + steps]
                hidden = np.zeros((self.n, self.latent), dtype=np.float16)
            elif local_idx == bt - 1:                                                               # :129-132
                obs = self.obs_buf[idx + global_idx + 1 - bt:idx + global_idx + 1 + steps]
                comm = self.comm_mask[global_idx * (S + 1):idx + global_idx + 1 + steps]
                hidden = np.zeros((self.n, self.latent), dtype=np.float16)
            else:                                                                                   # :134-137
                obs = self.obs_buf[idx + global_idx + 1 - bt:idx + global_idx + 1 + steps]
                comm = self.comm_mask[idx + global_idx + 1 - bt:idx + global_idx + 1 + steps]
                hidden = self.hid_buf[idx - bt]
            if obs.shape[0] < bt + fwd:                                                             # :139-142
                pad_len = bt + fwd - obs.shape[0]
                obs = np.pad(obs, ((0, pad_len), (0, 0), (0, 0), (0, 0), (0, 0)))
                comm = np.pad(comm, ((0, pad_len), (0, 0), (0, 0)))
            done = bool(local_idx == size - 1 and self.done_buf[global_idx])                       # :145-148
            b_obs.append(obs), b_action.append(self.act_buf[idx]), b_reward.append(self.rew_buf[idx])
            b_done.append(done), b_steps.append(steps), b_bt_steps.append(min(local_idx + 1, bt))
            b_comm_mask.append(comm), b_hidden.append(hidden)
        min_p = np.min(priorities)                                                                  # :165-166
        weights = np.power(priorities / min_p, -self.beta)
        return (np.stack(b_obs).astype(np.float16), np.asarray(b_action, dtype=np.int64)[:, None],
                np.asarray(b_reward, dtype=np.float16)[:, None], np.asarray(b_done, dtype=np.float16)[:, None],
                np.asarray(b_steps, dtype=np.float16)[:, None], np.asarray(b_bt_steps, dtype=np.int64),
                np.concatenate(b_hidden), np.stack(b_comm_mask), idxes, weights.astype(np.float16)[:, None], self.ptr)

    def update_priorities(self, idxes, priorities, old_ptr):  # :186-203
        S = self.S
        if self.ptr > old_ptr:
            mask = (idxes < old_ptr * S) | (idxes >= self.ptr * S)
            idxes, priorities = idxes[mask], priorities[mask]
        elif self.ptr < old_ptr:
            mask = (idxes < old_ptr * S) & (idxes >= self.ptr * S)
            idxes, priorities = idxes[mask], priorities[mask]
        self.priority_tree.batch_update(np.array(idxes, dtype=np.int64), priorities ** self.alpha)
