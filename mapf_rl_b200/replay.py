"""ReplayStore — device-resident mirror of the storage and sampling half of the reference's `GlobalBuffer`
(worker.py:21-203, without Ray and without the curriculum statistics).

The buffers keep the reference's logical layout (worker.py:36-42) as CUDA tensors: episode slot g owns
observation / comm-mask rows g*(max_steps+1)+f and action / reward / hidden rows g*max_steps+t, and the
priority tree leaf of transition (g, t) is g*max_steps+t.  Observations are stored as the bool bytes the
step kernel emits, so a batched actor can point `BatchedEnvironment.step(out_obs=...)` straight at rows
of `obs_buf`.  `sample_batch` is two launches: the sum-tree descent (mapf_per_sample) and the window
gather with bool->fp16 conversion (mapf_replay_gather); `update_priorities` is the fused stale-mask +
tree update.  Return values follow the reference's tuple (worker.py:168-182) with CUDA tensors in
place of CPU ones.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _native, config
from .buffer import SumTree


def _torch():
    import torch
    return torch


class ReplayStore:
    def __init__(self, capacity: int, alpha=config.prioritized_replay_alpha, beta=config.prioritized_replay_beta,
                 max_num_agents: int = config.max_num_agents, device=None, max_steps: int = config.max_steps,
                 bt_steps: int = config.bt_steps, forward_steps: int = config.forward_steps,
                 latent_dim: int = config.latent_dim):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("mapf_rl_b200 needs a CUDA device: the replay kernels have no CPU fallback")
        self._lib = _native.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.capacity, self.alpha, self.beta = int(capacity), alpha, beta
        self.max_num_agents, self.max_steps = int(max_num_agents), int(max_steps)
        self.bt_steps, self.forward_steps, self.latent_dim = int(bt_steps), int(forward_steps), int(latent_dim)
        self.size = 0       # stored transitions (worker.py:25)
        self.ptr = 0        # next episode slot (worker.py:26)
        self.counter = 0
        self.priority_tree = SumTree(self.capacity * self.max_steps, device=self.device)   # worker.py:27
        n, S, dev = self.max_num_agents, self.max_steps, self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.obs_buf = z(((S + 1) * capacity, n, *config.obs_shape), torch.uint8)          # worker.py:36
        self.act_buf = z((S * capacity,), torch.uint8)                                     # worker.py:37
        self.rew_buf = z((S * capacity,), torch.float16)                                   # worker.py:38
        self.hid_buf = z((S * capacity, n, self.latent_dim), torch.float16)                # worker.py:39
        self.done_buf = z((capacity,), torch.uint8)                                        # worker.py:40
        self.size_buf = z((capacity,), torch.int32)                                        # worker.py:41
        self.comm_mask = z(((S + 1) * capacity, n, n), torch.uint8)                        # worker.py:42
        self._size_host = np.zeros(capacity, dtype=np.int64)
        self._err = z((1,), torch.int32)

    def __len__(self):
        return self.size

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    # -- GlobalBuffer.add (worker.py:68-104) ------------------------------------------------------
    def add(self, buffer_list: List):
        """buffer_list: tuples as returned by LocalBuffer.finish — actor_id 0, num_agents 1, map_len 2, obs_buf 3,
        act_buf 4, rew_buf 5, hid_buf 6, td_errors 7, done 8, size 9, comm_mask 10 (worker.py:72)."""
        torch = _torch()
        S = self.max_steps
        dev = self.device

        def up(x, dt):
            return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=dt, non_blocking=True)

        for buffer in buffer_list:
            n, size = int(buffer[1]), int(buffer[9])
            idxes = np.arange(self.ptr * S, (self.ptr + 1) * S, dtype=np.int64)            # worker.py:88
            start_idx = self.ptr * S
            self.size -= int(self._size_host[self.ptr])
            self.size += size
            self.counter += size
            self.priority_tree.batch_update(idxes, np.asarray(buffer[7], dtype=np.float64) ** self.alpha)   # worker.py:94
            row0 = start_idx + self.ptr                                                    # = ptr * (S + 1)
            self.obs_buf[row0:row0 + size + 1, :n] = up(buffer[3], torch.uint8)            # worker.py:96
            self.act_buf[start_idx:start_idx + size] = up(buffer[4], torch.uint8)
            self.rew_buf[start_idx:start_idx + size] = up(np.asarray(buffer[5], dtype=np.float16), torch.float16)
            self.hid_buf[start_idx:start_idx + size, :n] = up(np.asarray(buffer[6], dtype=np.float16), torch.float16)
            self.done_buf[self.ptr] = int(bool(buffer[8]))
            self.size_buf[self.ptr] = size
            self._size_host[self.ptr] = size
            self.comm_mask[row0:row0 + size + 1, :n, :n] = up(buffer[10], torch.uint8)     # worker.py:102
            self.ptr = (self.ptr + 1) % self.capacity

    # -- direct-write path of a batched actor -------------------------------------------------------
    def obs_rows(self, slot: int, first_frame: int, count: int):
        """View of `count` consecutive observation frames of episode slot `slot` — the tensor a batched actor
        passes as `out_obs` so the step kernel writes into the replay store with no copy."""
        row0 = slot * (self.max_steps + 1) + first_frame
        return self.obs_buf[row0:row0 + count]

    # -- GlobalBuffer.sample_batch (worker.py:106-184) ------------------------------------------------
    def _view(self):
        return _native.ReplayView(self.obs_buf.data_ptr(), self.comm_mask.data_ptr(), self.hid_buf.data_ptr(),
                                  self.act_buf.data_ptr(), self.rew_buf.data_ptr(), self.done_buf.data_ptr(),
                                  self.size_buf.data_ptr(), self.max_num_agents, self.max_steps, self.bt_steps,
                                  self.forward_steps, self.latent_dim)

    def gather(self, idxes):
        """Window gather for the given leaf indices (int64 CUDA tensor [B]) -> dict of CUDA tensors."""
        torch = _torch()
        idx = torch.as_tensor(idxes, dtype=torch.int64).to(self.device).contiguous()
        B, n, W, dev = int(idx.numel()), self.max_num_agents, self.bt_steps + self.forward_steps, self.device
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        out = dict(obs=e((B, W, n, *config.obs_shape), torch.float16), comm_mask=e((B, W, n, n), torch.uint8),
                   hidden=e((B * n, self.latent_dim), torch.float16), action=e((B,), torch.int64),
                   reward=e((B,), torch.float16), done=e((B,), torch.float16), steps=e((B,), torch.float16),
                   bt_steps=e((B,), torch.int64))
        view = self._view()
        batch = _native.ReplayBatch(*[out[k].data_ptr() for k in ("obs", "comm_mask", "hidden", "action", "reward", "done",
                                                                   "steps", "bt_steps")])
        _native.check(self._lib.mapf_replay_gather(C.byref(view), C.c_void_p(idx.data_ptr()), B, C.byref(batch),
                                                   C.c_void_p(self._err.data_ptr()), self._stream()))
        return out

    def sample_batch(self, batch_size: int, uniforms=None, check: bool = True):
        """-> (obs f16[B,W,n,6,9,9], action i64[B,1], reward f16[B,1], done f16[B,1], steps f16[B,1], bt_steps i64[B],
        hidden f16[B*n,latent], comm_mask bool[B,W,n,n], idxes int64 numpy[B], weights f16[B,1], ptr) — the reference's
        tuple (worker.py:168-182) with CUDA tensors.  `uniforms` (optional, [B] in [0,1)) replaces the draw
        np.random.uniform makes inside SumTree.batch_sample (buffer.py:60)."""
        torch = _torch()
        if uniforms is None:
            uniforms = np.random.random_sample(batch_size)
        idx, prio, _ = self.priority_tree.sample_device(batch_size, uniforms)
        out = self.gather(idx)
        # importance sampling weights (worker.py:165-166), fp64 like numpy, then fp16 (:181)
        weights = (prio / prio.min()).pow(-self.beta).to(torch.float16).unsqueeze(1)
        idx_host = idx.cpu().numpy()
        if check:
            if int(self._err.item()) != 0:
                self._err.zero_()
                raise AssertionError("sampled transition lies beyond its episode (worker.py:120)")
        return (out["obs"], out["action"].unsqueeze(1), out["reward"].unsqueeze(1), out["done"].unsqueeze(1),
                out["steps"].unsqueeze(1), out["bt_steps"], out["hidden"], out["comm_mask"].bool(), idx_host, weights,
                self.ptr)

    # -- GlobalBuffer.update_priorities (worker.py:186-203) --------------------------------------------
    def update_priorities(self, idxes: np.ndarray, priorities: np.ndarray, old_ptr: int):
        idxes = np.asarray(idxes)
        priorities = np.asarray(priorities)
        S = self.max_steps
        if self.ptr > old_ptr:      # discard the slots overwritten since sampling: [old_ptr, ptr)
            mask = (idxes < old_ptr * S) | (idxes >= self.ptr * S)
            idxes, priorities = idxes[mask], priorities[mask]
        elif self.ptr < old_ptr:    # [0, ptr) and [old_ptr, capacity)
            mask = (idxes < old_ptr * S) & (idxes >= self.ptr * S)
            idxes, priorities = idxes[mask], priorities[mask]
        # numpy evaluates priorities**alpha in the dtype the learner sent (fp16 in the reference)
        self.priority_tree.batch_update(np.array(idxes, dtype=np.int64), priorities ** self.alpha)

    def update_priorities_device(self, q_online, q_target_next, action, reward, done, steps, idxes, old_ptr: int,
                                 gamma: float = 0.99, q_online_next=None):
        """Learner tail on the device: TD error, priority, stale mask and tree update in one launch
        (SumTree.td_update).  Returns (td, priority) CUDA tensors."""
        return self.priority_tree.td_update(q_online, q_target_next, action, reward, done, steps, idxes, old_ptr=old_ptr,
                                            ptr=self.ptr, slot_steps=self.max_steps, gamma=gamma, alpha=self.alpha,
                                            q_online_next=q_online_next)

    def ready(self):  # worker.py:228-232
        return len(self) >= config.learning_starts
