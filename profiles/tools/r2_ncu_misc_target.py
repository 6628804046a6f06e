#!/usr/bin/env python
"""Target for ONE ncu --set full capture of every kernel of the library other than the rollout kernel, each launched at
the size it runs at in BASELINE's configs (warm-up launches first, then exactly one profiled launch between two marker
prints):

    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o gpurun_out/r2_misc \
        python profiles/tools/r2_ncu_misc_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, ReplayStore, SumTree, config  # noqa: E402
from mapf_rl_b200.buffer import actor_td_errors  # noqa: E402

nvtx = torch.cuda.nvtx
dev = "cuda:0"


def cap(fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    nvtx.range_push("cap")
    fn()
    torch.cuda.synchronize()
    nvtx.range_pop()


# ---- C2 geometry: 8192 x 32 agents, 40x40 ----
B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L, device=dev)
cap(lambda: env.reset(seed=0, env_offset=0, density=0.3), warm=1)                 # reset_kernel<2,2> + bfs_navi_kernel<2,3,2>
obs = torch.empty((B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(0)
acts = torch.randint(0, 5, (B, N), generator=g, device=dev, dtype=torch.uint8)
cap(lambda: env.step(acts, out_obs=obs))                                          # step_observe_kernel<2,1,true,8,5>
cap(lambda: env.observe(out_obs=obs))                                             # step_observe_kernel<2,1,false,...>
cap(lambda: env.comm_mask())                                                      # comm_mask_kernel
ah = acts.cpu().pin_memory()
cap(lambda: env.step_host_codes(ah, device_obs=obs))                              # step_only_kernel<2,1> + observe
env.check()
env.close()
# ---- C4 geometry: 4096 x 64 agents, 80x80 ----
B4, N4, L4 = 4096, 64, 80
env4 = BatchedEnvironment(B4, N4, L4, device=dev)
cap(lambda: env4.reset(seed=0, env_offset=0, density=0.3), warm=1)                # reset_kernel<3,3> + bfs_navi_kernel<3,3,1>
obs4 = torch.empty((B4, N4, 6, 9, 9), dtype=torch.uint8, device=dev)
acts4 = torch.randint(0, 5, (B4, N4), generator=g, device=dev, dtype=torch.uint8)
cap(lambda: env4.step(acts4, out_obs=obs4))                                       # step_observe_kernel<3,2,true,...>
env4.check()
env4.close()
# ---- PER at the reference's capacity (2048 slots x 256 steps = 2^19 leaves), batch 192 ----
tree = SumTree(1 << 19, device=dev)
rng = np.random.default_rng(0)
idx0 = torch.as_tensor(rng.permutation(1 << 19)[:200000].astype(np.int64)).cuda()
tree.update_device(idx0, torch.rand(200000, dtype=torch.float64, device=dev) + 1e-3)   # per_claim / per_leaf / per_level kernels
n = config.batch_size
u = torch.rand(n, dtype=torch.float64, device=dev)
idx = torch.as_tensor(rng.integers(0, 1 << 19, size=n)).cuda()
pr = torch.rand(n, dtype=torch.float64, device=dev) + 1e-3
cap(lambda: tree.update_device(idx, pr))                                          # per_update_kernel
cap(lambda: tree.sample_device(n, u, beta=0.4))                                   # per_sample_kernel
upd = dict(q_online=torch.randn(n, 5, device=dev), q_target_next=torch.randn(n, 5, device=dev),
           action=torch.randint(0, 5, (n,), device=dev), reward=torch.zeros(n, device=dev), done=torch.zeros(n, device=dev),
           steps=torch.ones(n, device=dev), idx=idx)
cap(lambda: tree.td_update(upd["q_online"], upd["q_target_next"], upd["action"], upd["reward"], upd["done"], upd["steps"], idx))  # per_td_update_kernel
cap(lambda: tree.cycle(update=upd, sample_size=n, uniforms=u, beta=0.4))          # per_cycle_kernel
ep = torch.arange(256 * 64, device=dev)
cap(lambda: tree.update_device(ep, torch.rand(256 * 64, dtype=torch.float64, device=dev)))   # 64 episodes inserted at once
E = 64
cap(lambda: actor_td_errors(torch.zeros((E, 256), device=dev), torch.randn((E, 256, 5), device=dev),
                            torch.zeros((E, 256), dtype=torch.uint8, device=dev), torch.full((E,), 200, dtype=torch.int32, device=dev)))
# ---- replay window gather: 192 samples x 18 frames x 32 agents ----
store = ReplayStore(256, max_num_agents=32, device=dev, max_steps=64)
store.size_buf.fill_(64)
gi = torch.as_tensor(rng.integers(0, 256 * 64, size=n)).cuda()
cap(lambda: store.gather(gi))                                                     # replay_gather_kernel
print("ok")
