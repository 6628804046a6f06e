#!/usr/bin/env python
"""Target for an ncu capture of the persistent rollout kernel: one mapf_env_rollout call of 32 steps at 8192 x 32 agents, 40x40.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
        --clock-control none --cache-control none -k regex:step_rollout_kernel python profiles/tools/ncu_rollout_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402

B, N, L, T = 8192, 32, 40, 32
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
rr = torch.empty((2, B, N), dtype=torch.float32, device="cuda")
rd = torch.empty((2, B), dtype=torch.uint8, device="cuda")
rs = torch.empty((2, B), dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)
for _ in range(2):
    env.rollout(acts, num_steps=T, out_obs=ring, out_rewards=rr, out_done=rd, out_steps=rs)
torch.cuda.synchronize()
env.check()
print("ok", env.rollout_plan(T, 16, 4, 2))
