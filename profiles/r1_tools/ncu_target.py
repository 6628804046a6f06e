#!/usr/bin/env python
"""Target of the ncu captures of session 4 (GPU box): BASELINE configs[1] geometry; launch order:
reset_kernel, bfs_navi_kernel, 30 whole-batch step launches, 8 steps as 8 rollout chains (64 sub-batch launches, direct),
bfs_navi_kernel again.

    ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_observe_kernel -s 27 -c 8 \
        -o gpurun_out/step_full python profiles/tools/ncu_target.py
    ncu --set full --clock-control none --import-source on -k regex:bfs_navi_kernel -s 1 -c 1 -o gpurun_out/bfs_full \
        python profiles/tools/ncu_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, _native  # noqa: E402

B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)
for s in range(30):
    env.step(acts[s % 16], out_obs=ring[s % 4])
torch.cuda.synchronize()
env.rollout(acts, num_steps=8, out_obs=ring, chains=8)
torch.cuda.synchronize()
_native.check(env._lib.mapf_env_bfs_navi(env._h, None, B, None, env._stream()))
torch.cuda.synchronize()
env.check()
print("ok")
