#!/usr/bin/env python
"""Target for ncu captures of the persistent rollout kernel (one mapf_env_rollout call of K steps after a warm-up call).

    ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/X \
        python profiles/tools/r2_ncu_rollout_target.py [--config c2|c3|c4] [--K 64] [--reset 0|1] [--store 0|1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--K", type=int, default=64)
ap.add_argument("--reset", type=int, default=0)
ap.add_argument("--store", type=int, default=0)
ap.add_argument("--calls", type=int, default=2)
a = ap.parse_args()
B, N, L, cap = {"c2": (8192, 32, 40, 256), "c3": (8192, 64, 40, 256), "c4": (4096, 64, 80, 32)}[a.config]
_native.lib().mapf_debug_rollout_tuning(1, -1, -1, a.store, -1)
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, env_offset=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
rr = torch.empty((2, B, N), dtype=torch.float32, device="cuda")
rd = torch.empty((2, B), dtype=torch.uint8, device="cuda")
rs = torch.empty((2, B), dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)
if a.reset:
    env.set_autoreset(cap, seed=0, env_offset=B, stride=B, density=0.3)
    env.set_state(steps=((torch.arange(B, device="cuda", dtype=torch.int64) * 2654435761) % cap).to(torch.int32))
for _ in range(a.calls):
    env.rollout(acts, num_steps=a.K, out_obs=ring, out_rewards=rr, out_done=rd, out_steps=rs)
torch.cuda.synchronize()
env.check()
print("ok", a, int(env.episode_counts().sum()))
