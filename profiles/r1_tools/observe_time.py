#!/usr/bin/env python
"""us per launch of the observe-only kernel (Environment.observe) at 8192 x 32 agents, 40x40; MAPF_STEP_VARIANT selects the CTA shape."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mapf_rl_b200 import BatchedEnvironment
B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
for s in range(200):
    env.observe(out_obs=ring[s % 4])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(400):
        env.observe(out_obs=ring[s % 4])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e3 / 400)
print(json.dumps({"variant": os.environ.get("MAPF_STEP_VARIANT", "default"), "observe_us": round(best, 2)}))
