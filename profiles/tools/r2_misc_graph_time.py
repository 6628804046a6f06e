#!/usr/bin/env python
"""GPU-side time of the small kernels around the environment (GPU box): comm mask at 8192 x 32 agents, replay window gather at
batch 192 x 18 frames x 32 agents; 32 captured calls per graph replay so that the Python / launch cost drops out.

    python profiles/tools/r2_misc_graph_time.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, ReplayStore  # noqa: E402


def gtime(fn, reps=32, launches=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / (reps * launches), 2)


env = BatchedEnvironment(8192, 32, 40, device="cuda:0")
env.reset(seed=0, density=0.3)
out = env.comm_mask()
res = {"comm_mask_8192x32_us": gtime(lambda: env.comm_mask(out=out))}

store = ReplayStore(256, device="cuda:0")
store.obs_buf.random_(0, 2)
store.comm_mask.random_(0, 2)
store.size_buf.fill_(256)
g = torch.Generator(device="cuda:0")
g.manual_seed(0)
idx = torch.randint(0, 256 * 256, (192,), generator=g, device="cuda:0")
res["replay_gather_192x18x32_us"] = gtime(lambda: store.gather(idx), reps=8)
print(json.dumps(res))
