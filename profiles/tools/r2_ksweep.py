#!/usr/bin/env python
"""K-sweep of mapf_env_rollout at BASELINE configs[1] (8192 x 32 agents, 40x40): device time per step of ONE call of K steps,
(a) with the events recorded on an idle stream (the host's enqueue latency of the call sits inside the timed region) and
(b) behind a device-side gate (torch.cuda._sleep queued first, so the call is already enqueued when the first event fires).

    python profiles/tools/r2_ksweep.py [K ...]      -> one JSON line per K
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402

Ks = [int(a) for a in sys.argv[1:]] or [20, 64, 256, 2000]
B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
rr = torch.empty((2, B, N), dtype=torch.float32, device="cuda")
rd = torch.empty((2, B), dtype=torch.uint8, device="cuda")
rs = torch.empty((2, B), dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)


def run(k):
    env.rollout(acts, num_steps=k, out_obs=ring, out_rewards=rr, out_done=rd, out_steps=rs)


run(256)
torch.cuda.synchronize()
for K in Ks:
    res = {"K": K}
    for gate in (False, True):
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if gate:
                torch.cuda._sleep(400_000)   # ~200 us at 1.9 GHz: the rollout call is enqueued while this spins
            e0.record()
            run(K)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / K)
        ts.sort()
        res["gated_us_per_step" if gate else "idle_us_per_step"] = {"min": ts[0], "median": ts[len(ts) // 2], "max": ts[-1]}
    print(json.dumps(res), flush=True)
env.check()
