#!/usr/bin/env python
"""Where the time of the host-buffer step goes: per-call wall time of mapf_env_step_host_codes (and the fp32 form) in a tight
loop, plus the device time of the kernel alone, under the diagnosis switches of MAPF_HOSTCODES_DBG
(1 = no per-warp system fence, 2 = results into device memory, 4 = actions from device memory).

    MAPF_HOSTCODES_DBG=n python profiles/tools/r2_e2e_probe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402

B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, env_offset=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)
ah = acts.cpu().pin_memory()
ah = [ah[i] for i in range(16)]
slots = [ring[i] for i in range(4)]
res = {"dbg": os.environ.get("MAPF_HOSTCODES_DBG", "0")}
for name, fn in (("codes", lambda s: env.step_host_codes(ah[s % 16], device_obs=slots[s % 4])),
                 ("f32", lambda s: env.step_host(ah[s % 16], device_obs=slots[s % 4]))):
    for s in range(32):
        fn(s)
    torch.cuda.synchronize()
    T = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for s in range(T):
        fn(s)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / T * 1e6
    res[name] = {"wall_us_per_call": round(wall, 1), "device_us_per_call": round(e0.elapsed_time(e1) * 1e3 / T, 1)}
    # host time of a call alone (GPU idle between calls)
    t = 0.0
    for s in range(50):
        torch.cuda.synchronize()
        a = time.perf_counter()
        fn(s)
        t += time.perf_counter() - a
    res[name]["isolated_call_us"] = round(t / 50 * 1e6, 1)
# the fused kernel alone, device-resident actions
rew = torch.empty((B, N), dtype=torch.float32, device="cuda")
for s in range(20):
    env.step(acts[s % 16], out_obs=slots[s % 4])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(200):
    env.step(acts[s % 16], out_obs=slots[s % 4])
e1.record()
torch.cuda.synchronize()
res["fused_kernel_device_actions_us"] = round(e0.elapsed_time(e1) * 1e3 / 200, 1)
print(json.dumps(res))
