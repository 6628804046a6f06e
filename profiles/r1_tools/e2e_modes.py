#!/usr/bin/env python
"""Kernel duration of the host-buffer step per MAPF_STEP_HOST_MODE (GPU box; run under
`ncu --metrics gpu__time_duration.sum` for the kernel times, plain for the end-to-end times)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402

B, N, L = 8192, 32, 40
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, density=0.3)
ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(0)
acts = rng.integers(0, 5, size=(16, B, N)).astype(np.uint8)
if os.environ.get("E2E_PINNED_ACTIONS", "1") != "0":   # page-locked action buffers handed over in place (no staging copy)
    acts = torch.as_tensor(acts).pin_memory()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for s in range(10):
    env.step_host(acts[s % 16], device_obs=ring[s % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(steps):
    env.step_host(acts[s % 16], device_obs=ring[s % 4])
torch.cuda.synchronize()
el = time.perf_counter() - t0
print(f"mode {os.environ.get('MAPF_STEP_HOST_MODE', 'default')}: {el / steps * 1e6:.1f} us per host step, "
      f"{B * N * steps / el / 1e9:.2f} G agent-steps/s")
# GPU span of one host step (events on the stream around the call) next to its wall time
spans, walls = [], []
for s in range(50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    env.step_host(acts[s % 16], device_obs=ring[s % 4])
    e1.record()
    walls.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    spans.append(e0.elapsed_time(e1) * 1e3)
print(f"   GPU span median {np.median(spans):.1f} us, wall median {np.median(walls) * 1e6:.1f} us (includes the two event records)")
# host-side cost of the call path alone: copy of the actions into the pinned buffer
hb = env.host_actions
acts_np = acts.numpy() if isinstance(acts, torch.Tensor) else acts
t0 = time.perf_counter()
for s in range(200):
    np.copyto(hb, acts_np[s % 16])
print(f"   np.copyto of the actions: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
