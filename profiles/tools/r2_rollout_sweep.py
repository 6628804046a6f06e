#!/usr/bin/env python
"""Sweep of the persistent rollout kernel's knobs (mapf_debug_rollout_tuning) on one GPU: device time per step of ONE
mapf_env_rollout call, gated behind a device-side spin (as bench.py times it), median of 5.

    python profiles/tools/r2_rollout_sweep.py [--config c2|c3|c4] [--K 20 2000] [--grid quick|full] [--cap N]

One JSON line per (K, store_mode, warps_per_sm, chunk, stagger_ns, autoreset)."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--K", type=int, nargs="+", default=[20, 2000])
ap.add_argument("--grid", default="quick")
ap.add_argument("--cap", type=int, default=-1)
ap.add_argument("--store", type=int, nargs="+", default=[0, 1])
ap.add_argument("--warps", type=int, nargs="+", default=None)
ap.add_argument("--chunk", type=int, nargs="+", default=None)
ap.add_argument("--stagger", type=int, nargs="+", default=None)
ap.add_argument("--reset", type=int, nargs="+", default=[0, 1])
args = ap.parse_args()
B, N, L, cap = {"c2": (8192, 32, 40, 256), "c3": (8192, 64, 40, 256), "c4": (4096, 64, 80, 32)}[args.config]
if args.cap >= 0:
    cap = args.cap
lib = _native.lib()
env = BatchedEnvironment(B, N, L)
env.reset(seed=0, env_offset=0, density=0.3)
R = 4
ring = torch.empty((R, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
rr = torch.empty((2, B, N), dtype=torch.float32, device="cuda")
rd = torch.empty((2, B), dtype=torch.uint8, device="cuda")
rs = torch.empty((2, B), dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
acts = torch.randint(0, 5, (16, B, N), generator=g, device="cuda", dtype=torch.uint8)
stagger = ((torch.arange(B, device="cuda", dtype=torch.int64) * 2654435761) % max(cap, 1)).to(torch.int32)
peak = 6459.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ab = 486 + 40.5 + 1 + 4 + 2 + 4 + (L * L / 8.0 + 5) / N


def run(k):
    env.rollout(acts, num_steps=k, out_obs=ring, out_rewards=rr, out_done=rd, out_steps=rs)


def timed(k, reps=5):
    ts = []
    for _ in range(reps):
        if cap > 0:
            env.set_state(steps=stagger)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(400_000)
        e0.record()
        run(k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / k)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


warps = args.warps or ([0] if args.grid == "quick" else [0, 12, 20, 24])
chunks = args.chunk or ([0] if args.grid == "quick" else [0, 4, 8])
staggers = args.stagger or ([4000] if args.grid == "quick" else [0, 2000, 4000, 8000])
for K in args.K:
    for store, w, c, sg, rst in itertools.product(args.store, warps, chunks, staggers, args.reset):
        lib.mapf_debug_rollout_tuning(1, w, c, store, sg)
        if rst:
            env.set_autoreset(cap, seed=0, env_offset=B, stride=B, density=0.3)
        else:
            env.set_autoreset(0)
        if cap > 0:
            env.set_state(steps=stagger)
        run(max(K, 32) if K <= 64 else 64)
        torch.cuda.synchronize()
        ep0 = int(env.episode_counts().sum())
        med, best = timed(K)
        ep1 = int(env.episode_counts().sum())
        env.check()
        print(json.dumps({"config": args.config, "K": K, "store_mode": store, "warps_per_sm": w, "chunk": c, "stagger_ns": sg,
                          "autoreset": rst, "us_per_step": round(med, 3), "best_us_per_step": round(best, 3),
                          "frac": round(ab * B * N / (med * 1e-6) / 1e9 / peak, 4), "resets_per_call": (ep1 - ep0) / 5}), flush=True)
lib.mapf_debug_rollout_tuning(1, 0, 0, 0, 1000)
