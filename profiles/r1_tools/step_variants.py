#!/usr/bin/env python
"""Tuning harness (GPU box): time the fused step+observe kernel under the launch variants / flag sets that
mapf_step_kernels.cu reads from MAPF_STEP_VARIANT / MAPF_STEP_FLAGS, on the BASELINE configs[1] workload.

    python profiles/step_variants.py "0:0,1:0,1:1,1:3,2:1" [--envs 8192] [--agents 32] [--side 40] [--steps 300]

Each variant runs in its own process (the library reads the variables once).  Prints one line per variant:
us/step (CUDA events over `steps` back-to-back launches, obs rotating over a 4-slot ring > L2) and a
bit-exactness check of the first 16 envs against the C oracle.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    from bench import make_instances
    from mapf_rl_b200 import BatchedEnvironment
    B, N, L = args.envs, args.agents, args.side
    cache = f"/tmp/mapf_inst_{B}_{N}_{L}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        maps, agents, goals = z["m"], z["a"], z["g"]
    else:
        maps, agents, goals = make_instances(B, L, N, 0.3, 0, 0)
        np.savez(cache, m=maps, a=agents, g=goals)
    dev = torch.device("cuda", 0)
    env = BatchedEnvironment(B, N, L, device=dev)
    env.load(maps, agents, goals)
    R = 4
    replay = torch.empty((R, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    actions = torch.randint(0, 5, (16, B, N), generator=g, device=dev, dtype=torch.uint8)
    # parity of the first envs against the oracle for 4 steps
    from oracle import oracle
    chk = min(16, B)
    ora = []
    for k in range(chk):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        ora.append(o)
    ok = True
    for s in range(4):
        obs, rew, done = env.step(actions[s], out_obs=replay[s % R])
        a = actions[s, :chk].cpu().numpy()
        obs, rew = obs[:chk].cpu().numpy(), rew[:chk].cpu().numpy()
        for k in range(chk):
            (oo, op), orw, od, _ = ora[k].step(a[k])
            ok &= bool(np.array_equal(oo.astype(np.uint8), obs[k]) and np.array_equal(np.asarray(orw, dtype=np.float32), rew[k]))
    t0 = time.perf_counter()
    s = 0
    while time.perf_counter() - t0 < 0.3:
        for _ in range(64):
            env.step(actions[s % 16], out_obs=replay[s % R])
            s += 1
        torch.cuda.synchronize()
    best = None
    for rep in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for s in range(args.steps):
            env.step(actions[s % 16], out_obs=replay[s % R])
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / args.steps
        best = us if best is None else min(best, us)
    env.check()
    print(json.dumps({"variant": os.environ.get("MAPF_STEP_VARIANT"), "flags": os.environ.get("MAPF_STEP_FLAGS"), "ctas_per_sm": os.environ.get("MAPF_STEP_CTAS_PER_SM"),
                      "us_per_step": round(best, 2), "G_agent_steps_s": round(B * N / best / 1e3, 3),
                      "parity16": ok, "B": B, "N": N, "L": L}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="?", default="0:0,1:0,1:1,1:3")
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--agents", type=int, default=32)
    ap.add_argument("--side", type=int, default=40)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        child(args)
        return
    for spec in args.variants.split(","):
        parts = spec.split(":")
        v, f = parts[0], parts[1]
        env = dict(os.environ, MAPF_STEP_VARIANT=v, MAPF_STEP_FLAGS=f)
        if len(parts) > 2:
            env["MAPF_STEP_CTAS_PER_SM"] = parts[2]
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--envs", str(args.envs), "--agents", str(args.agents),
               "--side", str(args.side), "--steps", str(args.steps)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=180)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else f"variant {spec} FAILED rc={r.returncode}: {r.stderr[-600:]}", flush=True)


if __name__ == "__main__":
    main()
