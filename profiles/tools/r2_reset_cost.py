#!/usr/bin/env python
"""What one in-launch re-generation costs (generator + BFS of every agent by ONE warp inside rollout_kernel), against the
dedicated launches (reset_kernel + bfs_navi_kernel over the whole batch):

  * `B` environments, all at the step cap, ONE rollout step: every environment re-generates at once; B = one per SM, one per
    resident warp, the whole batch -> latency of a lone re-generation, under full contention, and per-reset warp time;
  * env.reset() of the same batch (two launches at full occupancy).

    python profiles/tools/r2_reset_cost.py [--config c2|c3|c4]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
args = ap.parse_args()
Bfull, N, L = {"c2": (8192, 32, 40), "c3": (8192, 64, 40), "c4": (4096, 64, 80)}[args.config]
cap = 8


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(400_000)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for B in (148, 148 * 4, 148 * 16, Bfull):
    env = BatchedEnvironment(B, N, L)
    env.reset(seed=0, env_offset=0, density=0.3)
    obs = torch.empty((1, B, N, 6, 9, 9), dtype=torch.uint8, device="cuda")
    rew = torch.empty((1, B, N), dtype=torch.float32, device="cuda")
    done = torch.empty((1, B), dtype=torch.uint8, device="cuda")
    steps = torch.empty((1, B), dtype=torch.int32, device="cuda")
    acts = torch.zeros((1, B, N), dtype=torch.uint8, device="cuda")
    at_cap = torch.full((B,), cap, dtype=torch.int32, device="cuda")
    zero = torch.zeros((B,), dtype=torch.int32, device="cuda")

    def one_step():
        env.rollout(acts, num_steps=1, out_obs=obs, out_rewards=rew, out_done=done, out_steps=steps)

    env.set_autoreset(cap, seed=0, env_offset=B, stride=B, density=0.3)
    ts = {}
    for name, st in (("all_reset", at_cap), ("no_reset", zero)):
        def run():
            env.set_state(steps=st)
            return timed(one_step, reps=1)
        run()
        xs = sorted(run() for _ in range(5))
        ts[name] = xs[2]
    env.set_autoreset(0)
    t_kernels = timed(lambda: env.reset(seed=1, env_offset=0, density=0.3))
    env.check()
    print(json.dumps({"config": args.config, "B": B, "N": N, "L": L, "one_step_all_envs_regenerate_us": round(ts["all_reset"], 1),
                      "one_step_no_regeneration_us": round(ts["no_reset"], 1),
                      "reset_kernel_plus_bfs_kernel_us": round(t_kernels, 1)}), flush=True)
    env.close()
