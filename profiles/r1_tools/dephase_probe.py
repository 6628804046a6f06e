#!/usr/bin/env python
"""Probe (GPU box): does running the batch as S independent sub-batches on S streams, so that their kernels are in
different phases (one computes while another stores), beat one launch over the whole batch?

    python profiles/tools/dephase_probe.py [--envs 8192] [--splits 1,2,4,8]

Each sub-batch is its own BatchedEnvironment (B / S envs) stepped on its own torch stream; step k+1 of a sub-batch
depends only on step k of the same sub-batch.  Prints us per whole-batch step (CUDA events bracketing all streams).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment  # noqa: E402


def run(B, N, L, S, steps, stagger):
    dev = torch.device("cuda", 0)
    b = B // S
    envs, rings, acts, streams = [], [], [], []
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    for j in range(S):
        e = BatchedEnvironment(b, N, L, device=dev)
        e.reset(seed=0, env_offset=j * b, density=0.3)
        envs.append(e)
        rings.append(torch.empty((4, b, N, 6, 9, 9), dtype=torch.uint8, device=dev))
        acts.append(torch.randint(0, 5, (16, b, N), generator=g, device=dev, dtype=torch.uint8))
        streams.append(torch.cuda.Stream(device=dev))
    torch.cuda.synchronize()

    def loop(n):
        for s in range(n):
            for j in range(S):
                with torch.cuda.stream(streams[j]):
                    envs[j].step(acts[j][s % 16], out_obs=rings[j][s % 4])

    loop(300)
    torch.cuda.synchronize()
    # the Python call path costs ~19 us per launch, so S > 1 would be launch-bound: replay 16 captured steps per stream
    graphs = []
    for j in range(S):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=streams[j]):
            for s in range(16):
                envs[j].step(acts[j][s % 16], out_obs=rings[j][s % 4])
        graphs.append(gr)
    torch.cuda.synchronize()

    def loop(n):  # noqa: F811
        for s in range(n // 16):
            for j in range(S):
                with torch.cuda.stream(streams[j]):
                    graphs[j].replay()

    steps = (steps // 16) * 16
    loop(64)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        ev0 = torch.cuda.Event(enable_timing=True)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
        torch.cuda.synchronize()
        ev0.record(torch.cuda.current_stream())
        for j in range(S):
            streams[j].wait_event(ev0)
        if stagger and S > 1:
            # offset the sub-batches by a fraction of a step: sub-batch j first spins on a tiny dummy workload
            for j in range(1, S):
                with torch.cuda.stream(streams[j]):
                    torch.cuda._sleep(int(1965 * 34 * j / S))
        loop(steps)
        for j in range(S):
            evs[j].record(streams[j])
        torch.cuda.synchronize()
        us = max(ev0.elapsed_time(e) for e in evs) * 1e3 / steps
        best = us if best is None else min(best, us)
    for e in envs:
        e.check()
    print(json.dumps({"probe": "dephase", "splits": S, "envs_per_split": b, "stagger": stagger, "us_per_batch_step": round(best, 2),
                      "G_agent_steps_s": round(B * N / best / 1e3, 3)}), flush=True)
    for e in envs:
        e.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--agents", type=int, default=32)
    ap.add_argument("--side", type=int, default=40)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--splits", default="1,2,4,8")
    args = ap.parse_args()
    for S in [int(x) for x in args.splits.split(",")]:
        for stagger in ([False] if S == 1 else [False, True]):
            run(args.envs, args.agents, args.side, S, args.steps, stagger)


if __name__ == "__main__":
    main()
