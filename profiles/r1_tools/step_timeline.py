#!/usr/bin/env python
"""Timeline of one launch of the split step kernel (GPU box): per-environment %globaltimer stamps written by the
kernel itself (mapf_debug_step_trace) -> when producers start / finish, when consumers get / finish each env.

    python profiles/step_timeline.py [--variants 10,17] [--ctas 0] [--envs 8192] [--agents 32] [--side 40]

Prints, per variant, percentiles of every stamp (us since the first producer started) and how many observation
bytes had been handed to the memory system by each microsecond.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["prod_reach", "slot_free", "step_done", "published", "cons_reach", "cons_got", "stores_issued", "inputs_loaded",
         "conflicts_done", "gather_begin", "gather_end"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="10,17")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--agents", type=int, default=32)
    ap.add_argument("--side", type=int, default=40)
    ap.add_argument("--dump", default=None)
    args = ap.parse_args()
    import torch
    from mapf_rl_b200 import BatchedEnvironment, _native
    lib = _native.lib()
    B, N, L = args.envs, args.agents, args.side
    dev = torch.device("cuda", 0)
    env = BatchedEnvironment(B, N, L, device=dev)
    env.reset(seed=0, density=0.3)
    ring = torch.empty((4, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    actions = torch.randint(0, 5, (16, B, N), generator=g, device=dev, dtype=torch.uint8)
    trace = torch.zeros((B, 16), dtype=torch.int64, device=dev)
    for v in args.variants.split(","):
        for ctas in args.ctas.split(","):
            lib.mapf_debug_step_tuning(int(v), -1, int(ctas))
            lib.mapf_debug_step_trace(None)
            for s in range(50):
                env.step(actions[s % 16], out_obs=ring[s % 4])
            torch.cuda.synchronize()
            lib.mapf_debug_step_trace(C.c_void_p(trace.data_ptr()))
            for s in range(3):
                env.step(actions[s % 16], out_obs=ring[s % 4])
            torch.cuda.synchronize()
            lib.mapf_debug_step_trace(None)
            t = trace.cpu().numpy()[:, :11].astype(np.float64)
            t0 = t[:, 0].min()
            t = (t - t0) / 1e3
            print(json.dumps({"variant": v, "ctas_per_sm": ctas, "span_us": round(float(t[:, 6].max()), 2)}))
            for k, name in enumerate(NAMES):
                q = np.percentile(t[:, k], [0, 10, 50, 90, 100])
                print(f"  {name:14s} min {q[0]:7.2f}  p10 {q[1]:7.2f}  p50 {q[2]:7.2f}  p90 {q[3]:7.2f}  max {q[4]:7.2f}")
            d = {"produce (slot_free->published)": t[:, 3] - t[:, 1], "step phase (slot_free->step_done)": t[:, 2] - t[:, 1],
                 "  inputs loaded (slot_free->in smem)": t[:, 7] - t[:, 1], "  conflict resolution": t[:, 8] - t[:, 7],
                 "  commit (stores of state)": t[:, 2] - t[:, 8],
                 "gather (step_done->published)": t[:, 3] - t[:, 2], "  agent bitmap + navi loads issued": t[:, 9] - t[:, 2],
                 "  field walk (waits for navi)": t[:, 10] - t[:, 9], "  tail (first words, syncs)": t[:, 3] - t[:, 10],
                 "wait for slot": t[:, 1] - t[:, 0],
                 "published->consumer got": t[:, 5] - t[:, 3], "consumer waits for stream": t[:, 5] - t[:, 4],
                 "expand+store (got->issued)": t[:, 6] - t[:, 5]}
            first = t[:, 0] < 1.0   # environments a producer took in its first round
            for name, x in d.items():
                q = np.percentile(x, [10, 50, 90])
                print(f"  {name:38s} p10 {q[0]:6.2f}  p50 {q[1]:6.2f}  p90 {q[2]:6.2f}  mean {x.mean():6.2f}   "
                      f"first round mean {x[first].mean():6.2f}  later {x[~first].mean() if (~first).any() else 0:6.2f}")
            done = np.sort(t[:, 6])
            edges = np.arange(0, done.max() + 2, 2.0)
            cum = np.searchsorted(done, edges, side="right")
            print("  envs with stores issued by t (us): " + "  ".join(f"{int(e)}:{c}" for e, c in zip(edges, cum)))
            pub = np.sort(t[:, 3])
            cum = np.searchsorted(pub, edges, side="right")
            print("  envs published by t (us):          " + "  ".join(f"{int(e)}:{c}" for e, c in zip(edges, cum)))
            if args.dump:
                np.save(f"{args.dump}_v{v}_c{ctas}.npy", t)


if __name__ == "__main__":
    main()
