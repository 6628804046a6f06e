#!/usr/bin/env python
"""Per-item timeline of ONE rollout launch with episode handling (diagnosis build of the library only):

    MAPF_ENABLE_DIAG=1 python -m mapf_rl_b200.build --diag
    MAPF_B200_LIB=mapf_rl_b200/libmapf_b200_diag.so python profiles/tools/r2_rollout_timeline.py [--config c2] [--steps 20]

Every work item of rollout_kernel records %globaltimer when it is claimed, when its environment's state is loaded, the time it
spent re-generating / adopting an instance and when it is done, plus SM and warp.  Prints where the launch's warp-time goes."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--pregen", type=int, default=1)
ap.add_argument("--cap", type=int, default=0)
ap.add_argument("--dump", default="")
ap.add_argument("--no-trace", action="store_true", help="product build: only run the launches (target of an ncu launch list)")
args = ap.parse_args()
B, N, L, cap = {"c2": (8192, 32, 40, 256), "c3": (8192, 64, 40, 256), "c4": (4096, 64, 80, 32)}[args.config]
cap = args.cap or cap
K, A, R = args.steps, 16, 4
lib = _native.lib()
if not args.no_trace:
    assert hasattr(lib, "mapf_diag_rollout_trace"), "needs the diagnosis build (MAPF_B200_LIB=...libmapf_b200_diag.so)"
    lib.mapf_diag_rollout_trace.argtypes = [C.c_void_p, C.c_int]
lib.mapf_debug_rollout_pregen(args.pregen)

dev = torch.device("cuda", 0)
env = BatchedEnvironment(B, N, L, device=dev)
env.reset(seed=0, env_offset=0, density=0.3)
replay = torch.empty((R, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(0)
actions = torch.randint(0, 5, (A, B, N), generator=g, device=dev, dtype=torch.uint8)
rew = torch.empty((2, B, N), dtype=torch.float32, device=dev)
done = torch.empty((2, B), dtype=torch.uint8, device=dev)
steps = torch.empty((2, B), dtype=torch.int32, device=dev)
stagger = ((torch.arange(B, device=dev, dtype=torch.int64) * 2654435761) % cap).to(torch.int32)
env.set_autoreset(cap, seed=0, env_offset=B, stride=B, density=0.3)


def rollout(k):
    env.rollout(actions, num_steps=k, out_obs=replay, out_rewards=rew, out_done=done, out_steps=steps)


for _ in range(3):
    env.set_state(steps=stagger)
    rollout(max(K, 32))
env.set_state(steps=stagger)
if args.no_trace:
    rollout(K)
    torch.cuda.synchronize()
    env.check()
    sys.exit(0)
CAP = 4 * B + 1024
trace = torch.zeros(1 + 8 * CAP, dtype=torch.int64, device=dev)
lib.mapf_diag_rollout_trace(C.c_void_p(trace.data_ptr()), CAP)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
torch.cuda._sleep(400_000)
ev0.record()
rollout(K)
ev1.record()
torch.cuda.synchronize()
lib.mapf_diag_rollout_trace(None, 0)
env.check()
t = trace.cpu().numpy().astype(np.uint64)
n = int(t[0])
rec = t[1:1 + 8 * n].reshape(n, 8)
item, env_t0, claimed, loaded, regen_ns, regens, end, where = (rec[:, i] for i in range(8))
e = (env_t0 & np.uint64(0xffffffff)).astype(np.int64)
t0s = (env_t0 >> np.uint64(32)).astype(np.int64)
warp = (where >> np.uint64(32)).astype(np.int64)
sm = (where & np.uint64(0xffffffff)).astype(np.int64)
start = int(claimed.min())
cl = (claimed - np.uint64(start)).astype(np.float64) / 1e3
ld = (loaded - np.uint64(start)).astype(np.float64) / 1e3
en = (end - np.uint64(start)).astype(np.float64) / 1e3
rg = (regen_ns & np.uint64(0xffffffff)).astype(np.float64) / 1e3
gen = (regen_ns >> np.uint64(32)).astype(np.float64) / 1e3   # the generator's share of a re-generation
n_regen = (regens & np.uint64(0xffff)).astype(np.int64)
n_adopt = (regens >> np.uint64(16)).astype(np.int64)
dur = en - cl
span = float(en.max())
nw = len(np.unique(warp))
out = {"config": args.config, "K": K, "pregen": args.pregen, "event_us": round(ev0.elapsed_time(ev1) * 1e3, 1),
       "kernel_span_us": round(span, 1), "items": n, "warps": nw, "sms": len(np.unique(sm)),
       "regenerated_in_kernel": int(n_regen.sum()), "adopted": int(n_adopt.sum()),
       "item_us": {"plain_median": round(float(np.median(dur[(n_regen + n_adopt) == 0])), 1) if ((n_regen + n_adopt) == 0).any() else None,
                   "with_regen_median": round(float(np.median(dur[n_regen > 0])), 1) if (n_regen > 0).any() else None,
                   "with_adopt_median": round(float(np.median(dur[n_adopt > 0])), 1) if (n_adopt > 0).any() else None},
       "regen_us": {"median": round(float(np.median(rg[n_regen > 0])), 1), "p90": round(float(np.percentile(rg[n_regen > 0], 90)), 1),
                    "max": round(float(rg[n_regen > 0].max()), 1),
                    "generator_median": round(float(np.median(gen[n_regen > 0])), 1)} if (n_regen > 0).any() else None,
       "adopt_us": {"median": round(float(np.median(rg[n_adopt > 0])), 1), "max": round(float(rg[n_adopt > 0].max()), 1)} if (n_adopt > 0).any() else None,
       "wait_plus_load_us_median": round(float(np.median(ld - cl)), 2), "wait_plus_load_us_p99": round(float(np.percentile(ld - cl, 99)), 1),
       "first_claim_spread_us": None, "warp_busy_frac": round(float(dur.sum() / (nw * span)), 3),
       "regen_share_of_warp_time": round(float(rg.sum() / (nw * span)), 3)}
# when each warp claimed its first item / finished its last one
first = {}
last = {}
for w, c, x in zip(warp, cl, en):
    first[w] = min(first.get(w, 1e18), c)
    last[w] = max(last.get(w, 0.0), x)
fv, lv = np.array(list(first.values())), np.array(list(last.values()))
out["first_claim_spread_us"] = [round(float(np.percentile(fv, q)), 1) for q in (0, 50, 99, 100)]
out["last_done_us_percentiles"] = [round(float(np.percentile(lv, q)), 1) for q in (0, 10, 50, 90, 100)]
# items in flight / regenerations in flight over time (20 bins)
bins = np.linspace(0, span, 21)
out["in_flight"] = [int(((cl <= b) & (en > b)).sum()) for b in bins[:-1]]
print(json.dumps(out))
if args.dump:
    np.save(args.dump, rec)
