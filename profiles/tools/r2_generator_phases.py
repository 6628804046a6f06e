#!/usr/bin/env python
"""Cycles per phase of (1) the generator (reset_env_warp, in reset_kernel) and (2) the heuristic-map searches of an in-launch
re-generation (inside rollout_kernel) for ONE environment running alone on its SM -- the latency an in-launch re-generation
sees (diagnosis build):  MAPF_B200_LIB=mapf_rl_b200/libmapf_b200_diag.so python profiles/tools/r2_generator_phases.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import BatchedEnvironment, _native  # noqa: E402

lib = _native.lib()
names = ["map (Philox draws + inserts)", "eligibility (dilate)", "agent draws", "first pick", "component (cache test / flood fill)",
         "second pick + bookkeeping", "stores", "bitmap store"]
for (B, N, L) in ((148, 32, 40), (148, 64, 40), (148, 64, 80)):
    env = BatchedEnvironment(B, N, L)
    tot = [0] * 8
    R = 8
    for rep in range(R):
        env.reset(seed=rep, env_offset=0, density=0.3)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)()
        lib.mapf_diag_reset_cycles(out)
        tot = [a + b for a, b in zip(tot, out)]
    rec = {"B": B, "N": N, "L": L, "generator_cycles_env0_mean": {n: round(t / R) for n, t in zip(names, tot)},
           "generator_us_at_1965MHz": round(sum(tot) / R / 1965, 1)}
    # one rollout step in which every environment re-generates (cap reached): the searches of environment 0
    cap = 8
    env.set_autoreset(cap, seed=0, env_offset=B, stride=B, density=0.3)
    acts = torch.zeros((1, B, N), dtype=torch.uint8, device="cuda")
    env.set_state(steps=torch.full((B,), cap, dtype=torch.int32, device="cuda"))
    env.rollout(acts, num_steps=1)
    torch.cuda.synchronize()
    lib.mapf_diag_bfs_cycles_occ8(None, 1)
    env.set_state(steps=torch.full((B,), cap, dtype=torch.int32, device="cuda"))
    env.rollout(acts, num_steps=1)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 8)()
    lib.mapf_diag_bfs_cycles_occ8(out, 0)
    n = max(int(out[3]), 1)
    rec["search_in_rollout_env0"] = {"searches": int(out[3]), "wave_triples_per_search": round(out[4] / n, 1),
                                     "cycles_per_search": {"init": round(out[0] / n), "waves": round(out[1] / n), "emit": round(out[2] / n)},
                                     "cycles_per_wave": round(out[1] / max(int(out[4]), 1) / 3, 1)}
    env.check()
    print(json.dumps(rec), flush=True)
    env.close()
