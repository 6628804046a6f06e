#!/usr/bin/env python
"""Time of search.create_test (test.py:23-79): N solvable instances minted per (agents, map side): instances + BFS distance maps
on the GPU, CBS on the host threads.    python profiles/tools/r2_cbs_generator.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from mapf_rl_b200 import search  # noqa: E402

for agents, side, n in ((4, 10, 200), (16, 40, 200), (32, 40, 200), (64, 40, 50)):
    t0 = time.perf_counter()
    tests = search.create_test(agents, side, test_num=n, density=0.3, seed=1, time_limit_s=5.0, node_limit=1 << 15)
    dt = time.perf_counter() - t0
    print(json.dumps({"agents": agents, "map": side, "instances": n, "seconds": round(dt, 2), "opt_mean_steps": round(tests["opt_mean_steps"], 2),
                      "host_threads": os.cpu_count(), "note": "density 0.3; CBS bounded by 5 s / 32768 high-level nodes per instance, unsolved ones replaced"}),
          flush=True)
