#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE — times the reference's OWN Python `Environment.step` (which ends in
`observe`, environment.py:430) on the host cores, one process per core, as BASELINE.json's north star asks
("reported next to the reference's Python Environment.step+observe timed on the same box's host cores").

The reference is loaded live through oracle/ref_loader.py (three shims, nothing else altered), so this runs only
where the reference is mounted (`/root/reference` in the dev container, or `MAPF_REFERENCE_DIR`).  The GPU box
has no reference tree: bench.py calls `measure()` when `ref_loader.available()` and otherwise quotes the
committed result of this script run in the dev container (profiles/r2_python_reference_cpu.json, core count
stated there).

    python oracle/time_python_reference.py [--agents 32] [--seconds 8] [--procs N] [--out FILE]

Workload per process: `load()` pkl instances round-robin (untimed — it is the 0.3 s BFS), then `step(actions)`
with uniform-random actions, 64 steps per instance, until the time budget is spent.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(args):
    rank, num_agents, seconds, steps_per_instance = args
    os.environ["OMP_NUM_THREADS"] = "1"   # train.py:2
    import numpy as np
    from oracle import ref_loader
    envmod = ref_loader.load_environment()
    maps, agents, goals = ref_loader.load_pkl(num_agents)
    rng = np.random.default_rng(1000 + rank)
    env = envmod.Environment()
    timed, agent_steps, k = 0.0, 0, rank
    while timed < seconds:
        k = (k + 1) % len(maps)
        env.load(maps[k], agents[k], goals[k])                       # untimed
        acts = rng.integers(0, 5, size=(steps_per_instance, num_agents))
        t0 = time.perf_counter()
        for s in range(steps_per_instance):
            env.step(acts[s].tolist())                               # step + observe
        timed += time.perf_counter() - t0
        agent_steps += steps_per_instance * num_agents
    return agent_steps, timed


def measure(num_agents: int = 32, seconds: float = 8.0, procs: int | None = None, steps_per_instance: int = 64):
    """-> dict(value agent-steps/s aggregate, per_core, cores, ...).  Raises if the reference is not mounted."""
    from oracle import ref_loader
    if not ref_loader.available():
        raise RuntimeError("reference not mounted")
    procs = procs or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(r, num_agents, seconds, steps_per_instance) for r in range(procs)])
    per_core = [a / t for a, t in res]
    return {"value": float(sum(per_core)), "unit": "agent-steps/s", "cores": procs, "kind": "reference",
            "per_core_mean": float(sum(per_core) / len(per_core)), "per_core_min": float(min(per_core)),
            "per_core_max": float(max(per_core)),
            "sample": f"live reference Environment.step (+observe) on test{num_agents}_40_0.3.pkl instances, uniform actions, "
                      f"{steps_per_instance} steps per instance, one process per core x {procs}, ~{seconds:.0f} s each, load() untimed",
            "python": sys.version.split()[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import platform
    r = measure(a.agents, a.seconds, a.procs or None)
    import numpy as np
    r["numpy"] = np.__version__
    r["where"] = f"dev container ({platform.processor() or platform.machine()}, {os.cpu_count()} vCPU)"
    try:
        with open("/proc/cpuinfo") as f:
            names = [ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")]
        if names:
            r["cpu_model"] = names[0]
    except OSError:
        pass
    print(json.dumps(r))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(r, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
