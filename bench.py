#!/usr/bin/env python
"""bench.py — env agent-steps/s of the fused step+observe hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one lockstep Environment.step (+ the observe it ends in) over the whole batch of
synthetic environments: configs[1] of BASELINE.json — 8192 envs x 32 agents on 40x40 maps with
obstacle density 0.3 (iid Bernoulli), uniform-random actions.  Multi-GPU: env batches shard with no
collective (weak scaling, 8192 envs per GPU); torch.distributed is used only for the barrier and the
max-over-ranks of the device time.

Printed JSON (rank 0, one line): see the field notes in DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_agent_steps_per_sec_step_observe"
UNIT = "agent-steps/s"
ALGO_BYTES_PER_AGENT_STEP = {  # 486 obs + 40.5 navi crop + 1 action + 4 pos r/w + 2 goal + 4 reward + map/done/steps amortised
    (32, 40): 544.0, (64, 40): 541.0, (64, 80): 550.0, (16, 40): 550.4,
}


def algo_bytes(N, L):
    if (N, L) in ALGO_BYTES_PER_AGENT_STEP:
        return ALGO_BYTES_PER_AGENT_STEP[(N, L)]
    return 486 + 40.5 + 1 + 4 + 2 + 4 + (L * L / 8.0 + 5) / N


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_instances(num, L, N, density, seed, first_index):
    """Synthetic batch following the reference generator (environment.py:100-138) at fixed density."""
    from concurrent.futures import ProcessPoolExecutor
    from mapf_rl_b200.instances import generate_batch
    workers = max(1, min(os.cpu_count() or 1, 32))
    chunk = (num + workers - 1) // workers
    jobs = [(min(chunk, num - s), L, N, density, seed, first_index + s) for s in range(0, num, chunk)]
    if workers == 1:
        parts = [generate_batch(*j) for j in jobs]
    else:
        with ProcessPoolExecutor(workers) as ex:
            parts = list(ex.map(generate_batch, *zip(*jobs)))
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML, in-process)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def run():
            while not self._stop.is_set():
                self._sample()
                time.sleep(0.002)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._sample()
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(maps, agents, goals, N, L, seconds=12.0, threads=None, max_envs=2048):
    """The oracle port (C restatement of environment.py:278-467) on the host cores, bounded sample."""
    from oracle import oracle
    threads = threads or (os.cpu_count() or 1)
    S = min(max_envs, maps.shape[0])
    m = np.ascontiguousarray(maps[:S])
    pos = np.ascontiguousarray(agents[:S].astype(np.int32))
    gl = np.ascontiguousarray(goals[:S].astype(np.int32))
    navi = np.empty((S, N, 4, L, L), dtype=np.uint8)
    for k in range(S):
        navi[k] = oracle.navi(m[k], gl[k])[1]
    rng = np.random.default_rng(1)
    T = 8
    acts = rng.integers(0, 5, size=(T, S, N)).astype(np.uint8)
    oracle.rollout(m, pos, gl, navi, acts[:1], threads=threads, want_rewards=False)  # warm
    t0 = time.perf_counter()
    done_steps = 0
    while True:
        oracle.rollout(m, pos, gl, navi, acts, threads=threads, want_rewards=False)
        done_steps += T
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": S * N * done_steps / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{S} of the batch's envs x {done_steps} lockstep steps (uniform actions), C oracle port "
                      f"(oracle/mapf_oracle.c) on {threads} host threads, {el:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path (oracle port; the reference itself is
    pure Python and /root/reference does not exist on the GPU box), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    N, L = args.num_agents, args.map_length
    threads = os.cpu_count() or 1
    S = args.ref_envs
    maps, agents, goals = make_instances(S, L, N, args.density, args.seed, 0)
    pos = np.ascontiguousarray(agents.astype(np.int32))
    gl = np.ascontiguousarray(goals.astype(np.int32))
    navi = np.empty((S, N, 4, L, L), dtype=np.uint8)
    for k in range(S):
        navi[k] = oracle.navi(maps[k], gl[k])[1]
    rng = np.random.default_rng(1)
    acts = rng.integers(0, 5, size=(16, S, N)).astype(np.uint8)
    for w in range(max(args.warmup, 1)):
        oracle.rollout(maps, pos, gl, navi, acts[w % 16:w % 16 + 1], threads=threads, want_rewards=False)
    t0 = time.perf_counter()
    for s in range(args.steps):
        oracle.rollout(maps, pos, gl, navi, acts[s % 16:s % 16 + 1], threads=threads, want_rewards=False)
    el = time.perf_counter() - t0
    value = S * N * args.steps / el
    sample = (f"each step = one lockstep step+observe over {S} envs ({N} agents, {L}x{L}, density {args.density}), "
              f"C oracle port of environment.py:278-467 on {threads} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "num_envs_per_step": S, "num_agents": N, "map_length": L,
                       "obstacle_density": args.density, "actions": "uniform"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"batched step+observe, {args.map_length}x{args.map_length} / {args.density} density, {args.num_agents} agents, "
            f"{args.num_envs} lockstep envs per GPU (BASELINE.json configs[1])")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mapf_rl_b200 import BatchedEnvironment, sharding

    rank, world, local = sharding.rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, N, L = args.num_envs, args.num_agents, args.map_length
    env = BatchedEnvironment(B, N, L, device=dev)
    # synthetic instances drawn on the device by the reference's procedure (environment.py:100-138) at fixed
    # density; slot e of rank r is global environment r*B + e whatever the number of GPUs (weak scaling)
    env.reset(seed=args.seed, env_offset=sharding.weak_offset(B, rank), density=args.density)
    env.check()

    R = args.obs_ring  # observation ring (device replay slots): R x B*N*486 bytes > L2
    replay = torch.empty((R, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    A = 16
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed + rank)
    actions = torch.randint(0, 5, (A, B, N), generator=g, device=dev, dtype=torch.uint8)
    actions_host = actions.cpu().pin_memory()  # [A,B,N] page-locked: step_host hands slot s % A to the GPU in place
    actions_host = [actions_host[i] for i in range(A)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up: W steps, then keep stepping until the clocks have had ~0.3 s of load
    for s in range(args.warmup):
        env.step(actions[s % A], out_obs=replay[s % R])
    t0 = time.perf_counter()
    s = 0
    while time.perf_counter() - t0 < 0.3:
        for _ in range(64):
            env.step(actions[s % A], out_obs=replay[s % R])
            s += 1
        torch.cuda.synchronize(dev)
    env.check()

    # ---- timed region: exactly K steps, device-resident inputs -----------------------------------
    # mapf_env_rollout: the K steps of the whole batch as `chains` independent sub-batch chains on internal streams
    # (environments are independent; step t+1 of a sub-batch waits for step t of that sub-batch only)
    out_ring = 2
    chains, per, graph_period = env.rollout_plan(args.steps, A, R, out_ring, args.chains)
    rew_ring = torch.empty((out_ring, B, N), dtype=torch.float32, device=dev)
    done_ring = torch.empty((out_ring, B), dtype=torch.uint8, device=dev)
    steps_ring = torch.empty((out_ring, B), dtype=torch.int32, device=dev)

    def rollout(k):
        env.rollout(actions, num_steps=k, out_obs=replay, out_rewards=rew_ring, out_done=done_ring, out_steps=steps_ring,
                    chains=args.chains)

    rollout(64)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    rollout(args.steps)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = sharding.max_over_ranks(ev0.elapsed_time(ev1), dev)
    value = world * B * N * args.steps / (ms * 1e-3)
    env.check()

    # ---- the same K steps as one whole-batch launch per step (mapf_env_step_observe from Python) ----
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record()
    for s in range(args.steps):
        env.step(actions[s % A], out_obs=replay[s % R])
    ev3.record()
    barrier()
    ms_single = sharding.max_over_ranks(ev2.elapsed_time(ev3), dev)

    # ---- e2e: through the host-buffer C-ABI entry point (mapf_env_step_host) ---------------------
    replay_slots = [replay[i] for i in range(R)]

    def e2e_run(steps, want_obs):
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            env.step_host(actions_host[s % A], want_obs=want_obs, device_obs=replay_slots[s % R])
        torch.cuda.synchronize(dev)
        el = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        return world * B * N * steps / el

    e2e_steps = max(10, min(args.steps, 400))
    e2e_run(2 * A, False)   # warm-up: every (action buffer, observation slot) pair has its captured launch sequence
    # the host side of this path (Python, graph launch, stream sync) is sensitive to what else runs on the box: three
    # repetitions, the median is reported and all three are kept in the JSON line
    e2e_reps = sorted(e2e_run(e2e_steps, False) for _ in range(3))
    e2e_value = e2e_reps[1]
    e2e_run(1, True)
    e2e_obs_value = e2e_run(max(3, min(args.steps, 20)), True)

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        per_launch_s = ms * 1e-3 / args.steps
        achieved = algo_bytes(N, L) * B * N / per_launch_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "step_observe_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "num_envs_per_gpu": B, "num_agents": N, "map_length": L,
                       "obstacle_density": args.density, "actions": "uniform iid {0..4}, 16 pre-generated device tensors",
                       "instances": "device-side generator (mapf_env_reset), global env index = rank*B + slot",
                       "l2": f"no flush: per-step output {B * N * 486 / 1e6:.0f} MB rotates over a {R}-slot device ring "
                             f"({R * B * N * 486 / 1e6:.0f} MB) + {env.arena_bytes / 1e6:.0f} MB state arena, both > 126 MB L2",
                       "extra_warmup": "0.3 s of untimed steps after W so SM clocks are under load"},
            "clocks": clocks,
            "gpu_launches": args.steps * chains if chains else 1,
            "launch": ({"api": "mapf_env_rollout (one call for the K steps)", "chains": chains, "envs_per_chain": per,
                        "graph_period_steps": graph_period,
                        "note": "each step of the batch = `chains` launches of step_observe_kernel over disjoint env ranges on "
                                "internal streams; chains run out of phase, so one's stores overlap another's conflict resolution"}
                       if chains else
                       {"api": "mapf_env_rollout (one call for the K steps)", "chains": 0, "kernel": "step_rollout_kernel (persistent)",
                        "note": "ONE launch for the K steps: every resident warp takes its environments (4 each at this size) through "
                                "all K steps, one environment after another; warps drift out of phase on their own and an "
                                "environment's heuristic lines are re-read from L1 / L2"}),
            "single_launch": {"api": "mapf_env_step_observe, one whole-batch launch per step", "ms_per_step": ms_single / args.steps,
                              "value": world * B * N * args.steps / (ms_single * 1e-3),
                              "roofline_frac": algo_bytes(N, L) * B * N / (ms_single * 1e-3 / args.steps) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * N, "d2h_bytes_per_step": B * N * 4 + B * 5,
                    "api": "mapf_env_step_host on one of 16 page-locked action buffers (read in place over PCIe) -> step kernel -> observe kernel || "
                           "D2H rewards/done/steps on a side stream -> sync (one CUDA-graph launch); observations stay in "
                           "the device replay ring (north star)", "steps": e2e_steps, "repetitions": e2e_reps,
                    "host_mode": os.environ.get("MAPF_STEP_HOST_MODE", "4")},
            "e2e_host_obs": {"value": e2e_obs_value, "unit": UNIT, "h2d_bytes_per_step": B * N,
                             "d2h_bytes_per_step": B * N * 4 + B * 5 + B * N * 486,
                             "api": "same call with the full observation tensor also copied to host (drop-in Environment.step)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": ("step_rollout_kernel<RW=2,K=1,2 warps,64 regs> (persistent)" if not chains
                                    else "step_observe_kernel<RW=2,K=1,DO_STEP,8 warps,64 regs>"),
                         "per_launch": ({"algorithmic_bytes": algo_bytes(N, L) * per * N, "avg_duration_us": per_launch_s * 1e6,
                                         "concurrent_launches": chains,
                                         "note": "each chain's K launches run back to back for the whole timed region, so a launch "
                                                 "lasts one step period while sharing the GPU with the other chains' launches"}
                                        if chains else
                                        {"algorithmic_bytes": algo_bytes(N, L) * B * N * args.steps, "avg_duration_us": ms * 1e3,
                                         "concurrent_launches": 1, "note": "the one persistent launch covers all K steps"}),
                         "achieved_is": "algorithmic bytes of one whole-batch step / step period in the timed region (steps of "
                                        "different environments overlap inside the rollout); `traffic` is the ncu DRAM bytes of a "
                                        "whole-batch single-step launch",
                         "algorithmic_bytes_per_agent_step": algo_bytes(N, L), "peak_source": peak_src,
                         "note": "peak = measured copy (read+write) bandwidth; the persistent rollout's traffic is almost write-only "
                                 "(2 MB of DRAM reads per step, profiles/r1_rollout_persistent.log) and a pure write stream reaches "
                                 "7.2 TB/s on a B200, so frac can exceed 1; plain write / mixed streams: profiles/r1_membw_probe.jsonl"},
        }
        if not args.no_cpu_baseline and world == 1:
            S = min(2048, B)
            maps = env.map[:S].cpu().numpy()
            agents = env.agents_pos[:S].cpu().numpy()
            goals = env.goals_pos[:S].cpu().numpy()
            line["cpu_baseline"] = cpu_baseline(maps, agents, goals, N, L, seconds=args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=8192)
    ap.add_argument("--num-agents", type=int, default=32)
    ap.add_argument("--map-length", type=int, default=40)
    ap.add_argument("--density", type=float, default=0.3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--obs-ring", type=int, default=4)
    ap.add_argument("--chains", type=int, default=0, help="sub-batch chains of mapf_env_rollout (0 = library default)")
    ap.add_argument("--ref-envs", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
