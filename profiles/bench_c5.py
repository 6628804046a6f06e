"""bench.py --config c5 — BASELINE.json configs[4]: the end-to-end DQN actor loop on device.

Per GPU: 2048 envs x 32 agents (40x40 / 0.3) stepped by `BatchedActor` (batched PyTorch Q-net forward in bf16 -> epsilon on
agent 0 -> fused step+observe kernel writing into the replay store -> actor-TD priorities -> sum-tree insert -> masked
device reset at episode ends), interleaved with `BatchedLearner` updates (window gather kernel -> two PyTorch bootstrap
passes -> loss / Adam -> ONE mapf_per_cycle launch: priorities in, next batch out).  With N > 1 ranks the learner's
gradients are averaged with one NCCL all-reduce per update; the env / replay path has no collective.
`value` = agent-steps/s of the whole loop (actor steps, learner updates included in the time).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run(args, METRIC, UNIT, ClockSampler, measured_hbm_peak, workload_name, bench_config):
    import torch
    import torch.distributed as dist
    from mapf_rl_b200 import BatchedEnvironment, ReplayStore, config, sharding
    from mapf_rl_b200.actor import BatchedActor
    from mapf_rl_b200.learner import BatchedLearner
    from mapf_rl_b200.qnet import Network

    rank, world, local = sharding.rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sharding.pin_to_cores(local, world)
    B, N, L = args.num_envs, args.num_agents, args.map_length
    cap = min(args.max_steps, 64)            # episode cap of the bench run: episodes must finish (and be published) within it
    K, W = max(args.steps, 3 * cap), max(args.warmup, 3)
    torch.manual_seed(args.seed)             # the same initial weights on every rank
    env = BatchedEnvironment(B, N, L, device=dev)
    net = Network().to(dev).to(memory_format=torch.channels_last)
    slots = 1
    while slots < 2 * B:
        slots *= 2
    store = ReplayStore(slots, max_num_agents=N, device=dev, max_steps=cap)
    eps = 0.4 ** (1 + 7 * torch.arange(B, device=dev, dtype=torch.float32) / max(B - 1, 1))   # train.py:25 over the env batch
    actor = BatchedActor(env, net, store, epsilon=eps, seed=args.seed + rank, density=args.density, max_steps=cap)
    learner = BatchedLearner(net, store, allreduce=world > 1, seed=args.seed + rank)
    min_transitions = 4 * config.batch_size

    class _AutocastStep:
        """actor forward under bf16 autocast (qnet.py header: 15.7 ms instead of 38 ms for 2048 x 32 agents)"""
        def __init__(self, net):
            self.net = net

        def step(self, *a, **k):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.net.step(*a, **k)

        def reset(self):
            self.net.reset()

    actor.net = _AutocastStep(net)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    updates = 0

    def loop(steps, timed):
        nonlocal updates
        for s in range(steps):
            actor.step()
            if (s + 1) % args.learner_every == 0 and learner.ready(min_transitions):
                learner.update()
                updates += timed

    loop(W + cap + 2, False)       # warm-up: at least one full episode per env is published, the learner has run
    for _ in range(2):
        if learner.ready(min_transitions):
            learner.update()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ep0, tr0 = actor.episodes, actor.transitions
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    loop(K, True)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms_all = sharding.gather_floats(ev0.elapsed_time(ev1), dev)
    ms = max(ms_all)
    clocks = sampler.stop()
    env.check()
    store.priority_tree.check()
    value = world * B * N * K / (ms * 1e-3)

    # ---- the pieces, timed alone on rank 0's GPU (CUDA events, median of 5) ---------------------------------------------------
    def timed(fn, reps=5, inner=1):
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b) * 1e3 / inner)
        return sorted(ts)[len(ts) // 2]

    tree = store.priority_tree
    u = torch.rand(config.batch_size, dtype=torch.float64, device=dev)
    idx, _, _ = tree.sample_device(config.batch_size, u)
    rng = np.random.default_rng(0)
    n = config.batch_size
    upd = dict(q_online=torch.randn(n, 5, device=dev), q_target_next=torch.randn(n, 5, device=dev),
               action=torch.randint(0, 5, (n,), device=dev), reward=torch.zeros(n, device=dev), done=torch.zeros(n, device=dev),
               steps=torch.ones(n, device=dev), idx=idx)
    outbuf = {}
    pieces = {
        "per_cycle_us (TD -> priorities -> tree update of 192 + next 192 samples + IS weights, ONE launch)":
            timed(lambda: outbuf.update(tree.cycle(update=upd, sample_size=n, uniforms=u, beta=0.4, out=outbuf)), inner=20),
        "per_td_update_us + per_sample_us (the same as two launches)":
            timed(lambda: (tree.td_update(upd["q_online"], upd["q_target_next"], upd["action"], upd["reward"], upd["done"],
                                          upd["steps"], upd["idx"]), tree.sample_device(n, u, beta=0.4)), inner=20),
        "replay_gather_us (192 x 18 frames x 32 agents, bool -> fp16)": timed(lambda: store.gather(idx), inner=5),
        "learner_update_us (gather + 2 bootstraps + backward + Adam + cycle)": timed(lambda: learner.update(), inner=2),
        "actor_step_us (Q-net forward + env step + bookkeeping)": timed(lambda: actor.step(), inner=4),
    }
    obs_t = torch.empty((B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    a8 = torch.randint(0, 5, (B, N), device=dev, dtype=torch.uint8)
    pieces["env_step_observe_kernel_us (the hand-written part of an actor step)"] = timed(lambda: env.step(a8, out_obs=obs_t), inner=20)
    pieces["comm_mask_kernel_us"] = timed(lambda: env.comm_mask(), inner=20)
    stats = learner.update(want_stats=True)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (env) / bf16 (Q-net) / f64 (sum tree)",
            "data": "synthetic",
            "config": bench_config(args),
            "method": {"learner_every_actor_steps": args.learner_every, "replay_slots": slots, "batch_size": config.batch_size,
                       "note": "Q-network GEMMs / convolutions stay in PyTorch (north star); the hand-written "
                               "kernels are the env step, comm mask, replay gather, actor TD and the PER cycle"},
            "clocks": clocks, "per_rank_ms": {"min": min(ms_all), "max": max(ms_all), "all": ms_all},
            "actor": {"episodes_published": actor.episodes - ep0, "transitions": actor.transitions - tr0, "resets": actor.resets},
            "learner": {"updates_in_timed_region": updates, "updates_total": learner.counter, "stats_last_update": stats,
                        "gradient_allreduce": "one NCCL all-reduce of the flattened fp32 gradients per update" if world > 1 else "none (1 rank)"},
            "pieces_us": pieces,
            "wall_s": wall,
            "gpu_launches": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
