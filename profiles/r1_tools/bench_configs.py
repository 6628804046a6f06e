#!/usr/bin/env python
"""Secondary measurements (GPU box, one GPU): the BASELINE.json configs that are not the bench.py line.

    python profiles/bench_configs.py [--out gpurun_out/configs.jsonl]

  C2-G   40x40 / 32 agents / 8192 envs, navi-greedy (eps 0.1) action stream — congestion-heavy conflicts
  C3     40x40 / 64 agents / 8192 envs (two agents per lane), uniform and greedy streams
  C4     80x80 / 64 agents, reset-heavy: device generator + per-agent BFS heuristic maps, and stepping
  K3     BFS heuristic maps per second at both geometries (load path)
  K4     PER sum tree at the reference's 2^19 leaves: batch_sample(192), fused TD+priority update(192),
         episode insert batch_update(256), actor TD for 64 episodes
Every number is CUDA-event time over back-to-back launches after warm-up; one JSON object per line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ev_time(torch, fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters  # us


def greedy_actions_torch(torch, obs, gen, eps=0.1):
    """navi-greedy stream G on the device: uniform among the set heuristic bits of the centre cell, else stay."""
    bits = obs[:, :, 2:6, 4, 4].to(torch.float32)                                  # [B,N,4]
    score = torch.rand(bits.shape, device=obs.device, generator=gen) * bits
    act = torch.where(bits.sum(-1) > 0, 1 + score.argmax(-1), torch.zeros_like(score.argmax(-1)))
    explore = torch.rand(act.shape, device=obs.device, generator=gen) < eps
    rnd = torch.randint(0, 5, act.shape, device=obs.device, generator=gen)
    return torch.where(explore, rnd, act).to(torch.uint8)


def env_config(torch, out, name, B, N, L, steps=200):
    from bench import algo_bytes, measured_hbm_peak
    from mapf_rl_b200 import BatchedEnvironment
    dev = torch.device("cuda", 0)
    env = BatchedEnvironment(B, N, L, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    t_reset = ev_time(torch, lambda: env.reset(seed=1, density=0.3), 3, warm=1)
    env.check()
    ids = torch.arange(B, dtype=torch.int32, device=dev)
    import ctypes as C
    from mapf_rl_b200 import _native
    t_bfs = ev_time(torch, lambda: _native.check(env._lib.mapf_env_bfs_navi(env._h, None, B, None, env._stream())), 3, warm=1)
    R = 4
    replay = torch.empty((R, B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
    peak, _ = measured_hbm_peak()
    for stream in ("U", "G"):
        env.reset(seed=2, density=0.3)
        if stream == "U":
            acts = torch.randint(0, 5, (16, B, N), generator=gen, device=dev, dtype=torch.uint8)
        else:
            # record a greedy stream by actually following it (actions depend on observations), then replay it
            obs, _ = env.observe()
            rec = []
            for s in range(16):
                a = greedy_actions_torch(torch, obs, gen)
                rec.append(a)
                obs, _, _ = env.step(a)
            acts = torch.stack(rec)
            env.reset(seed=2, density=0.3)
        k = [0]

        def one():
            env.step(acts[k[0] % 16], out_obs=replay[k[0] % R])
            k[0] += 1
        us = ev_time(torch, one, steps, warm=32)
        rew = env._rewards
        coll = float((rew == -0.5).float().mean().item())
        env.check()
        ach = algo_bytes(N, L) * B * N / (us * 1e-6) / 1e9
        # the same stream through mapf_env_rollout (independent sub-batch chains, graph replay)
        T = 640
        rr = torch.empty((2, B, N), dtype=torch.float32, device=dev)
        rd = torch.empty((2, B), dtype=torch.uint8, device=dev)
        rs = torch.empty((2, B), dtype=torch.int32, device=dev)
        roll = lambda: env.rollout(acts, num_steps=T, out_obs=replay, out_rewards=rr, out_done=rd, out_steps=rs)
        us_roll = ev_time(torch, roll, 2, warm=1) / T
        plan = env.rollout_plan(T, 16, R, 2)
        env.check()
        out({"config": name, "stream": stream, "num_envs": B, "num_agents": N, "map_length": L, "us_per_step": round(us, 2),
             "agent_steps_per_s": B * N / (us * 1e-6), "roofline_frac_of_measured_hbm": ach / peak,
             "us_per_step_rollout": round(us_roll, 2), "rollout_chains": plan[0], "rollout_agent_steps_per_s": B * N / (us_roll * 1e-6),
             "rollout_roofline_frac": algo_bytes(N, L) * B * N / (us_roll * 1e-6) / 1e9 / peak,
             "collision_fraction_last_step": coll})
    out({"config": name, "what": "reset = device generator + BFS heuristic maps, all envs", "num_envs": B, "num_agents": N,
         "map_length": L, "us_per_reset_batch": round(t_reset, 1), "envs_reset_per_s": B / (t_reset * 1e-6),
         "us_bfs_batch": round(t_bfs, 1), "heuristic_maps_per_s": B * N / (t_bfs * 1e-6),
         "navi_bytes_written_per_s": B * N * ((L + 7) // 8) ** 2 * 128 / (t_bfs * 1e-6)})
    env.close()
    del replay
    torch.cuda.empty_cache()


def per_config(torch, out):
    from mapf_rl_b200 import SumTree
    from mapf_rl_b200.buffer import actor_td_errors
    dev = torch.device("cuda", 0)
    cap = 2048 * 256  # train.py:21, config.py:29
    tree = SumTree(cap, device=dev)
    rng = np.random.default_rng(0)
    # fill: 2048 episode inserts of 256 leaves (worker.py:87-94)
    pr_all = torch.as_tensor(rng.random(cap) ** 0.6, dtype=torch.float64, device=dev)
    idx_all = torch.arange(cap, dtype=torch.int64, device=dev)
    for s in range(0, cap, 4096):
        tree.update_device(idx_all[s:s + 4096], pr_all[s:s + 4096])
    u = torch.rand(192, dtype=torch.float64, device=dev)
    t_sample = ev_time(torch, lambda: tree.sample_device(192, u, beta=0.4), 200)
    idx, pr, w = tree.sample_device(192, u, beta=0.4)
    q1 = torch.randn(192, 5, device=dev)
    q2 = torch.randn(192, 5, device=dev)
    a = torch.randint(0, 5, (192,), device=dev)
    r = torch.randn(192, device=dev)
    d = torch.zeros(192, device=dev)
    st = torch.full((192,), 2.0, device=dev)
    import ctypes as C
    from mapf_rl_b200 import _native
    td = torch.empty(192, device=dev)
    po = torch.empty(192, device=dev)

    def td_upd():
        _native.check(tree._lib.mapf_per_td_update(tree._h, q1.data_ptr(), q2.data_ptr(), None, a.data_ptr(), r.data_ptr(),
                                                   d.data_ptr(), st.data_ptr(), idx.data_ptr(), 192, 0.99, 0.6, 0, 0, 256,
                                                   td.data_ptr(), po.data_ptr(), tree._stream()))
    t_td = ev_time(torch, td_upd, 200)
    ep_idx = torch.arange(256, dtype=torch.int64, device=dev) + 256 * 7
    ep_pr = torch.rand(256, dtype=torch.float64, device=dev)

    def ins():
        _native.check(tree._lib.mapf_per_update(tree._h, ep_idx.data_ptr(), ep_pr.data_ptr(), 256, tree._stream()))
    t_ins = ev_time(torch, ins, 200)
    E = 64
    rew = torch.randn(E, 256, device=dev)
    q = torch.randn(E, 256, 5, device=dev)
    act = torch.randint(0, 5, (E, 256), device=dev, dtype=torch.uint8)
    size = torch.full((E,), 256, dtype=torch.int32, device=dev)
    t_actor = ev_time(torch, lambda: actor_td_errors(rew, q, act, size), 100)
    out({"config": "K4 PER sum tree, 2^19 leaves (reference capacity)", "batch_sample_192_us": round(t_sample, 2),
         "fused_td_priority_update_192_us": round(t_td, 2), "episode_insert_256_us": round(t_ins, 2),
         "actor_td_64_episodes_us": round(t_actor, 2),
         "reference_cpu_us": {"batch_sample_192": 348, "batch_update_192": 572, "batch_update_256": 396,
                              "finish_256_steps_per_episode": 168, "source": "SURVEY.md section 6 (measured, 1 core)"}})


def glue_config(torch, out):
    """Widened rows: comm-mask kernel, replay window gather (K5), and the whole actor loop (configs[4])."""
    from bench import measured_hbm_peak
    from mapf_rl_b200 import BatchedEnvironment, ReplayStore
    from mapf_rl_b200.actor import BatchedActor
    from mapf_rl_b200.qnet import Network
    from replay_cases import make_episode
    dev = torch.device("cuda", 0)
    peak, _ = measured_hbm_peak()
    # comm mask at the bench geometry
    env = BatchedEnvironment(8192, 32, 40, device=dev)
    env.reset(seed=0, density=0.3)
    m = torch.empty((8192, 32, 32), dtype=torch.uint8, device=dev)
    us = ev_time(torch, lambda: env.comm_mask(out=m), 200)
    out({"config": "comm mask (model.py:196-208), 8192 envs x 32 agents", "us_per_call": round(us, 2),
         "agent_rows_per_s": 8192 * 32 / (us * 1e-6), "bytes_written": 8192 * 32 * 32})
    env.close()
    # K5 window gather: reference geometry (6 agents, batch 192) and the bench geometry (32 agents)
    rng = np.random.default_rng(0)
    for n, cap in ((6, 64), (32, 64)):
        store = ReplayStore(cap, max_num_agents=n, device=dev)
        store.add([make_episode(rng, k, n, 256, False) for k in range(cap)])
        u = torch.rand(192, dtype=torch.float64, device=dev)
        idx, _, _ = store.priority_tree.sample_device(192, u)
        us = ev_time(torch, lambda: store.gather(idx), 100)
        # the torch allocations of gather() are included; bytes: up to 18 frames read as bool, written as fp16
        W = store.bt_steps + store.forward_steps
        rd = 192 * W * n * (486 + n) + 192 * n * 256 * 2
        wr = 192 * W * n * (972 + n) + 192 * n * 256 * 2
        out({"config": f"K5 replay window gather, batch 192, {n} agents, {W} frames", "us_per_call": round(us, 2),
             "algorithmic_bytes": rd + wr, "achieved_GBs": (rd + wr) / (us * 1e-6) / 1e9,
             "roofline_frac_of_measured_hbm": (rd + wr) / (us * 1e-6) / 1e9 / peak})
        del store
        torch.cuda.empty_cache()
    # configs[4]: the actor loop, env batch + Q-net (bf16 autocast) + replay recording
    B, N, L = 2048, 32, 40
    env = BatchedEnvironment(B, N, L, device=dev)
    net = Network().to(dev).eval().to(memory_format=torch.channels_last)   # NHWC convs: 24.6 -> 15.7 ms (qnet_forward_probe.py)
    store = ReplayStore(2 * B, max_num_agents=N, device=dev)
    actor = BatchedActor(env, net, store, epsilon=0.1, seed=0, density=0.3)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        actor.run(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 20
        actor.run(steps)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        obs = torch.zeros((B, N, 6, 9, 9), dtype=torch.uint8, device=dev)
        comm = env.comm_mask()
        t_net = ev_time(torch, lambda: net.step(obs, comm), 10, warm=2)
    t_env = ev_time(torch, lambda: env.step(torch.zeros((B, N), dtype=torch.uint8, device=dev)), 50)
    out({"config": "C5 actor loop: 2048 envs x 32 agents 40x40, Q-net bf16 autocast, replay recording, auto reset",
         "ms_per_step": round(el / steps * 1e3, 3), "agent_steps_per_s": B * N * steps / el,
         "qnet_forward_ms": round(t_net / 1e3, 3), "env_step_observe_us": round(t_env, 2),
         "episodes_published": actor.episodes, "store_GB": sum(t.numel() * t.element_size() for t in
                                                                 (store.obs_buf, store.hid_buf, store.comm_mask)) / 1e9})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    f = open(args.out, "w")

    def out(obj):
        obj["gpu"] = torch.cuda.get_device_name(0)
        line = json.dumps(obj)
        print(line, flush=True)
        f.write(line + "\n")
        f.flush()

    want = set(args.only.split(",")) if args.only else None
    if not want or "C2" in want:
        env_config(torch, out, "C2 40x40/0.3, 32 agents, 8192 envs", 8192, 32, 40)
    if not want or "C3" in want:
        env_config(torch, out, "C3 40x40/0.3, 64 agents, 8192 envs", 8192, 64, 40)
    if not want or "C4" in want:
        env_config(torch, out, "C4 80x80/0.3, 64 agents, 4096 envs", 4096, 64, 80, steps=100)
    if want and "C4B" in want:
        # same geometry at batch sizes that fill whole waves of the 20 resident warps per SM (2960 per wave)
        env_config(torch, out, "C4 80x80/0.3, 64 agents, 5920 envs (2 full waves)", 5920, 64, 80, steps=100)
        env_config(torch, out, "C4 80x80/0.3, 64 agents, 8192 envs", 8192, 64, 80, steps=100)
    if not want or "K4" in want:
        per_config(torch, out)
    if not want or "GLUE" in want:
        glue_config(torch, out)
    f.close()


if __name__ == "__main__":
    main()
