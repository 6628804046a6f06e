"""Generate the committed golden fixtures from the LIVE reference (dev container only).

    python tests/golden/make_golden.py

Everything written here is an output of `/root/reference` itself, loaded through
oracle/ref_loader.py (three numpy/matplotlib shims, nothing else changed).  The fixtures travel
to the GPU box; the reference does not.

Files (all under tests/golden/):
  instances_40_0.3.npz   compact copy of the INPUT instances test{16,32,64}_40_0.3.pkl
                          (bit-packed maps, uint8 coordinates) + sha256 of the source pkls
  traces.npz              reference step/observe traces on selected instances under two recorded
                          action streams: U (uniform) and G (navi-greedy, eps = 0.1)
  navi.npz                reference navi maps (packed) for selected instances, sha256 of all 200,
                          and search.compute_heuristics distance maps for selected goals
  crafted.npz             hand-built conflict cases run through the reference
  per.npz                 buffer.SumTree update/sample rounds and LocalBuffer.finish TD vectors
  generator_stats.npz     histograms over instances drawn by the reference's own Environment.__init__ / reset()
                          (density, start-goal distance, component share, components used), SURVEY 8f-3
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402

TRACE_INSTANCES = {16: [0, 7, 42, 199], 32: [0, 7, 42, 101, 150, 199], 64: [0, 7, 42, 199]}
TRACE_STEPS = 96
NAVI_FULL_INSTANCES = [0, 199]


def sha8(b: bytes) -> np.uint64:
    return np.frombuffer(hashlib.sha256(b).digest()[:8], dtype=np.uint64)[0]


def greedy_actions(obs: np.ndarray, rng: np.random.Generator, eps: float = 0.1) -> np.ndarray:
    """navi-greedy stream 'G' (SURVEY 8d): uniform among set heuristic bits of the centre cell."""
    N = obs.shape[0]
    acts = np.zeros(N, dtype=np.int64)
    for i in range(N):
        if rng.random() < eps:
            acts[i] = rng.integers(0, 5)
            continue
        dirs = np.flatnonzero(obs[i, 2:6, 4, 4])
        acts[i] = 0 if dirs.size == 0 else 1 + dirs[rng.integers(0, dirs.size)]
    return acts


def make_instances():
    out = {}
    maps_ref = None
    for N in (16, 32, 64):
        maps, agents, goals = ref_loader.load_pkl(N)
        m = np.stack([np.asarray(x) != 0 for x in maps]).astype(np.uint8)  # [200,40,40]
        if maps_ref is None:
            maps_ref = m
        assert np.array_equal(maps_ref, m), "the three pkls share their maps (SURVEY §2)"
        out[f"agents{N}"] = np.stack(agents).astype(np.uint8)
        out[f"goals{N}"] = np.stack(goals).astype(np.uint8)
        with open(os.path.join(ref_loader.REFERENCE_DIR, f"test{N}_40_0.3.pkl"), "rb") as f:
            out[f"pkl_sha256_{N}"] = np.frombuffer(hashlib.sha256(f.read()).digest(), dtype=np.uint8)
    out["maps_packed"] = np.packbits(maps_ref.reshape(200, -1), axis=1)  # [200,200]
    out["map_side"] = np.int32(40)
    np.savez_compressed(os.path.join(HERE, "instances_40_0.3.npz"), **out)
    print("instances: ok")


def make_traces():
    env_mod = ref_loader.load_environment()
    out = {}
    for N, insts in TRACE_INSTANCES.items():
        maps, agents, goals = ref_loader.load_pkl(N)
        for stream in ("U", "G"):
            A = np.zeros((len(insts), TRACE_STEPS, N), dtype=np.uint8)
            P = np.zeros((len(insts), TRACE_STEPS + 1, N, 2), dtype=np.uint8)
            R = np.zeros((len(insts), TRACE_STEPS, N), dtype=np.float32)
            D = np.zeros((len(insts), TRACE_STEPS), dtype=np.uint8)
            H = np.zeros((len(insts), TRACE_STEPS + 1), dtype=np.uint64)
            O_last = np.zeros((len(insts), N, 6, 9, 9), dtype=np.uint8)
            for q, k in enumerate(insts):
                rng = np.random.default_rng(1000 * N + k + (0 if stream == "U" else 500000))
                env = env_mod.Environment()
                env.load(maps[k], agents[k], goals[k])
                obs, pos = env.observe()
                P[q, 0] = pos
                H[q, 0] = sha8(obs.astype(np.uint8).tobytes())
                for s in range(TRACE_STEPS):
                    a = rng.integers(0, 5, size=N) if stream == "U" else greedy_actions(obs, rng)
                    (obs, pos), r, d, info = env.step(a.tolist())
                    assert info == {"step": s}
                    A[q, s] = a
                    P[q, s + 1] = pos
                    R[q, s] = np.asarray(r, dtype=np.float32)
                    D[q, s] = d
                    H[q, s + 1] = sha8(obs.astype(np.uint8).tobytes())
                O_last[q] = obs
            pre = f"n{N}_{stream}_"
            out[pre + "instances"] = np.asarray(insts, dtype=np.int32)
            out[pre + "actions"] = A
            out[pre + "pos"] = P
            out[pre + "rewards"] = R
            out[pre + "done"] = D
            out[pre + "obs_sha8"] = H
            out[pre + "obs_last_packed"] = np.packbits(O_last.reshape(len(insts), -1), axis=1)
            print(f"traces N={N} {stream}: collisions={(R == -0.5).sum()} done={int(D.sum())}")
    np.savez_compressed(os.path.join(HERE, "traces.npz"), **out)


def make_navi():
    env_mod = ref_loader.load_environment()
    search = ref_loader.load_module("search")
    out = {}
    for N in (16, 32, 64):
        maps, agents, goals = ref_loader.load_pkl(N)
        shas = np.zeros((200, 32), dtype=np.uint8)
        for k in range(200):
            env = env_mod.Environment()
            env.load(maps[k], agents[k], goals[k])
            nv = env.navi_map[:, :, 4:-4, 4:-4].astype(np.uint8)  # unpadded [N,4,40,40]
            assert env.navi_map.sum() == nv.sum()
            shas[k] = np.frombuffer(hashlib.sha256(nv.tobytes()).digest(), dtype=np.uint8)
            if k in NAVI_FULL_INSTANCES:
                out[f"navi{N}_{k}_packed"] = np.packbits(nv.reshape(-1))
        out[f"navi{N}_sha256"] = shas
        print(f"navi N={N}: ok")
    # distances from search.compute_heuristics (search.py:24-55) for instance 0 / 199 of test32
    maps, agents, goals = ref_loader.load_pkl(32)
    for k in NAVI_FULL_INSTANCES:
        m = np.asarray(maps[k])
        D = np.full((32, 40, 40), 2147483647, dtype=np.int32)
        for i in range(32):
            h = search.compute_heuristics(m, tuple(int(v) for v in goals[k][i]))
            for (x, y), c in h.items():
                D[i, x, y] = c
        out[f"dist32_{k}"] = D
    np.savez_compressed(os.path.join(HERE, "navi.npz"), **out)


CRAFTED = [
    # name, L, obstacles, agents, goals, actions          
    ("swap01", 3, [], [(0, 0), (0, 1)], [(2, 2), (2, 0)], [4, 3]),
    ("swap_flipped", 3, [], [(0, 1), (0, 0)], [(2, 2), (2, 0)], [3, 4]),
    ("swap12_agent0_parked", 3, [], [(2, 2), (0, 0), (0, 1)], [(2, 2), (2, 0), (2, 1)], [0, 4, 3]),
    ("blocked_train", 4, [(0, 3)], [(0, 2), (0, 1), (0, 0)], [(3, 3), (3, 2), (3, 1)], [4, 4, 4]),
    ("moving_train", 4, [], [(0, 2), (0, 1), (0, 0)], [(3, 3), (3, 2), (3, 1)], [4, 4, 4]),
    ("vacated_ok", 4, [], [(1, 0), (0, 1), (1, 1)], [(3, 3), (3, 2), (3, 0)], [4, 2, 2]),
    ("vacated_blocked", 4, [(2, 1)], [(1, 0), (0, 1), (1, 1)], [(3, 3), (3, 2), (3, 0)], [4, 2, 2]),
    ("contested_lowest_wins", 4, [], [(0, 1), (1, 0)], [(3, 3), (3, 2)], [2, 4]),
    ("contested_lowest_wins_flipped", 4, [], [(1, 0), (0, 1)], [(3, 3), (3, 2)], [4, 2]),
    ("into_parked", 3, [], [(0, 0), (0, 1)], [(2, 2), (2, 0)], [4, 0]),
    ("oob_all_sides", 2, [], [(0, 0), (0, 1), (1, 1), (1, 0)], [(1, 1), (1, 0), (0, 0), (0, 1)], [1, 4, 2, 3]),
    ("on_goal_into_wall", 3, [], [(0, 0), (2, 2)], [(0, 0), (0, 2)], [1, 0]),
    ("stay_on_off_goal", 3, [], [(0, 0), (2, 2)], [(0, 0), (0, 2)], [0, 0]),
    ("cycle4_plus_intruder", 3, [], [(0, 0), (0, 1), (1, 1), (1, 0), (2, 0)],
     [(2, 2), (2, 1), (0, 2), (1, 2), (0, 0)], [4, 2, 3, 1, 1]),
    ("cycle4_2x2", 2, [], [(0, 0), (0, 1), (1, 1), (1, 0)], [(1, 1), (1, 0), (0, 0), (0, 1)], [4, 2, 3, 1]),
    ("cycle3_with_obstacle", 2, [], [(0, 0), (0, 1), (1, 1)], [(1, 1), (1, 0), (0, 0)], [4, 2, 3]),
    ("finish", 3, [], [(0, 0), (2, 2)], [(0, 1), (2, 1)], [4, 3]),
    ("three_into_one", 3, [], [(0, 1), (1, 0), (1, 2), (2, 1)], [(2, 2), (0, 0), (2, 0), (0, 2)], [2, 4, 3, 1]),
    ("chain_behind_loser", 5, [], [(0, 2), (2, 0), (3, 0), (4, 0)], [(4, 4), (4, 3), (4, 2), (4, 1)], [2, 4, 1, 1]),
    ("mover_into_bounced", 3, [(0, 2)], [(0, 1), (0, 0)], [(2, 2), (2, 0)], [4, 4]),
    ("swap_then_third_into_cell", 3, [], [(0, 0), (0, 1), (1, 0)], [(2, 2), (2, 1), (2, 0)], [4, 3, 1]),
]


def make_crafted():
    env_mod = ref_loader.load_environment()
    out = {"names": np.asarray([c[0] for c in CRAFTED])}
    for name, L, obst, ag, gl, acts in CRAFTED:
        m = np.zeros((L, L), dtype=np.int64)
        for (x, y) in obst:
            m[x, y] = 1
        env = env_mod.Environment()
        env.load(m, np.asarray(ag, dtype=np.int64), np.asarray(gl, dtype=np.int64))
        (obs, pos), r, d, info = env.step(list(acts))
        # a second, all-stay step pins "step after done" and the step counter
        (obs2, pos2), r2, d2, info2 = env.step([0] * len(acts))
        out[name + "_map"] = m.astype(np.uint8)
        out[name + "_agents"] = np.asarray(ag, dtype=np.uint8)
        out[name + "_goals"] = np.asarray(gl, dtype=np.uint8)
        out[name + "_actions"] = np.asarray(acts, dtype=np.uint8)
        out[name + "_pos"] = pos.astype(np.uint8)
        out[name + "_rewards"] = np.asarray(r, dtype=np.float32)
        out[name + "_done"] = np.uint8(d)
        out[name + "_obs"] = obs.astype(np.uint8)
        out[name + "_rewards2"] = np.asarray(r2, dtype=np.float32)
        out[name + "_done2"] = np.uint8(d2)
        out[name + "_info2"] = np.int32(info2["step"])
        print(f"crafted {name}: pos={pos.tolist()} r={r} done={d}")
    np.savez_compressed(os.path.join(HERE, "crafted.npz"), **out)


def make_per():
    buffer = ref_loader.load_module("buffer")
    out = {}
    cap = 1 << 12
    tree = buffer.SumTree(cap)
    rng = np.random.default_rng(7)
    rounds = 24
    B = 64
    out["capacity"] = np.int64(cap)
    out["rounds"] = np.int32(rounds)
    for rd in range(rounds):
        n = int(rng.integers(1, 300))
        idx = rng.integers(0, cap, size=n).astype(np.int64)
        if rd % 3 == 0:  # contiguous episode-style insert with duplicates elsewhere
            start = int(rng.integers(0, cap - 256))
            idx = np.arange(start, start + 256, dtype=np.int64)
            n = 256
        if rd % 4 == 1:
            idx[: n // 2] = idx[n // 2: n // 2 * 2]  # forced duplicates: last writer wins
        pr = rng.random(n) ** 3 * 5.0
        pr[rng.random(n) < 0.1] = 0.0
        out[f"upd_idx_{rd}"] = idx.copy()
        out[f"upd_prio_{rd}"] = pr.copy()
        tree.batch_update(idx.copy(), pr.copy())
        out[f"tree_sha_{rd}"] = np.frombuffer(hashlib.sha256(tree.tree.tobytes()).digest(), dtype=np.uint8)
        # sample with recorded uniforms: np.random.uniform(0, interval, B) == interval * random_sample(B)
        seed = 100 + rd
        u = np.random.RandomState(seed).random_sample(B)
        np.random.seed(seed)
        sidx, sprio = tree.batch_sample(B)
        out[f"smp_u_{rd}"] = u
        out[f"smp_idx_{rd}"] = sidx.astype(np.int64)
        out[f"smp_prio_{rd}"] = sprio.astype(np.float64)
    out["tree_final"] = tree.tree.copy()

    # LocalBuffer.finish TD (buffer.py:153-179)
    for case, size in enumerate((1, 2, 5, 100, 256)):
        rng = np.random.default_rng(50 + case)
        init_obs = np.zeros((2, 6, 9, 9), dtype=bool)
        lb = buffer.LocalBuffer(0, 2, 10, init_obs)
        rews = rng.choice([-0.075, -0.5, 0.0, 3.0], size=size)
        qs = rng.normal(size=(size, 5)).astype(np.float32)
        acts = rng.integers(0, 5, size=size)
        for t in range(size):
            lb.add(qs[t], int(acts[t]), rews[t], init_obs, np.zeros((2, 256), dtype=np.float16),
                   np.zeros((2, 2), dtype=bool))
        res = lb.finish(None if case % 2 == 0 else rng.normal(size=5).astype(np.float32),
                        None if case % 2 == 0 else np.zeros((2, 2), dtype=bool))
        out[f"td_rew_{case}"] = np.asarray(rews, dtype=np.float16)
        out[f"td_q_{case}"] = qs
        out[f"td_act_{case}"] = acts.astype(np.uint8)
        out[f"td_out_{case}"] = res[7].astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "per.npz"), **out)
    print("per: ok")


def make_replay():
    """GlobalBuffer.add / sample_batch / update_priorities of the live reference (worker.py imported with a no-op
    ray stub) on the scenario of tests/test_replay.py::drive."""
    sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))
    from replay_cases import BATCH, CAPACITY
    from test_replay import drive
    worker = ref_loader.load_worker()

    def ref_sample(store, seed):
        np.random.seed(seed)
        return store.sample_batch(BATCH)

    outs = drive(worker.GlobalBuffer(CAPACITY), ref_sample)
    out = {}
    for r, o in enumerate(outs):
        o = [x.numpy() if hasattr(x, "numpy") else x for x in o]
        out[f"obs_packed_{r}"] = np.packbits(o[0].astype(bool).reshape(-1))
        assert set(np.unique(o[0]).tolist()) <= {0.0, 1.0}
        out[f"action_{r}"], out[f"reward_{r}"], out[f"done_{r}"], out[f"steps_{r}"] = o[1], o[2], o[3], o[4]
        out[f"bt_steps_{r}"], out[f"hidden_{r}"] = o[5], o[6]
        out[f"comm_packed_{r}"] = np.packbits(o[7].reshape(-1))
        out[f"idxes_{r}"], out[f"weights_{r}"], out[f"ptr_{r}"] = np.asarray(o[8], dtype=np.int64), o[9], np.int64(o[10])
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)
    print("replay: ok")


def make_generator_stats(samples: int = 3000):
    """Instances drawn by the reference generator itself (environment.py:100-138 in __init__, :146-192 in reset()),
    summarised by tests/helpers.generator_stats.  get_navi_map (which the constructor and reset() end in) is
    replaced by a no-op for this run only: it consumes no randomness and does not touch the instance."""
    import random
    sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))
    from helpers import GEN_CONFIGS, generator_stats
    env_mod = ref_loader.load_environment()
    real_navi = env_mod.Environment.get_navi_map
    env_mod.Environment.get_navi_map = lambda self: None
    env_mod.Environment.observe = lambda self: None
    out = {"samples": np.int64(samples)}
    try:
        for L, N in GEN_CONFIGS:
            np.random.seed(1000 + L)
            random.seed(2000 + L)
            maps, agents, goals = [], [], []
            env = env_mod.Environment(num_agents=N, map_length=L)   # __init__ draws the first instance
            for k in range(samples):
                if k:
                    env.reset()                                     # same procedure, float32 map
                maps.append(np.asarray(env.map != 0, dtype=np.uint8))
                agents.append(np.array(env.agents_pos)), goals.append(np.array(env.goals_pos))
            st = generator_stats(np.stack(maps), np.stack(agents), np.stack(goals))
            for name, h in st.items():
                out[f"{name}_{L}_{N}"] = h
            print("generator stats", L, N, {k: v.tolist() for k, v in st.items()})
    finally:
        env_mod.Environment.get_navi_map = real_navi
    np.savez_compressed(os.path.join(HERE, "generator_stats.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.available(), "reference not mounted"
    if len(sys.argv) > 1 and sys.argv[1] == "replay":
        make_replay()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "generator":
        make_generator_stats()
        sys.exit(0)
    make_instances()
    make_crafted()
    make_per()
    make_traces()
    make_navi()
    make_replay()
    make_generator_stats()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
