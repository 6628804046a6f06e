"""Host-side instance generation following the reference's procedure (environment.py:21-70, 100-138).

Used by the drop-in `Environment` constructor / `reset()` and to make synthetic batches for tests and
benchmarks.  The RNG stream of the reference (np.random + random globals) is NOT reproduced; the
distribution is: obstacle map iid Bernoulli(density) (density ~ triangular(0, 0.33, 0.5) when not
given, environment.py:100), start uniform over cells of components with >= 2 remaining cells (choosing a
component with probability proportional to its size and then a uniform cell in it, :120-131, is a
uniform cell), goal uniform over the remaining cells of the same component (:133-135), components
with < 2 remaining cells dropped (:137).
"""
from __future__ import annotations

import numpy as np


def map_partition(map_: np.ndarray):
    """4-connected components of free cells -> int32 label map (-1 obstacle) and component sizes.
    Same partition as environment.py:21-70."""
    L0, L1 = map_.shape
    free = map_ == 0
    label = np.full((L0, L1), -1, dtype=np.int32)
    sizes = []
    for sx, sy in np.argwhere(free):
        if label[sx, sy] >= 0:
            continue
        cid = len(sizes)
        stack = [(int(sx), int(sy))]
        label[sx, sy] = cid
        count = 0
        while stack:
            x, y = stack.pop()
            count += 1
            for nx, ny in ((x - 1, y), (x + 1, y), (x, y - 1), (x, y + 1)):
                if 0 <= nx < L0 and 0 <= ny < L1 and free[nx, ny] and label[nx, ny] < 0:
                    label[nx, ny] = cid
                    stack.append((nx, ny))
        sizes.append(count)
    return label, np.asarray(sizes, dtype=np.int64)


def generate_instance(rng: np.random.Generator, map_length: int, num_agents: int, density=None, max_tries: int = 1000,
                      return_density: bool = False):
    """-> (map uint8[L,L], agents int64[N,2], goals int64[N,2]) [+ the density the map was drawn with]"""
    L, N = map_length, num_agents
    d = rng.triangular(0, 0.33, 0.5) if density is None else float(density)
    for _ in range(max_tries):
        m = (rng.random((L, L)) < d).astype(np.uint8)
        label, sizes = map_partition(m)
        if not np.any(sizes >= 2):
            continue  # environment.py:107-110 regenerates the map
        remaining = sizes.copy()
        avail = label >= 0
        agents = np.empty((N, 2), dtype=np.int64)
        goals = np.empty((N, 2), dtype=np.int64)
        ok = True
        for i in range(N):
            elig = avail & (remaining[np.maximum(label, 0)] >= 2) & (label >= 0)
            cells = np.argwhere(elig)
            if len(cells) == 0:
                ok = False  # the reference would raise here (random.randint(0, -1)); we redraw the map
                break
            s = cells[rng.integers(0, len(cells))]
            c = label[s[0], s[1]]
            avail[s[0], s[1]] = False
            same = np.argwhere(avail & (label == c))
            g = same[rng.integers(0, len(same))]
            avail[g[0], g[1]] = False
            remaining[c] -= 2
            agents[i], goals[i] = s, g
        if ok:
            return (m, agents, goals, d) if return_density else (m, agents, goals)
    raise RuntimeError("no empty position")  # environment.py:31


def generate_batch(num: int, map_length: int, num_agents: int, density=0.3, seed: int = 0, first_index: int = 0):
    """Independent instances; instance k is drawn from SeedSequence([seed, first_index + k]) so any shard
    of a multi-GPU job gets the same instances a single process would."""
    maps = np.empty((num, map_length, map_length), dtype=np.uint8)
    agents = np.empty((num, num_agents, 2), dtype=np.uint8)
    goals = np.empty((num, num_agents, 2), dtype=np.uint8)
    for k in range(num):
        rng = np.random.default_rng(np.random.SeedSequence([seed, first_index + k]))
        m, a, g = generate_instance(rng, map_length, num_agents, density)
        maps[k], agents[k], goals[k] = m, a, g
    return maps, agents, goals
