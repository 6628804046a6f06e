"""Seeded CBS instances shared by tests/golden/make_golden_cbs.py (reference side) and tests/test_cbs.py."""
import numpy as np

# (map side, agents, obstacle density)
CBS_CASES = [(8, 3, 0.15), (8, 4, 0.2), (8, 5, 0.1), (10, 4, 0.2), (10, 5, 0.25), (10, 6, 0.15), (12, 4, 0.25), (12, 6, 0.2),
             (12, 8, 0.1), (6, 4, 0.1), (6, 5, 0.0), (7, 6, 0.1), (9, 7, 0.15), (14, 6, 0.25), (16, 8, 0.2), (5, 4, 0.0),
             (10, 8, 0.1), (12, 5, 0.3), (16, 6, 0.3), (20, 8, 0.25), (8, 6, 0.2), (9, 5, 0.3), (11, 7, 0.2), (13, 9, 0.1)]


def cbs_instance(k, L, N, density):
    """Random map; starts and goals are 2N distinct cells of its largest 4-connected component."""
    rng = np.random.default_rng(1000 + k)
    while True:
        m = (rng.random((L, L)) < density).astype(np.int64)
        comp = -np.ones((L, L), dtype=np.int64)
        best, best_cells = -1, []
        for x0 in range(L):
            for y0 in range(L):
                if m[x0, y0] or comp[x0, y0] >= 0:
                    continue
                cells, stack = [], [(x0, y0)]
                comp[x0, y0] = x0 * L + y0
                while stack:
                    x, y = stack.pop()
                    cells.append((x, y))
                    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                        u, v = x + dx, y + dy
                        if 0 <= u < L and 0 <= v < L and not m[u, v] and comp[u, v] < 0:
                            comp[u, v] = x0 * L + y0
                            stack.append((u, v))
                if len(cells) > len(best_cells):
                    best_cells = cells
        if len(best_cells) >= 2 * N:
            break
    pick = rng.permutation(len(best_cells))[:2 * N]
    cells = np.asarray(best_cells, dtype=np.int64)[pick]
    return m, cells[:N], cells[N:]
