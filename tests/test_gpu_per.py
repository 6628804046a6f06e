"""GPU parity for K4: sum tree (fp64, bit-exact vs buffer.SumTree goldens and the C oracle), the actor-side
TD (bit-exact) and the fused learner TD/priority tail (1e-5 relative, fp32 — north star tolerance)."""
import hashlib

import numpy as np
import pytest

from helpers import golden
from oracle import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: TD-errors and priorities within 1e-5 relative (fp32)


def test_sumtree_golden_rounds():
    from mapf_rl_b200 import SumTree
    z = golden("per.npz")
    tree = SumTree(int(z["capacity"]))
    assert tree.layer == 13
    for rd in range(int(z["rounds"])):
        idx = z[f"upd_idx_{rd}"].copy()
        tree.batch_update(idx, z[f"upd_prio_{rd}"])
        assert np.array_equal(idx, z[f"upd_idx_{rd}"] + tree.capacity - 1)  # buffer.py:96
        got = np.frombuffer(hashlib.sha256(tree.tree.cpu().numpy().tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(got, z[f"tree_sha_{rd}"]), rd
        np.random.seed(100 + rd)  # the reference consumed np.random.uniform under this seed
        sidx, sprio = tree.batch_sample(len(z[f"smp_u_{rd}"]))
        assert np.array_equal(sidx, z[f"smp_idx_{rd}"]), rd
        assert np.array_equal(sprio, z[f"smp_prio_{rd}"]), rd
    t = tree.tree.cpu().numpy()
    assert np.array_equal(t, z["tree_final"])
    assert abs(t[-tree.capacity:].sum() - t[0]) < 0.1  # buffer.py:105
    assert tree.sum() == t[0] and tree[5] == t[tree.capacity - 1 + 5]


@pytest.mark.parametrize("cap,n", [(1 << 19, 192), (1 << 19, 256), (1 << 10, 5000), (1 << 4, 64), (1, 3)])
def test_sumtree_vs_oracle_reference_capacity(cap, n):
    """2^19 leaves = 2048 episode slots x 256 steps (train.py:21, config.py:29), batch 192 (config.py:25)."""
    from mapf_rl_b200 import SumTree
    rng = np.random.default_rng(cap + n)
    tree, ref = SumTree(cap), oracle.OracleSumTree(cap)
    for rd in range(6):
        idx = rng.integers(0, cap, size=n).astype(np.int64)
        if rd == 1 and cap >= 512:
            idx = np.arange(256, 512, dtype=np.int64)[:n]   # contiguous episode insert, worker.py:87-94
        if rd == 2:
            idx[: n // 3] = idx[-(n // 3):] if n >= 3 else idx[: n // 3]  # duplicates: last writer wins
        pr = rng.random(idx.shape[0]) ** 2 * 3
        pr[rng.random(idx.shape[0]) < 0.2] = 0.0
        if rd == 0:
            pr[:] = np.maximum(pr, 1e-3)
        tree.batch_update(idx.copy(), pr)
        ref.batch_update(idx.copy(), pr)
        assert np.array_equal(tree.tree.cpu().numpy(), ref.tree), rd
        if ref.tree[0] > 0:
            B = 192
            u = rng.random(B)
            gi, gp, gw = tree.sample_device(B, u, beta=0.4)
            ri, rp = ref.batch_sample(B, u)
            assert np.array_equal(gi.cpu().numpy(), ri) and np.array_equal(gp.cpu().numpy(), rp)
            w = np.power(rp / rp.min(), -0.4)  # worker.py:165-166
            np.testing.assert_allclose(gw.cpu().numpy(), w, rtol=RTOL)


@pytest.mark.parametrize("cap", [1 << 19, 1 << 12, 1 << 11, 1 << 10, 1 << 4, 2, 1])
@pytest.mark.parametrize("n", [1, 37, 192, 256])
def test_sumtree_sorted_batches_vs_oracle(cap, n):
    """Non-decreasing leaf indices (what the stratified sampler returns, what an episode insert is) take the update path
    that runs its level loop in shared memory (neighbour groups instead of L2 round trips): random sorted batches,
    heavy duplicates (the last one wins), contiguous runs, one leaf n times, a batch at the tree's right edge; the tree
    bit-equal to the oracle's after every round, and the sampler (top of the tree from shared memory) on top of it."""
    from mapf_rl_b200 import SumTree
    rng = np.random.default_rng(cap * 7 + n)
    tree, ref = SumTree(cap), oracle.OracleSumTree(cap)
    base = rng.random(min(cap, 4096)) + 0.01
    for s in range(0, cap, 4096):   # a full tree to start from
        ii = np.arange(s, min(cap, s + 4096), dtype=np.int64)
        tree.batch_update(ii.copy(), base[: len(ii)])
        ref.batch_update(ii.copy(), base[: len(ii)])
    assert np.array_equal(tree.tree.cpu().numpy(), ref.tree)
    for rd in range(6):
        if rd == 0:
            idx = np.sort(rng.integers(0, cap, size=n))
        elif rd == 1:
            idx = np.sort(rng.integers(0, max(1, min(cap, n // 3 + 1)), size=n) + rng.integers(0, max(1, cap - n)))
        elif rd == 2:
            idx = np.minimum(np.arange(n) + rng.integers(0, max(1, cap - n + 1)), cap - 1)
        elif rd == 3:
            idx = np.full(n, rng.integers(0, cap))
        elif rd == 4:
            idx = np.sort(np.maximum(cap - 1 - rng.integers(0, min(cap, 2 * n), size=n), 0))
        else:
            idx = np.sort(np.r_[rng.integers(0, cap, size=n - n // 2), np.repeat(rng.integers(0, cap), n // 2)])
        idx = idx.astype(np.int64)
        assert (np.diff(idx) >= 0).all() and idx.min() >= 0 and idx.max() < cap
        pr = rng.random(n) ** 2 * 3
        pr[rng.random(n) < 0.2] = 0.0
        tree.batch_update(idx.copy(), pr)
        ref.batch_update(idx.copy(), pr)
        assert np.array_equal(tree.tree.cpu().numpy(), ref.tree), rd
        tree.check()
        if ref.tree[0] > 0:
            B = 192
            u = rng.random(B)
            gi, gp, gw = tree.sample_device(B, u, beta=0.4)
            ri, rp = ref.batch_sample(B, u)
            assert np.array_equal(gi.cpu().numpy(), ri) and np.array_equal(gp.cpu().numpy(), rp)
            if rp.min() > 0:
                np.testing.assert_allclose(gw.cpu().numpy(), np.power(rp / rp.min(), -0.4), rtol=RTOL)


def test_actor_td_golden_and_oracle():
    from mapf_rl_b200 import LocalBuffer
    from mapf_rl_b200.buffer import actor_td_errors
    z = golden("per.npz")
    for case in range(5):
        rew, q, act = z[f"td_rew_{case}"], z[f"td_q_{case}"], z[f"td_act_{case}"]
        size = len(rew)
        init_obs = np.zeros((2, 6, 9, 9), dtype=bool)
        lb = LocalBuffer(0, 2, 10, init_obs)
        for t in range(size):
            lb.add(q[t], int(act[t]), rew[t], init_obs, np.zeros((2, 256), dtype=np.float16), np.zeros((2, 2), dtype=bool))
        res = lb.finish()
        assert len(res) == 11 and res[9] == size and res[8] is True
        assert res[3].shape == (size + 1, 2, 6, 9, 9) and res[7].shape == (256,)
        assert np.array_equal(res[7], z[f"td_out_{case}"]), case
    # batched, many episodes, vs the oracle
    rng = np.random.default_rng(3)
    E, cap = 64, 256
    size = rng.integers(1, cap + 1, size=E).astype(np.int32)
    rew = rng.choice([-0.075, -0.5, 0.0, 3.0], size=(E, cap)).astype(np.float16).astype(np.float32)
    q = rng.normal(size=(E, cap, 5)).astype(np.float32)
    act = rng.integers(0, 5, size=(E, cap)).astype(np.uint8)
    td = actor_td_errors(rew, q, act, size).cpu().numpy()
    for e in range(E):
        want = oracle.actor_td(rew[e, :size[e]].astype(np.float16), q[e], act[e, :size[e]])
        assert np.array_equal(td[e], want), e


@pytest.mark.parametrize("sorted_idx", [False, True])
@pytest.mark.parametrize("double_q", [False, True])
def test_learner_td_update(double_q, sorted_idx):
    import torch
    from mapf_rl_b200 import SumTree
    cap, slot = 1 << 12, 256
    rng = np.random.default_rng(11)
    tree, ref = SumTree(cap), oracle.OracleSumTree(cap)
    base_idx = np.arange(cap, dtype=np.int64)
    base_pr = rng.random(cap) + 0.01
    tree.batch_update(base_idx.copy(), base_pr)
    ref.batch_update(base_idx.copy(), base_pr)
    n = 192
    for old_ptr, ptr in [(0, 0), (2, 5), (14, 3)]:  # no overwrite / plain window / wrapped window (worker.py:192-201)
        qo = rng.normal(size=(n, 5)).astype(np.float32)
        qt = rng.normal(size=(n, 5)).astype(np.float32)
        qn = rng.normal(size=(n, 5)).astype(np.float32)
        act = rng.integers(0, 5, size=n)
        rew = rng.choice([-0.075, -0.5, 0.0, 3.0], size=n).astype(np.float32)
        done = (rng.random(n) < 0.1).astype(np.float32)
        steps = rng.integers(1, 3, size=n).astype(np.float32)
        idx = rng.integers(0, cap, size=n).astype(np.int64)
        idx[:8] = idx[8:16]
        if sorted_idx:   # the learner's indices come from the stratified sampler: sorted (stale ones masked in the middle)
            idx = np.sort(idx)
        td, pr = tree.td_update(qo, qt, act, rew, done, steps, idx, old_ptr=old_ptr, ptr=ptr, slot_steps=slot,
                                q_online_next=qn if double_q else None)
        if double_q:
            boot = qt[np.arange(n), qn.argmax(1)]
            want_td = qo[np.arange(n), act] - (rew + np.power(np.float32(0.99), steps) * ((1 - done) * boot))
            want_pr = np.maximum(np.abs(want_td), 1e-6)
        else:
            want_td, want_pr = oracle.learner_td(qo, qt, act, rew, done, steps)
        np.testing.assert_allclose(td.cpu().numpy(), want_td, rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(pr.cpu().numpy(), want_pr, rtol=RTOL, atol=1e-7)
        # reference update_priorities on the same numbers
        if ptr > old_ptr:
            mask = (idx < old_ptr * slot) | (idx >= ptr * slot)
        elif ptr < old_ptr:
            mask = (idx < old_ptr * slot) & (idx >= ptr * slot)
        else:
            mask = np.ones(n, dtype=bool)
        assert 0 < mask.sum() and (mask.sum() < n or ptr == old_ptr)
        ref.batch_update(idx[mask].copy(), want_pr[mask].astype(np.float64) ** 0.6)
        got = tree.tree.cpu().numpy()
        np.testing.assert_allclose(got, ref.tree, rtol=RTOL)
        leaves = got[cap - 1:]
        assert abs(leaves.sum() - got[0]) < 1e-6 * got[0]
