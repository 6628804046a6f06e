"""GPU: BASELINE configs[4] as a test case — the device-resident actor loop (env batch + batched PyTorch Q-net +
comm mask + episode recording into the ReplayStore + actor-TD priorities + auto reset), replayed episode by
episode through the oracle environment and the reference's LocalBuffer.finish arithmetic."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def run_actor(B, N, L, density, max_steps, steps, eps):
    import torch
    from mapf_rl_b200 import BatchedEnvironment, ReplayStore
    from mapf_rl_b200.actor import BatchedActor
    from mapf_rl_b200.qnet import Network
    torch.manual_seed(0)
    dev = "cuda:0"
    env = BatchedEnvironment(B, N, L, device=dev)
    net = Network().to(dev).eval()
    store = ReplayStore(4 * B, max_num_agents=N, device=dev)
    log = dict(inst={}, steps=[], episodes=[])

    def on_begin(actor, ids):
        m, p, g = env.map.cpu().numpy(), env.agents_pos.cpu().numpy(), env.goals_pos.cpu().numpy()
        for e in ids:
            log["inst"][int(e)] = (m[e].copy(), p[e].copy(), g[e].copy(), len(log["steps"]))

    def on_step(actor, a8, rewards, done):
        log["steps"].append((a8.cpu().numpy(), rewards.cpu().numpy(), done.cpu().numpy()))

    def on_episode(actor, ids, slots, sizes, dones, td):
        q = actor.q_buf[torch.as_tensor(ids, device=dev)].cpu().numpy()
        for k, e in enumerate(ids):
            log["episodes"].append(dict(env=int(e), slot=int(slots[k]), size=int(sizes[k]), done=bool(dones[k]),
                                        td=td[k].cpu().numpy(), q=q[k], inst=log["inst"][int(e)]))

    actor = BatchedActor(env, net, store, epsilon=eps, seed=1, density=density, max_steps=max_steps,
                         on_step=on_step, on_episode=on_episode, on_begin=on_begin)
    actor.run(steps)
    env.check()
    return actor, store, log


def verify(actor, store, log, N, max_steps):
    S = store.max_steps
    tree = store.priority_tree.tree.cpu().numpy()
    cap = store.priority_tree.capacity
    obs_buf, comm_buf = store.obs_buf.cpu().numpy(), store.comm_mask.cpu().numpy()
    act_buf, rew_buf = store.act_buf.cpu().numpy(), store.rew_buf.cpu().numpy()
    hid_buf = store.hid_buf.cpu().numpy()
    size_buf, done_buf = store.size_buf.cpu().numpy(), store.done_buf.cpu().numpy()
    live = {}
    for ep in log["episodes"]:
        live[ep["slot"]] = ep              # later episodes overwrite earlier ones in the same slot
    running = {int(x) for x in actor._slot_host}
    live = {s: ep for s, ep in live.items() if s not in running}   # slots re-taken by a running episode are being overwritten
    assert len(live) >= 2
    n_done = 0
    for slot, ep in live.items():
        e, size = ep["env"], ep["size"]
        m, p, g, s0 = ep["inst"]
        o = oracle.OracleEnv()
        o.load(m, p.astype(np.int64), g.astype(np.int64))
        row0 = slot * (S + 1)
        assert np.array_equal(obs_buf[row0], o.observe()[0].astype(np.uint8))
        assert size_buf[slot] == size and bool(done_buf[slot]) == ep["done"]
        assert ep["done"] or size == max_steps
        for t in range(size):
            a, r, d = log["steps"][s0 + t]
            assert np.array_equal(comm_buf[row0 + t], oracle.comm_mask(o.agents_pos)), (slot, t)
            prev_comm = oracle.comm_mask(o.agents_pos)
            (oo, op), orw, od, _ = o.step(a[e])
            assert np.array_equal(obs_buf[row0 + t + 1], oo.astype(np.uint8)), (slot, t)
            assert np.array_equal(np.asarray(orw, dtype=np.float32), r[e]) and int(od) == d[e]
            assert act_buf[slot * S + t] == a[e, 0]
            assert rew_buf[slot * S + t] == np.float16(orw[0])
            assert (hid_buf[slot * S + t] == hid_buf[slot * S + t, 0]).all()      # agent 0's vector in every row (q8)
            assert od == (ep["done"] and t == size - 1)
        # comm_buf[size]: last mask again when cut at max_steps, zeros when done (worker.py:395-401)
        want_last = np.zeros((N, N), dtype=np.uint8) if ep["done"] else prev_comm
        assert np.array_equal(comm_buf[row0 + size], want_last)
        # initial priorities: LocalBuffer.finish (buffer.py:170-177) then ** alpha (worker.py:94)
        td = oracle.actor_td(rew_buf[slot * S:slot * S + size], ep["q"], act_buf[slot * S:slot * S + size], capacity=S)
        np.testing.assert_allclose(ep["td"], td, rtol=1e-6, atol=1e-7)   # q is fp32; north star: within 1e-5 relative
        leaves = tree[cap - 1 + slot * S: cap - 1 + (slot + 1) * S]
        np.testing.assert_allclose(leaves, td ** store.alpha, rtol=1e-6, atol=1e-9)
        assert (leaves[size:] == 0).all()
        n_done += ep["done"]
    # running episodes are not sampleable: their leaves are zero
    for e in range(actor.B):
        s = int(actor._slot_host[e])
        assert (tree[cap - 1 + s * S: cap - 1 + (s + 1) * S] == 0).all()
    assert abs(tree[0] - tree[cap - 1:].sum()) < 1e-6 * max(1.0, tree[0])      # root == sum of leaves (buffer.py:30)
    assert store.size == sum(int(x) for x in store._size_host)
    return n_done


def test_actor_loop_truncated_episodes():
    actor, store, log = run_actor(B=16, N=3, L=8, density=0.2, max_steps=12, steps=40, eps=0.3)
    verify(actor, store, log, 3, 12)
    assert actor.episodes >= 16 * 3


def test_actor_loop_done_episodes_and_learner_tail():
    import torch
    actor, store, log = run_actor(B=32, N=1, L=4, density=0.0, max_steps=20, steps=60, eps=1.0)
    assert verify(actor, store, log, 1, 20) >= 1          # a single random walker on a 4x4 board reaches its goal
    # learner side (worker.py:287-331): sample -> two bootstraps -> fused TD / priority / tree update
    batch = store.sample_batch(32)
    obs, action, reward, done, steps, bt_steps, hidden, comm, idxes, weights, old_ptr = batch
    net = actor.net
    with torch.no_grad():
        q_next = net.bootstrap(obs, bt_steps + steps.squeeze(1).long(), hidden, comm)
        q = net.bootstrap(obs[:, :-store.forward_steps], bt_steps, hidden, comm[:, :-store.forward_steps])
    before = store.priority_tree.tree.clone()
    td, prio = store.update_priorities_device(q, q_next, action.squeeze(1), reward.squeeze(1), done.squeeze(1),
                                              steps.squeeze(1), torch.as_tensor(idxes, device="cuda:0"), old_ptr)
    ref_td, ref_pr = oracle.learner_td(q.float().cpu().numpy(), q_next.float().cpu().numpy(), action.cpu().numpy(),
                                       reward.float().cpu().numpy(), done.float().cpu().numpy(), steps.float().cpu().numpy())
    np.testing.assert_allclose(td.cpu().numpy(), ref_td, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prio.cpu().numpy(), ref_pr, rtol=1e-5, atol=1e-6)
    assert not torch.equal(before, store.priority_tree.tree)
