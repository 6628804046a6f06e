"""GPU: the learner half of BASELINE configs[4] — BatchedLearner.update (window gather kernel -> two PyTorch bootstrap
passes -> loss / Adam -> ONE mapf_per_cycle launch) on a store filled by the device-resident actor loop.  The kernel's TD
errors are checked against the PyTorch restatement of worker.py:300-308 at 1e-5 relative, the tree against its own leaves,
and the sampled indices against the episodes that exist."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_learner_updates_on_actor_filled_store():
    import torch
    from mapf_rl_b200 import BatchedEnvironment, ReplayStore
    from mapf_rl_b200.actor import BatchedActor
    from mapf_rl_b200.learner import BatchedLearner
    from mapf_rl_b200.qnet import Network
    torch.manual_seed(0)
    dev = "cuda:0"
    B, N, L, cap = 48, 4, 12, 16
    env = BatchedEnvironment(B, N, L, device=dev)
    net = Network().to(dev)
    store = ReplayStore(128, max_num_agents=N, device=dev, max_steps=cap)
    actor = BatchedActor(env, net, store, epsilon=0.3, seed=1, density=0.15, max_steps=cap)
    actor.run(2 * cap + 3)
    assert actor.episodes >= 2 * B and len(store) >= 192
    learner = BatchedLearner(net, store, batch_size=64, target_update_freq=3, seed=5)
    assert learner.ready(64)
    w0 = [p.detach().clone() for p in net.parameters()]
    tree = store.priority_tree
    for it in range(5):
        idx_used = None if learner._next is None else learner._next[0].clone()
        stats = learner.update(want_stats=True)
        assert np.isfinite(stats["loss"]) and stats["prio_min"] >= 1e-6
        assert stats["kernel_td_vs_torch"] <= 1e-5 * max(1.0, stats["td_abs_mean"]) + 1e-6, stats
        t = tree.tree.cpu().numpy()
        cap_leaves = tree.capacity
        # every ancestor is the sum of its children (buffer.py:99-105), root == sum of leaves
        inner = t[:cap_leaves - 1]
        assert np.array_equal(inner, t[1:2 * cap_leaves - 1:2] + t[2:2 * cap_leaves - 1:2])
        assert abs(t[0] - t[cap_leaves - 1:].sum()) < 1e-6 * max(1.0, t[0])
        nxt = learner._next[0].cpu().numpy()
        size = store.size_buf.cpu().numpy()
        assert ((nxt % cap) < size[nxt // cap]).all(), "a sampled transition lies beyond its episode (worker.py:120)"
    tree.check()
    env.check()
    assert any(not torch.equal(a, b.detach()) for a, b in zip(w0, net.parameters())), "the optimizer never stepped"
    assert learner.counter == 5
    # the target network followed at update 3 only
    same = all(torch.equal(a, b) for a, b in zip(learner.tar_model.parameters(), net.parameters()))
    assert not same
