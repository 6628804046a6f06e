"""CPU, world_size 2 over gloo: the multi-GPU path has no data-path collective — what must hold is that
shards tile the batch exactly, that every rank generates the instances a single process would for its
global indices, and that the timing reductions (max / sum over ranks) behave."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mapf_rl_b200 import sharding
from mapf_rl_b200.instances import generate_batch


def test_shard_bounds_tile_the_batch():
    for total in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)
    assert sharding.weak_offset(8192, 3) == 24576


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_gpu, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.rank_world() == (rank, world, rank)
        off = sharding.weak_offset(per_gpu, rank)
        maps, agents, goals = generate_batch(per_gpu, 12, 5, density=0.3, seed=9, first_index=off)
        # strong split of a fixed job gives the same instances too
        lo, hi = sharding.shard_bounds(world * per_gpu, world, rank)
        m2, a2, g2 = generate_batch(hi - lo, 12, 5, density=0.3, seed=9, first_index=lo)
        assert np.array_equal(maps, m2) and np.array_equal(agents, a2) and np.array_equal(goals, g2)
        gathered = [None] * world
        dist.all_gather_object(gathered, (maps, agents, goals))
        t_max = sharding.max_over_ranks(10.0 + rank)
        t_sum = sharding.sum_over_ranks(float(per_gpu))
        assert sharding.gather_floats(3.0 + rank) == [3.0 + r for r in range(world)]
        # the learner's gradient averaging (the package's only collective): one all-reduce of the flattened gradients
        import torch
        from mapf_rl_b200.learner import BatchedLearner
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        for k, p in enumerate(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (k + 1))
        shim = type("L", (), {"allreduce": True, "model": model})()
        BatchedLearner._allreduce_grads(shim)
        mean = sum(range(1, world + 1)) / world
        for k, p in enumerate(model.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, mean * (k + 1)))
        dist.barrier()
        if rank == 0:
            q.put((gathered, t_max, t_sum))
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_equal_single_process_batch():
    world, per_gpu = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t_max, t_sum = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t_max == 11.0 and t_sum == 12.0
    maps, agents, goals = generate_batch(world * per_gpu, 12, 5, density=0.3, seed=9, first_index=0)
    assert np.array_equal(np.concatenate([g[0] for g in gathered]), maps)
    assert np.array_equal(np.concatenate([g[1] for g in gathered]), agents)
    assert np.array_equal(np.concatenate([g[2] for g in gathered]), goals)
