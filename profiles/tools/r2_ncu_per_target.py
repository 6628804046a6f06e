#!/usr/bin/env python
"""Target for one ncu --set full capture of the sum-tree kernels on their batch-sized paths (2^19 leaves, batch 192 with the
sorted indices the stratified sampler returns; a 256-leaf episode insert):

    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o gpurun_out/r2_per \
        python profiles/tools/r2_ncu_per_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mapf_rl_b200 import SumTree, config  # noqa: E402

nvtx = torch.cuda.nvtx
dev = "cuda:0"


def cap(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    nvtx.range_push("cap")
    fn()
    torch.cuda.synchronize()
    nvtx.range_pop()


cap_leaves = 1 << 19
tree = SumTree(cap_leaves, device=dev)
rng = np.random.default_rng(0)
allidx = torch.arange(cap_leaves, dtype=torch.int64, device=dev)
allpr = torch.rand(cap_leaves, dtype=torch.float64, device=dev) + 1e-3
for s in range(0, cap_leaves, 4096):
    tree.update_device(allidx[s:s + 4096], allpr[s:s + 4096])
n = config.batch_size
u = torch.rand(n, dtype=torch.float64, device=dev)
idx, _, _ = tree.sample_device(n, u, beta=0.4)          # sorted: what the learner hands back to update_priorities
assert bool((idx[1:] >= idx[:-1]).all())
pr = torch.rand(n, dtype=torch.float64, device=dev) + 1e-3
cap(lambda: tree.update_device(idx, pr))                 # per_update_kernel, sorted path
cap(lambda: tree.sample_device(n, u, beta=0.4))          # per_sample_kernel
upd = dict(q_online=torch.randn(n, 5, device=dev), q_target_next=torch.randn(n, 5, device=dev),
           action=torch.randint(0, 5, (n,), device=dev), reward=torch.zeros(n, device=dev), done=torch.zeros(n, device=dev),
           steps=torch.ones(n, device=dev), idx=idx)
out = tree.cycle(update=upd, sample_size=n, uniforms=u, beta=0.4)
cap(lambda: tree.cycle(update=upd, sample_size=n, uniforms=u, beta=0.4, out=out))   # per_cycle_kernel
ep = torch.arange(256, dtype=torch.int64, device=dev) + 256 * 77
eppr = torch.rand(256, dtype=torch.float64, device=dev)
cap(lambda: tree.update_device(ep, eppr))                # episode insert (worker.py:87-94)
tree.check()
print("done")
