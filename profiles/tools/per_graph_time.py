#!/usr/bin/env python
"""GPU-side time of the PER kernels (GPU box): each call is captured 64 times into one CUDA graph and the graph is replayed,
so the Python / launch cost of the eager call path (~15 us, which is what bench_configs.py's K4 line mostly measures) drops out.

    python profiles/tools/per_graph_time.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200 import SumTree  # noqa: E402

dev = torch.device("cuda", 0)
cap = 2048 * 256
tree = SumTree(cap, device=dev)
rng = np.random.default_rng(0)
pr_all = torch.as_tensor(rng.random(cap) ** 0.6, dtype=torch.float64, device=dev)
idx_all = torch.arange(cap, dtype=torch.int64, device=dev)
for s in range(0, cap, 4096):
    tree.update_device(idx_all[s:s + 4096], pr_all[s:s + 4096])
u = torch.rand(192, dtype=torch.float64, device=dev)
idx, pr, w = tree.sample_device(192, u, beta=0.4)
newp = torch.rand(192, dtype=torch.float64, device=dev)
ep_idx = torch.arange(256, dtype=torch.int64, device=dev) + 256 * 77
ep_pr = torch.rand(256, dtype=torch.float64, device=dev)


def graph_time(fn, reps=64, launches=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * launches)


import ctypes as C  # noqa: E402
from mapf_rl_b200 import _native  # noqa: E402


def raw_update(i, p):  # SumTree.update_device synchronises the stream after the launch; the bare launch is what is timed
    _native.check(tree._lib.mapf_per_update(tree._h, C.c_void_p(i.data_ptr()), C.c_void_p(p.data_ptr()), int(i.numel()), tree._stream()))


upd = dict(q_online=torch.randn(192, 5, device=dev), q_target_next=torch.randn(192, 5, device=dev),
           action=torch.randint(0, 5, (192,), device=dev), reward=torch.randn(192, device=dev),
           done=torch.zeros(192, device=dev), steps=torch.full((192,), 2.0, device=dev), idx=idx)
cyc_out = tree.cycle(update=upd, sample_size=192, uniforms=u, beta=0.4)

out = {"cycle_192_192_us": round(graph_time(lambda: tree.cycle(update=upd, sample_size=192, uniforms=u, beta=0.4, out=cyc_out)), 2),
       "batch_sample_192_us": round(graph_time(lambda: tree.sample_device(192, u, beta=0.4)), 2),
       "batch_update_192_us": round(graph_time(lambda: raw_update(idx, newp)), 2),
       "episode_insert_256_us": round(graph_time(lambda: raw_update(ep_idx, ep_pr)), 2)}
print(json.dumps(out))
