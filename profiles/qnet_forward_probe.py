#!/usr/bin/env python
"""How fast can the (out-of-scope, PyTorch) Q-network forward of the actor loop go on a B200 without touching its
arithmetic?  Times qnet.Network.step on 2048 envs x 32 agents under the options an actor can switch on:
autocast bf16 / weights in bf16, channels_last, CUDA-graph replay.  (GPU box.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mapf_rl_b200.qnet import Network  # noqa: E402

B, N = 2048, 32
dev = torch.device("cuda", 0)
obs = (torch.rand(B, N, 6, 9, 9, device=dev) < 0.3).to(torch.uint8)
comm = torch.rand(B, N, N, device=dev) < 0.1
comm |= torch.eye(N, dtype=torch.bool, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, net, autocast=None, chlast=False):
    x = obs
    if chlast:
        net = net.to(memory_format=torch.channels_last)

    def fn():
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                net.step(x, comm)
        else:
            net.step(x, comm)
    ms = timeit(fn)
    print(f"{name:48s} {ms:8.2f} ms  {B * N / ms / 1e3:8.2f} M agent-forwards/s", flush=True)


torch.backends.cudnn.benchmark = True
run("fp32 eager", Network().to(dev).eval())
run("autocast bf16", Network().to(dev).eval(), autocast=torch.bfloat16)
run("autocast bf16 + channels_last", Network().to(dev).eval(), autocast=torch.bfloat16, chlast=True)
run("weights bf16", Network().to(dev).eval().to(torch.bfloat16))
run("weights bf16 + channels_last", Network().to(dev).eval().to(torch.bfloat16), chlast=True)
run("weights fp16 + channels_last", Network().to(dev).eval().to(torch.float16), chlast=True)
