import sys, hashlib
sys.path.insert(0, '/root/repo')
import torch
from mapf_rl_b200 import BatchedEnvironment
h = hashlib.sha256()
for (B, N, L, dens) in ((512, 32, 40, 0.3), (128, 64, 80, 0.3), (256, 7, 13, None), (64, 100, 120, 0.2)):
    env = BatchedEnvironment(B, N, L)
    env.reset(seed=5, env_offset=17, density=dens)
    env.check()
    for t in (env.map, env.agents_pos, env.goals_pos):
        h.update(t.cpu().numpy().tobytes())
print(h.hexdigest())
