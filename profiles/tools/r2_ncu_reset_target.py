import sys
sys.path.insert(0, '/root/repo')
import torch
from mapf_rl_b200 import BatchedEnvironment
env = BatchedEnvironment(8192, 32, 40)
for i in range(3):
    env.reset(seed=0, env_offset=i * 8192, density=0.3)
torch.cuda.synchronize()
env.check()
