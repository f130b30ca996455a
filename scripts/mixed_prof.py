import sys, os
sys.path.insert(0, os.getcwd())
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
n = 65536
env = LeoPowerAttVecEnv(n, device=0, seed=77, precision=1, use_j2=1, rw_set=1)
env.reset()
a = torch.randint(0, 3, (4, n), dtype=torch.int32, device="cuda")
for t in range(4):
    env.step(a[t])
torch.cuda.synchronize()
