import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
from basilisk_env_b200.opnav_env import OpNavVecEnv
env = LeoPowerAttVecEnv(200, device=0, auto_reset=True, step_duration=20.0, max_length=2)
env.reset(seed=1)
for t in range(4):
    env.step(torch.full((200,), t % 3, dtype=torch.int32, device="cuda"))
env.step_host(np.zeros(200, np.int32))
print("leo", env.episode_stats()); env.close()
on = OpNavVecEnv(100, device=0, auto_reset=True, step_duration_min=2.0, max_length=2, camera_reenable=1, sample_orbit=1)
on.reset(seed=2)
for t in range(4):
    on.step(torch.full((100,), t % 2, dtype=torch.int32, device="cuda"))
on.step_host(np.zeros(100, np.int32))
print("opnav", on.episode_stats()); on.close()
