"""Quick device timing of the opNav step kernel (development aid; bench.py is the judged measurement)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basilisk_env_b200.opnav_env import OpNavVecEnv

import os
for n in [int(v) for v in os.environ.get("OPNAV_N", "4096,32768,65536").split(",")]:
    env = OpNavVecEnv(n, device=0, auto_reset=True, sample_orbit=1, camera_reenable=1, noise_seed=1)
    env.reset(seed=1)
    torch.manual_seed(0)
    acts = [torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
    for a in acts[:2]:
        env.step(a)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    K = 3
    for k in range(K):
        env.step(acts[k % 4])
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    fl = env.flops_per_step()
    print(f"opnav n={n}: {ms:.2f} ms/step, {n / ms * 1e3:.0f} env-steps/s, model {fl * n / ms / 1e9:.2f} TFLOP/s", env.episode_stats())
    env.close()
