#!/usr/bin/env python
"""opNav: device-timed launch with i.i.d. actions against all-0 (OpNavOD task set) and all-1 (sun-safe task set) launches --
which bucket holds the slow warps (the work queue should hand those out first).   python scripts/opnav_bucket_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.opnav_env import OpNavVecEnv
n = int(os.environ.get("OPNAV_N", "113664"))
out = {}
for name in ("iid", "all0", "all1"):
    env = OpNavVecEnv(n, device=0, auto_reset=True, sample_orbit=1, camera_reenable=1, noise_seed=1)
    env.reset(seed=1)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = torch.randint(0, 2, (6, n), dtype=torch.int32, device="cuda", generator=g)
    if name != "iid":
        acts[2:] = 0 if name == "all0" else 1
    for t in range(3):
        env.step(acts[t])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for t in range(3):
        env.step(acts[3 + t]); ev[t + 1].record()
    torch.cuda.synchronize()
    out[name] = round(float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(3)])), 2)
    env.close()
print(json.dumps({"envs": n, "ms_per_step": out}))
