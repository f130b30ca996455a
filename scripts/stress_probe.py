#!/usr/bin/env python
"""Times the stress configuration (BASELINE configs[4]: J2 + four wheels + dumping) at several batch sizes.
    python scripts/stress_probe.py [--envs 56832,65536,113664]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="56832,65536,113664")
a = ap.parse_args()
out = []
for n in [int(x) for x in a.envs.split(",")]:
    env = LeoPowerAttVecEnv(n, device=0, seed=17, auto_reset=True, use_j2=1, rw_set=1)
    env.reset()
    acts = torch.randint(0, 3, (13, n), dtype=torch.int32, device="cuda")
    for t in range(3):
        env.step(acts[t])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for t in range(10):
        env.step(acts[3 + t]); ev[t + 1].record()
    torch.cuda.synchronize()
    ms = float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(10)]))
    out.append((n, round(ms, 3), round(n / ms / 1e3, 3), env.kernel_name()))
    env.close()
print(out)
