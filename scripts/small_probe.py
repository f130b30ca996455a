#!/usr/bin/env python
"""Times the LEO step at small batch sizes (BASELINE configs[1] = 4096 envs): device-timed ms per step, steady state.
    python scripts/small_probe.py [--envs 4096,8192,...] [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="4096")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
out = []
for n in [int(x) for x in a.envs.split(",")]:
    env = LeoPowerAttVecEnv(n, device=0, seed=5, auto_reset=True)
    env.reset()
    acts = torch.randint(0, 3, (a.steps + a.warmup, n), dtype=torch.int32, device="cuda")
    for t in range(a.warmup):
        env.step(acts[t])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for t in range(a.steps):
        env.step(acts[a.warmup + t]); ev[t + 1].record()
    torch.cuda.synchronize()
    per = [ev[t].elapsed_time(ev[t + 1]) for t in range(a.steps)]
    ms = float(np.median(per))
    out.append({"envs": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "kernel": env.kernel_name(), "checksum": float(env.obs.sum())})
    env.close()
print(json.dumps(out))
