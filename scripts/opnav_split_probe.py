#!/usr/bin/env python
"""opNav: three-kernel interval (noise walk -> buffer -> dynamics -> filter) against the fused first pass (the default; the three-kernel form is opt-in: BSKENV_OPNAV_NOISE_SPLIT=1)
on the same envs and actions: device-timed medians and a bitwise comparison of outputs and final state.
    python scripts/opnav_split_probe.py [--envs 113664,75776]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.opnav_env import OpNavVecEnv
ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="113664")
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
for n in [int(x) for x in a.envs.split(",")]:
    res = {}
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = torch.randint(0, 2, (a.steps + 2, n), dtype=torch.int32, device="cuda", generator=g)
    for name in ("fused", "split"):
        os.environ["BSKENV_OPNAV_NOISE_SPLIT"] = "0" if name == "fused" else "1"
        env = OpNavVecEnv(n, device=0, auto_reset=True, sample_orbit=1, camera_reenable=1, noise_seed=1)
        env.reset(seed=1)
        for t in range(2):
            env.step(acts[t])
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        chk = torch.zeros(3, dtype=torch.float64, device="cuda")
        l0 = env.launch_count()
        ev[0].record()
        for t in range(a.steps):
            out = env.step(acts[2 + t]); ev[t + 1].record()
            chk += torch.stack([out[0].sum(), out[1].sum(), out[2].double().sum()])
        torch.cuda.synchronize()
        ms = float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(a.steps)]))
        S, I = env.get_state()
        res[name] = (ms, chk.clone(), S.clone(), I.clone(), (env.launch_count() - l0) // a.steps)
        env.close()
    f, s = res["fused"], res["split"]
    print(json.dumps({"envs": n, "ms_fused": round(f[0], 2), "ms_split": round(s[0], 2), "M_env_steps_per_s": [round(n / f[0] / 1e3, 4), round(n / s[0] / 1e3, 4)],
                      "kernels_per_step": [f[4], s[4]], "checksum_equal": bool(torch.equal(f[1], s[1])),
                      "state_equal": bool(torch.equal(f[2], s[2]) and torch.equal(f[3], s[3]))}), flush=True)
