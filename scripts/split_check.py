#!/usr/bin/env python
"""Two-warp (split) organisation against the one-thread organisation: bitwise comparison of obs / reward / done / state over
several decision steps with random actions (incl. mode 2) and auto-reset, then device-timed ms per step of both.
    python scripts/split_check.py [--envs 4096,8192] [--steps 6] [--stress]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="4096")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--time-steps", type=int, default=10)
ap.add_argument("--stress", action="store_true")
a = ap.parse_args()
kw = dict(use_j2=1, rw_set=1) if a.stress else {}
for n in [int(x) for x in a.envs.split(",")]:
    res = {}
    for org in ("thread", "split"):
        env = LeoPowerAttVecEnv(n, device=0, seed=7, auto_reset=True, max_length=4, **kw)
        env.set_organisation(org)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        acts = torch.randint(0, 3, (a.steps + a.time_steps + 3, n), dtype=torch.int32, device="cuda", generator=g)
        outs = []
        for t in range(a.steps):
            o, r, d, info = env.step(acts[t])
            outs.append((o.clone(), r.clone(), d.clone(), info["done_reason"].clone()))
        S, I = env.get_state()
        name = env.kernel_name()
        for t in range(3):
            env.step(acts[a.steps + t])
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.time_steps + 1)]
        ev[0].record()
        for t in range(a.time_steps):
            env.step(acts[a.steps + 3 + t]); ev[t + 1].record()
        torch.cuda.synchronize()
        ms = float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(a.time_steps)]))
        res[org] = (outs, S.clone(), I.clone(), name, ms)
        env.close()
    (o1, S1, I1, k1, ms1), (o2, S2, I2, k2, ms2) = res["thread"], res["split"]
    worst = 0.0
    equal = True
    for t in range(a.steps):
        for k in range(4):
            x, y = o1[t][k].double(), o2[t][k].double()
            same = bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())
            equal &= same
            if not same:
                worst = max(worst, float(((x - y).abs() / (x.abs() + 1e-300)).max()))
    sd = (S1 - S2).abs()
    st_equal = bool((S1 == S2).all()) and bool((I1 == I2).all())
    rel = float((sd / (S1.abs() + 1e-30)).max())
    dones = int(sum(int(o[2].sum()) for o in o1))
    print(json.dumps({"envs": n, "thread": k1, "split": k2, "outputs_bit_equal": equal, "worst_rel_output": worst, "state_bit_equal": st_equal,
                      "state_worst_rel": rel, "int_state_equal": bool((I1 == I2).all()), "episodes_ended": dones,
                      "ms_thread": ms1, "ms_split": ms2, "M_env_steps_s_split": n / ms2 / 1e3}))
