#!/usr/bin/env python
"""Action bucketing of the throughput organisation (bskenv.cu: LeoSched): device-timed A/B against lanes in index order on the
same envs and actions, bitwise comparison of the final state, and the single-action launch times that bound what bucketing can
give.    python scripts/bucket_probe.py [--envs 131072] [--steps 12]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="131072")
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--stress", type=int, default=1)
a = ap.parse_args()


def run(n, org, acts, kw, warm=3):
    env = LeoPowerAttVecEnv(n, device=0, seed=17, auto_reset=True, organisation=org, **kw)
    env.reset()
    for t in range(warm):
        env.step(acts[t])
    torch.cuda.synchronize()
    k = acts.shape[0] - warm
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    ev[0].record()
    chk = torch.zeros(3, dtype=torch.float64, device="cuda")
    for t in range(k):
        o, r, d, info = env.step(acts[warm + t]); ev[t + 1].record()
        chk += torch.stack([o.sum(), r.sum(), d.double().sum()])
    torch.cuda.synchronize()
    ms = float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(k)]))
    S, I = env.get_state()
    res = (ms, chk.clone(), S.clone(), I.clone(), env.kernel_name())
    env.close()
    return res


for n in [int(x) for x in a.envs.split(",")]:
    for kw in ([{}, dict(use_j2=1, rw_set=1)] if a.stress else [{}]):
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        acts = torch.randint(0, 3, (a.steps + 3, n), dtype=torch.int32, device="cuda", generator=g)
        b = run(n, "thread", acts, kw)
        i = run(n, "thread_index", acts, kw)
        single = {}
        for m in (0, 1, 2):
            single[m] = round(run(n, "thread_index", torch.full_like(acts, m), kw)[0], 3)
        print(json.dumps({"envs": n, "cfg": kw, "kernel": b[4], "ms_bucketed": round(b[0], 3), "ms_index_order": round(i[0], 3),
                          "M_env_steps_per_s": [round(n / b[0] / 1e3, 3), round(n / i[0] / 1e3, 3)],
                          "ms_single_action": single, "mean_single": round(sum(single.values()) / 3, 3),
                          "checksum_equal": bool(torch.equal(b[1], i[1])),
                          "state_equal": bool(torch.equal(b[2], i[2]) and torch.equal(b[3], i[3]))}), flush=True)
