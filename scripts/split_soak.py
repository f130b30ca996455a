#!/usr/bin/env python
"""Soak run of the split organisation against the one-thread organisation (and, with --pair thread,thread_index --envs 60001,
of the action-bucketed throughput organisation against lanes in index order): many decision steps with random actions, episodes
ending for every reason, auto-reset inside the launch; bitwise comparison of the final state and of a running checksum of the
outputs.  python scripts/split_soak.py [--steps 300]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--pair", default="thread,split")
ap.add_argument("--envs", default="")
a = ap.parse_args()
A, B = a.pair.split(",")
CASES = ((4096, {}), (4000, {}), (16384, {}), (4096, dict(use_j2=1, rw_set=1)))
if a.envs:
    CASES = tuple((int(x), kw) for x in a.envs.split(",") for kw in ({}, dict(use_j2=1, rw_set=1)))
for n, kw in CASES:
    res = {}
    for org in (A, B):
        env = LeoPowerAttVecEnv(n, device=0, seed=23, auto_reset=True, organisation=org, **kw)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(11)
        chk = torch.zeros(4, dtype=torch.float64, device="cuda")
        reasons = 0
        for t in range(a.steps):
            o, r, d, info = env.step(torch.randint(0, 3, (n,), dtype=torch.int32, device="cuda", generator=g))
            chk += torch.stack([o.sum(), r.sum(), d.double().sum(), info["done_reason"].double().sum()])
            reasons |= int(torch.bitwise_or(info["done_reason"][d.bool()].to(torch.int32), torch.zeros((), dtype=torch.int32, device="cuda")).unique().sum()) if bool(d.any()) else 0
        S, I = env.get_state()
        res[org] = (chk.clone(), S.clone(), I.clone(), env.kernel_name(), env.episode_stats())
        env.close()
    t_, s_ = res[A], res[B]
    print(json.dumps({"envs": n, "cfg": kw, "steps": a.steps, "kernels": [t_[3], s_[3]], "checksum_equal": bool(torch.equal(t_[0], s_[0])),
                      "state_equal": bool(torch.equal(t_[1], s_[1]) and torch.equal(t_[2], s_[2])), "episodes": t_[4]["episodes"],
                      "ends": {k: t_[4][k] for k in ("wheel_failures", "power_failures", "orbit_decays", "max_length_ends")},
                      "counts_equal": all(t_[4][k] == s_[4][k] for k in ("episodes", "wheel_failures", "power_failures", "orbit_decays", "max_length_ends", "env_steps"))}))
