#!/usr/bin/env python
"""SURVEY 8(f)-4 cost: ms per decision step of the LEO kernel with (a) the reference force model, (b) inertial zonal J2,
(c) the planet-fixed degree-2 field, (d) the same plus Sun / orientation tables.  One GPU, CUDA events, i.i.d. actions.

    python scripts/degree2_bench.py [n_envs] > gpurun_out/degree2_bench.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basilisk_env_b200 import ephemeris as eph                      # noqa: E402
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
out = {"envs": n, "cases": {}}
sun = eph.ChebTable.fit(eph.analytic_sun, 0.0, 86400.0, 2, 9)
orient = eph.ChebTable.fit(eph.iau_earth_angles, 0.0, 86400.0, 2, 5)
for name, kw, deg2, tables in (("reference", {}, False, False), ("use_j2", {"use_j2": 1}, False, False),
                               ("degree2_pfix", {}, True, False), ("degree2_pfix_tables", {}, True, True),
                               ("stress_4rw_j2", {"use_j2": 1, "rw_set": 1}, False, False),
                               ("stress_4rw_degree2_tables", {"rw_set": 1}, True, True)):
    env = LeoPowerAttVecEnv(n, device=0, seed=3, auto_reset=True, **kw)
    if deg2:
        env.set_gravity_degree2(True)
    if tables:
        env.set_ephemeris("sun", sun); env.set_ephemeris("orientation", orient)
    env.reset()
    g = torch.Generator("cuda").manual_seed(7)
    acts = torch.randint(0, 3, (13, n), dtype=torch.int32, device="cuda", generator=g)
    for t in range(3):
        env.step(acts[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(3, 13):
        env.step(acts[t])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["cases"][name] = {"ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "flop_per_env_step": env.flops_per_step()}
    env.close()
print(json.dumps(out))
