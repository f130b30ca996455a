#!/usr/bin/env python
"""BASELINE config 5: FP32-vs-FP64 accuracy / throughput trade-off of the LEO step kernel on one GPU.

For the reference configuration and the stress configuration (J2 + drag + eclipse + four wheels with momentum dumping),
65536 envs: device-timed ms per decision step of the FP64 kernel and of the mixed-precision kernel (precision = 1: FP32
stage arithmetic, FP64 state accumulation / clocks / flight software / events), and the differences between the two
trajectories (same initial conditions, same actions) after 1, 5 and 20 decision intervals.  Prints one JSON object.

    python scripts/mixed_tradeoff.py [--envs 65536] [--steps 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
from basilisk_env_b200 import _native


def F(name):
    return _native.state_field(name)[0]


def run(cfg_kw, n, steps, report_at):
    envs = {p: LeoPowerAttVecEnv(n, device=0, seed=77, precision=p, **cfg_kw) for p in (0, 1)}
    for e in envs.values():
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 3, (steps, n), dtype=torch.int32, device="cuda", generator=g)
    ms, diffs = {}, []
    for p, e in envs.items():
        e.step(acts[0]); torch.cuda.synchronize()          # first interval (1801 ticks) + warm-up, not timed
    for p, e in envs.items():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        st = {}
        per = []
        for t in range(1, steps):
            ev[0].record(); e.step(acts[t]); ev[1].record(); torch.cuda.synchronize()
            per.append(ev[0].elapsed_time(ev[1]))
            if (t + 1) in report_at:
                d, i = e.get_state()
                st[t + 1] = (d.clone(), i.clone(), e.obs.clone(), e.done.clone())
        ms[p] = float(np.median(per))
        envs[p].snap = st
    for k in report_at:
        if k not in envs[0].snap:
            continue
        d0, i0, o0, dn0 = envs[0].snap[k]; d1, i1, o1, dn1 = envs[1].snap[k]
        def q(x):
            x = x.flatten().double().cpu().numpy()
            return {"median": float(np.median(x)), "p99": float(np.percentile(x, 99)), "max": float(x.max())}
        r0 = d0[F("r_BN_N"):F("r_BN_N") + 3]; r1 = d1[F("r_BN_N"):F("r_BN_N") + 3]
        v0 = d0[F("v_BN_N"):F("v_BN_N") + 3]; v1 = d1[F("v_BN_N"):F("v_BN_N") + 3]
        same = (dn0 == dn1)
        diffs.append({"after_steps": k,
                      "position_m": q((r0 - r1).norm(dim=0)), "velocity_m_s": q((v0 - v1).norm(dim=0)),
                      "sigma_BN": q((d0[F("sigma_BN"):F("sigma_BN") + 3] - d1[F("sigma_BN"):F("sigma_BN") + 3]).abs().max(dim=0).values),
                      "omega_rad_s": q((d0[F("omega_BN_B"):F("omega_BN_B") + 3] - d1[F("omega_BN_B"):F("omega_BN_B") + 3]).abs().max(dim=0).values),
                      "wheel_rad_s": q((d0[F("Omega"):F("Omega") + 4] - d1[F("Omega"):F("Omega") + 4]).abs().max(dim=0).values),
                      "charge_J": q((d0[F("storedCharge")] - d1[F("storedCharge")]).abs()),
                      "obs_abs": [q((o0[:, c] - o1[:, c]).abs()) for c in range(5)],
                      "done_flags_equal_frac": float(same.double().mean()),
                      "fire_counters_equal_frac": float((i0[F("fireCounter"):F("fireCounter") + 8] == i1[F("fireCounter"):F("fireCounter") + 8]).all(dim=0).double().mean())})
    for e in envs.values():
        e.close()
    return {"ms_per_step": {"fp64": ms[0], "mixed": ms[1]}, "env_steps_per_s": {"fp64": n / ms[0] * 1e3, "mixed": n / ms[1] * 1e3},
            "speedup": ms[0] / ms[1], "differences": diffs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    out = {"envs": a.envs, "steps": a.steps,
           "reference_config": run({}, a.envs, a.steps, (2, 5, a.steps)),
           "stress_config": run({"use_j2": 1, "rw_set": 1}, a.envs, a.steps, (2, 5, a.steps))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
