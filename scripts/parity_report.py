#!/usr/bin/env python
"""Measured size of the CUDA-path-vs-oracle deviations (what the parity tests bound at 1e-9): a random subset of a large,
device-sampled batch that runs the shipped throughput path (queued work items, envs bucketed by action) is stepped beside the
oracle without re-synchronisation; discrete outcomes are asserted exact (tests/parity.py), the continuous deviations are
reported as maxima / percentiles per field and per interval.  Run on the GPU box:
    python scripts/parity_report.py > gpurun_out/parity_report.json
"Oracle" = the in-repo FP64 restatement of the Basilisk 1.x algorithms (PARITY UNPINNED)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
from basilisk_env_b200.opnav_env import OpNavVecEnv
from oracle import oracle as orc
from oracle import opnav as on
from tests import parity, opnav_parity

report = {}


def leo(name, n, sub, steps, **cfg):
    g = torch.Generator("cuda").manual_seed(41)
    acts = torch.randint(0, 3, (steps, n), dtype=torch.int32, device="cuda", generator=g)
    env = LeoPowerAttVecEnv(n, device=0, seed=77, **cfg)
    env.reset()
    ics = env.initial_conditions().cpu().numpy()
    idx = np.sort(np.random.RandomState(3).choice(n, sub, replace=False))
    ocfg = orc.default_cfg(**{k: v for k, v in cfg.items() if k in ("use_j2", "rw_set")})
    batch = orc.LeoEnvBatch(ics[idx], ocfg)
    a_host = acts.cpu().numpy()
    per_step, n_done, reasons = [], 0, set()
    alive = np.ones(sub, bool)                       # no auto-reset: an env is compared up to and including its terminal interval
    for t in range(steps):
        o, r, d, info = env.step(acts[t])
        ob, rew, dn, rs = o.cpu().numpy()[idx], r.cpu().numpy()[idx], d.cpu().numpy()[idx], info["done_reason"].cpu().numpy()[idx]
        S, I = env.get_state()
        S, I = S.cpu().numpy(), I.cpu().numpy()
        o_ob, o_rew, o_done, o_reason = batch.step(a_host[t, idx])
        errs = {}
        for k, e in enumerate(idx):
            if not alive[k]:
                continue
            where = f"{name} step {t} env {e}"
            parity.compare_obs(ob[k], o_ob[k], where)
            assert bool(dn[k]) == bool(o_done[k]) and int(rs[k]) == int(o_reason[k]), where
            assert abs(rew[k] - o_rew[k]) <= 1e-12, where
            for f, v in parity.compare_state(batch.envs[k].state(), S[:, e], I[:, e], where).items():
                errs.setdefault(f, []).append(v)
            if dn[k]:
                alive[k] = False; n_done += 1; reasons.add(int(rs[k]))
        per_step.append({f: {"max": float(np.max(v)), "p50": float(np.median(v))} for f, v in errs.items()})
    report[name] = {"envs_in_batch": n, "envs_compared": sub, "intervals": steps, "kernel": env.kernel_name(), "episodes_ended": n_done,
                    "done_reasons_seen": sorted(reasons), "discrete": "exact (flags, reasons, counters, masks, clock)",
                    "max_over_run": {f: max(s[f]["max"] for s in per_step if f in s) for f in per_step[0]},
                    "first_interval": per_step[0], "last_interval": per_step[-1]}
    env.close()


def opnav(name, n, steps):
    rows = opnav_parity.sample_rows(on, n, seed=21)
    acts = np.random.RandomState(22).randint(0, 2, size=(steps, n))
    env = OpNavVecEnv(n, device=0, first_env_index=5000, noise_seed=123, camera_reenable=1)
    batch = on.OpNavEnvBatch(rows, on.default_cfg(seed=123, camera_reenable=1), first_env_index=5000)
    env.reset_ics(rows)
    worst = {}
    for t in range(steps):
        o, r, d, info = env.step(acts[t])
        obs, dbg = o.cpu().numpy(), info["full_states"].cpu().numpy()
        o_ob, o_rew, o_done, o_reason, o_dbg = batch.step(acts[t])
        S, I = env.get_state()
        S, I = S.cpu().numpy(), I.cpu().numpy()
        for e, st in enumerate(batch.states()):
            where = f"{name} step {t} env {e}"
            opnav_parity.compare_obs(obs[e], o_ob[e], where)
            opnav_parity.compare_debug(dbg[e], o_dbg[e], where)
            opnav_parity.compare_state(st, S[:, e], I[:, e], where)
        np.testing.assert_array_equal(d.cpu().numpy().astype(bool), o_done)
        worst[f"obs_interval_{t}"] = float(np.max(np.abs(obs - o_ob) / np.maximum(np.abs(o_ob), 1.0)))
    report[name] = {"envs": n, "intervals": steps, "all_assertions_of_tests/opnav_parity.py": "passed",
                    "max_relative_obs_deviation": worst}
    env.close()


leo("leo_reference_60001_bucketed", 60001, 384, 8)
leo("leo_stress_60001_bucketed", 60001, 256, 6, use_j2=1, rw_set=1)
leo("leo_reference_4096_split", 4096, 256, 8)
opnav("opnav_768", 768, 3)
print(json.dumps(report, indent=1))
