"""TEST INFRASTRUCTURE: ingest of a recorded Basilisk trajectory (docs/TRACE_SCHEMA.md).

PARITY IS UNPINNED in this repository because Basilisk cannot be built or run in the build image.  Anybody who has a
Basilisk 1.x installation can pin it with two commands:

    python scripts/record_basilisk_trace.py --out trace.npz          (on the Basilisk machine; uses the reference's own classes)
    python scripts/compare_basilisk_trace.py trace.npz --backend both (here; `gpu` needs a B200, `oracle` runs anywhere)

The trace carries the initial conditions, the action list, the per-decision-step message samples the reference's
`run_sim` pulls (simulators/leoPowerAttitudeSimulator.py:598-619, plus velocity and attitude from the same logged message)
and the Sun state of every SPICE tick, which replaces this repository's analytic Sun (deviation D1) through
`bskenv_set_ephemeris` / `orc_set_ephemeris`.  This module replays the trace through the CUDA path (C ABI) and / or the CPU
oracle and reports the deviations at the tolerances of tests/parity.py."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCHEMA_VERSION = 1
PER_STEP = {"obs": 5, "r_BN_N": 3, "v_BN_N": 3, "sigma_BN": 3, "omega_BN_B": 3, "wheelSpeeds": 3, "sigma_BR": 3, "sigma_RN": 3,
            "storageLevel": 0, "shadowFactor": 0}
# field -> (tolerance, kind): rel = |d| / max(|ref|, floor); abs = |d|
TOL = {"r_BN_N": (1e-9, "rel", 0.0), "v_BN_N": (1e-9, "rel", 0.0), "sigma_BN": (1e-9, "rel", 1.0),
       "omega_BN_B": (1e-9, "rel", 1e-3), "wheelSpeeds": (1e-9, "rel", 1.0), "sigma_BR": (1e-9, "rel", 1.0),
       "sigma_RN": (1e-9, "rel", 1.0), "storageLevel": (1e-9, "rel", 1.0), "shadowFactor": (1e-7, "abs", 0.0),
       "obs": (1e-7, "abs", 0.0)}


class TraceError(ValueError):
    pass


def load_trace(path):
    """Read and validate a trace file; returns a dict of numpy arrays (docs/TRACE_SCHEMA.md)."""
    z = np.load(path, allow_pickle=False)
    t = {k: z[k] for k in z.files}
    need = ["schema_version", "ic", "actions", "dynRate", "fswRate", "step_duration", "sun_r", "sun_v"] + list(PER_STEP)
    missing = [k for k in need if k not in t]
    if missing:
        raise TraceError(f"{path}: missing arrays {missing}")
    if int(t["schema_version"]) != SCHEMA_VERSION:
        raise TraceError(f"{path}: schema_version {int(t['schema_version'])}, this tool reads {SCHEMA_VERSION}")
    T = int(np.asarray(t["actions"]).shape[0])
    if np.asarray(t["ic"]).shape != (19,):
        raise TraceError("ic must be 19 doubles in BSKENV_IC_DIM order (include/bskenv.h)")
    for k, w in PER_STEP.items():
        want = (T, w) if w else (T,)
        if tuple(np.asarray(t[k]).shape) != want:
            raise TraceError(f"{k} must have shape {want}, found {tuple(np.asarray(t[k]).shape)}")
    for k in ("sun_r", "sun_v"):
        if tuple(np.asarray(t[k]).shape) != (T + 1, 3):
            raise TraceError(f"{k} must have shape {(T + 1, 3)}: one row per SPICE tick t = 0, step_duration, ..., T * step_duration")
    return t


def save_trace(path, **arrays):
    arrays.setdefault("schema_version", np.int32(SCHEMA_VERSION))
    np.savez(path, **arrays)


def sun_table(trace):
    from basilisk_env_b200.ephemeris import ChebTable
    return ChebTable.from_nodes(0.0, float(trace["step_duration"]), trace["sun_r"], trace["sun_v"])


def _cfg_kw(trace, extra):
    kw = dict(dynRate=float(trace["dynRate"]), fswRate=float(trace["fswRate"]), step_duration=float(trace["step_duration"]))
    kw.update(extra)
    return kw


def run_oracle(trace, **cfg):
    """Replay through oracle/bsk_oracle.c with the trace's Sun table; returns the same per-step arrays as the trace."""
    from oracle import oracle as orc
    orc.set_ephemeris(0, sun_table(trace))
    try:
        sim = orc.LeoSim(np.asarray(trace["ic"], float), orc.default_cfg(**_cfg_kw(trace, cfg)))
        out = {k: [] for k in PER_STEP}
        for a in np.asarray(trace["actions"]):
            ob, _ = sim.run_sim(int(a))
            st = sim.state()
            out["obs"].append(ob)
            for k, f in (("r_BN_N", st.r_BN_N), ("v_BN_N", st.v_BN_N), ("sigma_BN", st.sigma_BN), ("omega_BN_B", st.omega_BN_B),
                         ("sigma_BR", st.sigma_BR), ("sigma_RN", st.sigma_RN)):
                out[k].append(np.array(f[:]))
            out["wheelSpeeds"].append(np.array(st.Omega[:3]))
            out["storageLevel"].append(st.storedCharge)
            out["shadowFactor"].append(st.shadowFactor)
        return {k: np.asarray(v) for k, v in out.items()}
    finally:
        orc.set_ephemeris(0, None)


def run_gpu(trace, device=0, **cfg):
    """Replay through the CUDA path (C ABI of include/bskenv.h: bskenv_set_ephemeris, bskenv_reset_ics, bskenv_step)."""
    import torch
    from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
    from basilisk_env_b200 import _native
    F = lambda name: _native.state_field(name)[0]       # noqa: E731
    T = len(trace["actions"])
    kw = _cfg_kw(trace, cfg)
    kw.setdefault("max_length", max(T, 1))
    env = LeoPowerAttVecEnv(1, device=device, **kw)
    env.set_ephemeris("sun", sun_table(trace).extended(int(kw["max_length"]) + 1))
    env.reset_ics(np.asarray(trace["ic"], float).reshape(1, 19))
    out = {k: [] for k in PER_STEP}
    for a in np.asarray(trace["actions"]):
        env.step(torch.tensor([int(a)], dtype=torch.int32, device=env.device))
        d, _ = env.get_state()
        S = d.cpu().numpy()[:, 0]
        out["obs"].append(S[F("sim_obs"):F("sim_obs") + 5])
        for k, name in (("r_BN_N", "r_BN_N"), ("v_BN_N", "v_BN_N"), ("sigma_BN", "sigma_BN"), ("omega_BN_B", "omega_BN_B"),
                        ("wheelSpeeds", "Omega"), ("sigma_BR", "att_guidance"), ("sigma_RN", "att_reference")):
            out[k].append(S[F(name):F(name) + 3])
        out["storageLevel"].append(S[F("storedCharge")])
        out["shadowFactor"].append(S[F("shadowFactor")])
    env.close()
    return {k: np.asarray(v) for k, v in out.items()}


def compare(trace, got):
    """Per field: worst deviation over the trace, the step where it occurs and whether it is within tolerance."""
    rep = {}
    for k, (tol, kind, floor) in TOL.items():
        ref, g = np.asarray(trace[k], float), np.asarray(got[k], float)
        if ref.ndim == 1:
            ref, g = ref[:, None], g[:, None]
        if k == "obs":      # |sigma_BR|, |omega| , |Omega|, Wh, shadow: relative on the first four, absolute on the shadow factor
            d = np.abs(g - ref) / np.maximum(np.abs(ref), [1.0, 1e-3, 1.0, 1e-3, 1.0])
            d[:, 4] = np.abs(g[:, 4] - ref[:, 4])
            err = d.max(axis=1)
            tol_eff = tol
        elif kind == "rel":
            err = np.linalg.norm(g - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), max(floor, 1e-300))
            tol_eff = tol
        else:
            err = np.abs(g - ref).max(axis=1)
            tol_eff = tol
        i = int(np.argmax(err))
        rep[k] = {"worst": float(err[i]), "step": i, "tol": tol_eff, "ok": bool(err[i] <= tol_eff)}
    rep["ok"] = all(v["ok"] for v in rep.values())
    return rep


def record_oracle_trace(ic_row, actions, path=None, **cfg):
    """A trace in the file format, produced by the oracle with the analytic Sun: the fixture that proves the tool chain
    (tests/test_trace_tool.py) and an example of what scripts/record_basilisk_trace.py writes."""
    from oracle import oracle as orc
    kw = dict(dynRate=0.1, fswRate=1.0, step_duration=180.0)
    kw.update(cfg)
    sim = orc.LeoSim(np.asarray(ic_row, float), orc.default_cfg(**kw))
    r0, v0, et = np.zeros(3), np.zeros(3), orc.C.c_double(0.0)
    orc.lib().orc_sun_ephemeris(0.0, orc._p(r0), orc._p(v0), orc.C.byref(et))
    out = {k: [] for k in PER_STEP}
    sun_r, sun_v = [r0], [v0]
    for a in actions:
        ob, _ = sim.run_sim(int(a))
        st = sim.state()
        out["obs"].append(ob)
        for k, f in (("r_BN_N", st.r_BN_N), ("v_BN_N", st.v_BN_N), ("sigma_BN", st.sigma_BN), ("omega_BN_B", st.omega_BN_B),
                     ("sigma_BR", st.sigma_BR), ("sigma_RN", st.sigma_RN)):
            out[k].append(np.array(f[:]))
        out["wheelSpeeds"].append(np.array(st.Omega[:3]))
        out["storageLevel"].append(st.storedCharge)
        out["shadowFactor"].append(st.shadowFactor)
        sun_r.append(np.array(st.sun_r[:])); sun_v.append(np.array(st.sun_v[:]))
    tr = {k: np.asarray(v) for k, v in out.items()}
    tr.update(schema_version=np.int32(SCHEMA_VERSION), ic=np.asarray(ic_row, float), actions=np.asarray(actions, np.int32),
              dynRate=np.float64(kw["dynRate"]), fswRate=np.float64(kw["fswRate"]), step_duration=np.float64(kw["step_duration"]),
              sun_r=np.asarray(sun_r), sun_v=np.asarray(sun_v), source=np.str_("oracle/bsk_oracle.c (NOT Basilisk)"))
    if path:
        np.savez(path, **tr)
    return tr


def main(argv=None):
    ap = argparse.ArgumentParser(description="Replay a recorded Basilisk trajectory (docs/TRACE_SCHEMA.md) and report the deviations.")
    ap.add_argument("trace")
    ap.add_argument("--backend", default="oracle", choices=["oracle", "gpu", "both"])
    ap.add_argument("--hill-cel-pun", type=int, default=0, help="quirk decision D3 (DESIGN.md section 9)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    trace = load_trace(a.trace)
    cfg = {"hill_cel_pun": a.hill_cel_pun}
    reports = {}
    if a.backend in ("oracle", "both"):
        reports["oracle"] = compare(trace, run_oracle(trace, **cfg))
    if a.backend in ("gpu", "both"):
        reports["gpu"] = compare(trace, run_gpu(trace, **cfg))
    src = str(trace.get("source", "unknown"))
    print(f"trace {a.trace}: {len(trace['actions'])} decision steps of {float(trace['step_duration'])} s, source: {src}")
    for name, rep in reports.items():
        print(f"[{name}] {'WITHIN TOLERANCE' if rep['ok'] else 'DEVIATES'}")
        for k, v in rep.items():
            if isinstance(v, dict):
                print(f"   {k:14s} worst {v['worst']:.3e} at step {v['step']:4d}   tol {v['tol']:.0e}   {'ok' if v['ok'] else 'FAIL'}")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(reports, f, indent=1)
    return 0 if all(r["ok"] for r in reports.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
