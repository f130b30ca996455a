#!/usr/bin/env python
"""Record one trajectory of the REFERENCE simulator (needs AVS-Lab Basilisk 1.x and atharris/basilisk_env installed; NOT
runnable in this repository's build image, which has neither) in the format of docs/TRACE_SCHEMA.md.

    python scripts/record_basilisk_trace.py --out trace.npz [--steps 20] [--seed 7] [--actions 0,1,2,0,...]

It drives the reference's own classes and pulls exactly the messages `run_sim` pulls:
    LEOPowerAttitudeSimulator(dynRate, fswRate, step_duration, initial_conditions)   simulators/leoPowerAttitudeSimulator.py:67
    .run_sim(action)                                                                :535-644
    .pullMessageLogData(<msg>.<field>, range)                                       as at :598-619
The only addition is one more logged message, `sun_planet_data` (the reference has that line commented out at :509), added
by overriding `set_logging`, so that the Sun state of every SPICE tick is in the trace: replaying it through
`bskenv_set_ephemeris` removes this repository's analytic-Sun deviation from the comparison.
Then, anywhere:  python scripts/compare_basilisk_trace.py trace.npz --backend both"""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--actions", default=None, help="comma-separated list; default: seeded uniform {0,1,2}")
    a = ap.parse_args()

    from Basilisk.utilities import macros as mc                                   # noqa: F401  (reference: :13)
    from basilisk_env.simulators import leoPowerAttitudeSimulator as sim_mod

    class RecordingSimulator(sim_mod.LEOPowerAttitudeSimulator):
        def set_logging(self):
            sim_mod.LEOPowerAttitudeSimulator.set_logging(self)
            self.TotalSim.logThisMessage("sun_planet_data", mc.sec2nano(self.step_duration))     # cf. :509

    np.random.seed(a.seed)                       # the reference samples its ICs from numpy's global legacy stream (:119-193)
    sim = RecordingSimulator(0.1, 1.0, 180.0)
    ic = sim.initial_conditions
    row = np.concatenate([np.asarray(ic["rN"], float).reshape(3), np.asarray(ic["vN"], float).reshape(3),
                          np.asarray(ic["sigma_init"], float).reshape(3), np.asarray(ic["omega_init"], float).reshape(3),
                          np.asarray(ic["disturbance_vector"], float).reshape(3), np.asarray(ic["wheelSpeeds"], float).reshape(3),
                          [float(ic["storedCharge_Init"])]])
    if a.actions:
        actions = np.array([int(x) for x in a.actions.split(",")], dtype=np.int32)
    else:
        actions = np.random.RandomState(a.seed + 1).randint(0, 3, a.steps).astype(np.int32)

    sc, nav_att = sim.scObject.scStateOutMsgName, sim.simpleNavObject.outputAttName
    pulls = {"r_BN_N": (sc + ".r_BN_N", 3), "v_BN_N": (sc + ".v_BN_N", 3), "sigma_BN": (sc + ".sigma_BN", 3),
             "omega_BN_B": (nav_att + ".omega_BN_B", 3), "wheelSpeeds": (sim.rwStateEffector.OutputDataString + ".wheelSpeeds", 3),
             "sigma_BR": (sim.trackingErrorData.outputDataName + ".sigma_BR", 3), "sigma_RN": ("att_reference.sigma_RN", 3),
             "storageLevel": (sim.powerMonitor.batPowerOutMsgName + ".storageLevel", 1),
             "shadowFactor": (sim.solarPanel.sunEclipseInMsgName + ".shadowFactor", 1)}
    out = {k: [] for k in pulls}
    out["obs"] = []
    for act in actions:
        obs, _, _ = sim.run_sim(int(act))
        out["obs"].append(np.asarray(obs, float).reshape(5))
        for k, (name, width) in pulls.items():
            last = sim.pullMessageLogData(name, list(range(width)))[-1, 1:1 + width]   # newest sample = this decision boundary
            out[k].append(np.asarray(last, float) if width > 1 else float(last[0]))
    sun_r = sim.pullMessageLogData("sun_planet_data.PositionVector", list(range(3)))
    sun_v = sim.pullMessageLogData("sun_planet_data.VelocityVector", list(range(3)))
    T = len(actions)
    assert sun_r.shape[0] >= T + 1, "expected one Sun sample per SPICE tick (t = 0, 180 s, ...)"
    try:
        import Basilisk
        version = str(getattr(Basilisk, "__version__", "unknown"))
    except Exception:
        version = "unknown"
    np.savez(a.out, schema_version=np.int32(1), ic=row, actions=actions, dynRate=np.float64(0.1), fswRate=np.float64(1.0),
             step_duration=np.float64(180.0), sun_r=sun_r[:T + 1, 1:4], sun_v=sun_v[:T + 1, 1:4],
             source=np.str_("Basilisk " + version + " via atharris/basilisk_env LEOPowerAttitudeSimulator"),
             **{k: np.asarray(v) for k, v in out.items()})
    sim.close_gracefully()
    print("wrote", a.out, "-", T, "decision steps")


if __name__ == "__main__":
    main()
