"""Generates the opNav golden fixture of tests/golden/ from the CPU oracle (oracle/opnav_oracle.c).

PARITY UNPINNED: these vectors are outputs of the in-repo FP64 restatement, not of Basilisk (which cannot be
built or imported in this image; the reference ships no recorded trajectories).  Inputs follow the reference:
env 0 flies the fixed orbit of simulators/opNavSimulator.py:173-178 with the filter error drawn from numpy's legacy
stream (:187-188) and is driven by the reference's own recorded action list `actHist = [1,1,0,0,1,1,1,0,0,1]`
(:327); the other envs use orbits from the commented-out element ranges (:166-171) and seeded random actions.

    python tests/golden/make_golden_opnav.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import opnav as on  # noqa: E402

ACT_HIST = [1, 1, 0, 0, 1, 1, 1, 0, 0, 1]
SEED = 2019
FIRST_ENV = 7


def main():
    rng = np.random.RandomState(12345)
    n = 8
    rows = np.stack([on.sample_ic_row(rng, sample_orbit=(k > 0)) for k in range(n)])
    acts = np.random.RandomState(778).randint(0, 2, size=(len(ACT_HIST), n)).astype(np.int32)
    acts[:, 0] = ACT_HIST
    out = {}
    for tag, kw in (("ref", dict()), ("cam", dict(camera_reenable=1))):
        batch = on.OpNavEnvBatch(rows, on.default_cfg(seed=SEED, **kw), first_env_index=FIRST_ENV)
        obs, rew, done, reason, dbg, nmeas = [], [], [], [], [], []
        for t in range(len(ACT_HIST)):
            o, r, d, w, g = batch.step(acts[t])
            obs.append(o); rew.append(r); done.append(d); reason.append(w); dbg.append(g)
            nmeas.append([s.n_meas for s in batch.states()])
        sts = batch.states()
        out.update({f"{tag}_obs": np.array(obs), f"{tag}_reward": np.array(rew), f"{tag}_done": np.array(done),
                    f"{tag}_reason": np.array(reason, np.int32), f"{tag}_debug": np.array(dbg),
                    f"{tag}_n_meas": np.array(nmeas, np.int64),
                    f"{tag}_filt_state": np.array([s.filt_state[:] for s in sts]),
                    f"{tag}_filt_covar": np.array([s.filt_covar[:] for s in sts]),
                    f"{tag}_Omega": np.array([s.Omega[:] for s in sts])})
    np.savez(os.path.join(HERE, "opnav_batch8.npz"), ics=rows, actions=acts, seed=np.int64(SEED), first_env=np.int64(FIRST_ENV), **out)


if __name__ == "__main__":
    main()
