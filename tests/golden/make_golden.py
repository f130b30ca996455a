"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/bsk_oracle.c).

PARITY UNPINNED: these vectors are outputs of the in-repo FP64 restatement, not of Basilisk (which
cannot be built or imported in this image; the reference ships no recorded trajectories).  Inputs
follow the reference's own demos: initial conditions drawn from numpy's legacy stream in the
reference's order (np.random.seed(12345), the seed of ENV:225), a constant-0 action sequence
(ENV:226-227, SIM:672) for a full 541-call episode, and a recorded random action sequence.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402


def episode(ic_row, actions):
    env = orc.LeoEnv()
    ob0 = env.reset(ic_row)
    obs, rew, done, reason = [], [], [], []
    for a in actions:
        o, r, d, why = env.step(int(a))
        obs.append(o); rew.append(r); done.append(d); reason.append(why)
        if d:
            break
    st = env.state()
    n = len(obs)
    return dict(ic=ic_row, ob0=ob0, actions=np.asarray(actions[:n], np.int32), obs=np.array(obs), reward=np.array(rew),
                done=np.array(done), reason=np.array(reason, np.int32), final_r=np.array(st.r_BN_N[:]),
                final_v=np.array(st.v_BN_N[:]), final_sigma=np.array(st.sigma_BN[:]), final_Omega=np.array(st.Omega[:3]),
                final_switch=np.int64(st.mrp_switch_count), final_fire=np.array(st.thr_fire_count[:], np.int64))


def main():
    rng = np.random.RandomState(12345)
    ic = orc.ic_to_row(orc.sample_ic_dict(rng))
    np.savez(os.path.join(HERE, "leo_episode_const0.npz"), **episode(ic, np.zeros(541, np.int32)))
    ic2 = orc.ic_to_row(orc.sample_ic_dict(rng))
    acts = np.random.RandomState(777).randint(0, 3, 541).astype(np.int32)
    np.savez(os.path.join(HERE, "leo_episode_random.npz"), **episode(ic2, acts))
    # a small multi-env fixture: 16 envs x 6 steps, all three modes
    rows = np.stack([orc.ic_to_row(orc.sample_ic_dict(rng)) for _ in range(16)])
    rows[:4, 15:18] *= 3.0                                  # fast wheels: mode 2 fires thrusters
    a = np.random.RandomState(778).randint(0, 3, size=(6, 16)).astype(np.int32)
    batch = orc.LeoEnvBatch(rows)
    obs, rew, done, reason = [], [], [], []
    for t in range(6):
        o, r, d, w = batch.step(a[t])
        obs.append(o); rew.append(r); done.append(d); reason.append(w)
    fire = np.array([e.state().thr_fire_count[:] for e in batch.envs], np.int64)
    np.savez(os.path.join(HERE, "leo_batch16.npz"), ics=rows, ob0=batch.obs0, actions=a, obs=np.array(obs), reward=np.array(rew),
             done=np.array(done), reason=np.array(reason, np.int32), fire=fire)


if __name__ == "__main__":
    main()
