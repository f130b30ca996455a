"""TEST INFRASTRUCTURE: steps a batch through whichever build of the library BSKENV_LIB names and dumps what the eclipse
comparison of tests/test_gpu_round2.py needs (run in a subprocess: one process loads one build of libbskenv.so).

    BSKENV_LIB=basilisk_env_b200/libbskenv_literal.so python -m tests.eclipse_probe out.npz [n_envs] [steps] [step_duration]"""
import sys

import numpy as np
import torch

from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
from basilisk_env_b200 import _native


def main():
    out, n, steps, dur = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    env = LeoPowerAttVecEnv(n, device=0, seed=123, step_duration=dur, max_length=10 ** 6)
    env.reset()
    ics = env.initial_conditions().cpu().numpy()
    acts = np.random.RandomState(7).randint(0, 3, size=(steps, n)).astype(np.int32)
    obs4 = np.zeros((steps, n))
    rr = np.zeros((steps, n, 3))
    F = _native.state_field("r_BN_N")[0]
    for t in range(steps):
        o, r, d, info = env.step(torch.as_tensor(acts[t], device="cuda"))
        obs4[t] = o[:, 4].cpu().numpy()
        rr[t] = env.get_state()[0][F:F + 3].cpu().numpy().T
    np.savez(out, ics=ics, acts=acts, obs4=obs4, r=rr, lib=np.str_(_native.lib_path()))
    env.close()


if __name__ == "__main__":
    main()
