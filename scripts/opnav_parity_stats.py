"""Development aid: distribution of kernel-vs-oracle differences of the opNav path over a batch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from basilisk_env_b200.opnav_env import OpNavVecEnv
from oracle import opnav as on
from tests import opnav_parity as par

n, steps = int(os.environ.get("N", 1024)), int(os.environ.get("STEPS", 3))
rows = par.sample_rows(on, n, seed=21)
acts = np.random.RandomState(22).randint(0, 2, size=(steps, n))
env = OpNavVecEnv(n, device=0, first_env_index=5000, noise_seed=123, camera_reenable=1)
batch = on.OpNavEnvBatch(rows, on.default_cfg(seed=123, camera_reenable=1), first_env_index=5000)
env.reset_ics(rows)
F = par.F
for t in range(steps):
    env.step(acts[t]); batch.step(acts[t])
    d, i = env.get_state(); S, I = d.cpu().numpy(), i.cpu().numpy()
    sts = batch.states()
    for mode in (0, 1):
        idx = [e for e in range(n) if sts[e].mode == mode]
        if not idx:
            continue
        def mx(f):
            return max(f(e) for e in idx)
        print(f"step {t} mode {mode} ({len(idx)} envs):",
              "r %.1e" % mx(lambda e: par.rel(S[F('r_BN_N'):F('r_BN_N') + 3, e], sts[e].r_BN_N[:])),
              "sigma %.1e" % mx(lambda e: np.abs(S[F('sigma_BN'):F('sigma_BN') + 3, e] - np.array(sts[e].sigma_BN[:])).max()),
              "omega_abs %.1e" % mx(lambda e: np.abs(S[F('omega_BN_B'):F('omega_BN_B') + 3, e] - np.array(sts[e].omega_BN_B[:])).max()),
              "Omega_abs %.1e" % mx(lambda e: np.abs(S[F('Omega'):F('Omega') + 4, e] - np.array(sts[e].Omega[:])).max()),
              "|Omega| %.1f" % mx(lambda e: np.linalg.norm(sts[e].Omega[:])),
              "rwCmd_abs %.1e" % mx(lambda e: np.abs(S[F('reactionwheel_cmds'):F('reactionwheel_cmds') + 4, e] - np.array(sts[e].rwCmd[:])).max()),
              "filt_r %.1e" % mx(lambda e: par.rel(S[F('filter_state'):F('filter_state') + 3, e], sts[e].filt_state[:3])),
              "sBR %.1e" % mx(lambda e: np.linalg.norm(sts[e].sigma_BR[:])))
