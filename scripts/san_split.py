"""compute-sanitizer --tool synccheck workload: the split organisation only (named barriers with divergent-prone code between
them), whole and ragged groups, mode 2, fresh + running envs in one warp, both configurations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
for n, kw in ((64, {}), (200, {}), (33, dict(use_j2=1, rw_set=1))):
    env = LeoPowerAttVecEnv(n, device=0, auto_reset=True, step_duration=20.0, max_length=2, seed=8, organisation="split", **kw)
    env.reset()
    for t in range(4):
        env.step(torch.randint(0, 3, (n,), dtype=torch.int32, device="cuda"))
    print("leo split", n, env.kernel_name(), env.episode_stats()["episodes"]); env.close()
