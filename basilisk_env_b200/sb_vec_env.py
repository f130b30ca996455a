"""stable-baselines-shaped `VecEnv` over the batched step.

What an RL trainer written against the reference would call: the reference env is a single-process gym.Env whose `step`
returns `info['episode'] = {'r': reward_total, 'l': curr_step}` when the episode ends
(/root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:130-142); trainers of that generation wrap N copies in
a `SubprocVecEnv`.  This class has that interface -- `reset`, `step_async`, `step_wait`, `step`, `close`, `get_attr`,
`set_attr`, `env_method`, `seed`, numpy in / numpy out, list-of-dict infos with `episode` and `terminal_observation`
for the envs that finished, automatic reset of finished envs -- but all N envs advance in ONE CUDA launch through
`bskenv_step_host_async` / `_wait` with page-locked zero-copy buffers.  When `stable_baselines3` (or the older
`stable_baselines`) is importable the class also inherits its `VecEnv`, so `isinstance` checks of its wrappers pass."""
import numpy as np

from . import spaces
from .vec_env import LeoPowerAttVecEnv, OBS_DIM

_Base = object
for _mod in ("stable_baselines3.common.vec_env", "stable_baselines.common.vec_env"):
    try:
        _Base = __import__(_mod, fromlist=["VecEnv"]).VecEnv
        break
    except Exception:       # not installed
        pass


class LeoPowerAttSBVecEnv(_Base):
    """N LEO power/attitude envs behind the stable-baselines VecEnv interface (observations [N, 5, 1] float64, as the
    reference's Box(shape=(5, 1)) observation space, ENV:45)."""

    metadata = {"render.modes": []}

    def __init__(self, num_envs, device=0, seed=0, first_env_index=0, **config):
        self.vec = LeoPowerAttVecEnv(num_envs, device=device, first_env_index=first_env_index, seed=seed, auto_reset=True,
                                     **config)
        self.num_envs = int(num_envs)
        self.observation_space = spaces.Box(-1e16, 1e16, shape=(OBS_DIM, 1))
        self.action_space = spaces.Discrete(3)
        self._actions, self._out = self.vec.host_buffers(episode=True)
        self._waiting = False
        self.buf_infos = [{} for _ in range(self.num_envs)]

    # ---- VecEnv interface ---------------------------------------------------------------------
    def reset(self):
        obs = self.vec.reset()
        return obs.cpu().numpy().reshape(self.num_envs, OBS_DIM, 1)

    def seed(self, seed=None):
        if seed is not None:
            self.vec.seed_value = int(seed)
        return [self.vec.seed_value + i for i in range(self.num_envs)]

    def step_async(self, actions):
        if self._waiting:
            raise RuntimeError("step_async called twice without step_wait")
        np.copyto(self._actions, np.asarray(actions).reshape(self.num_envs), casting="unsafe")
        self.vec.step_host_async(self._actions, self._out)
        self._waiting = True

    def step_wait(self):
        if not self._waiting:
            raise RuntimeError("step_wait without step_async")
        self.vec.step_host_wait()
        self._waiting = False
        obs, rew, done, reason, ep_r, ep_l, term = self._out
        dones = done.astype(bool)
        infos = [{} for _ in range(self.num_envs)]
        idx = np.flatnonzero(dones)
        if idx.size:
            # the kernel reset these envs in the same launch: `obs` already holds the first observation of the new
            # episode, `term` the last one of the old
            for e in idx:
                infos[e] = {"episode": {"r": float(ep_r[e]), "l": int(ep_l[e])},
                            "terminal_observation": term[e].reshape(OBS_DIM, 1).copy(), "done_reason": int(reason[e])}
        self.buf_infos = infos
        return obs.reshape(self.num_envs, OBS_DIM, 1).copy(), rew.copy(), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if self._waiting:
            self.vec.step_host_wait()
            self._waiting = False
        self.vec.close()

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.vec, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.vec, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.vec, method_name)(*method_args, **method_kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def get_images(self):
        return []

    def render(self, mode="human"):
        return None
