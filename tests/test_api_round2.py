"""CPU suite for the round-2 API additions: gym.Env inheritance when a gym is importable, the stable-baselines-shaped
adapter, the new C-ABI entry points without a device, and the oracle's episode record (ENV:130-136)."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_env_classes_subclass_gym_env_when_gym_is_importable(tmp_path):
    """The reference declares `class leoPowerAttEnv(gym.Env)` (envs/leoPowerAttitudeEnvironment.py:14).  gym is not in this
    image, so a stand-in `gym` package on sys.path plays its part; without it the classes derive from object."""
    pkg = tmp_path / "gym"
    (pkg / "envs").mkdir(parents=True)
    (pkg / "__init__.py").write_text("class Env:\n    marker = 'fake gym'\n")
    (pkg / "envs" / "__init__.py").write_text("")
    (pkg / "envs" / "registration.py").write_text("registry = {}\ndef register(id, entry_point, **kw):\n    registry[id] = entry_point\n")
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import gym, gym.envs.registration as reg
        import basilisk_env_b200 as b
        from basilisk_env_b200.envs import leoPowerAttEnv
        from basilisk_env_b200.opnav_env import opNavEnv
        assert issubclass(leoPowerAttEnv, gym.Env) and issubclass(opNavEnv, gym.Env)
        assert reg.registry['leo_power_att_env-v0'] == 'basilisk_env_b200.envs:leoPowerAttEnv'
        print('ok')
    """ % (str(tmp_path), ROOT))
    out = subprocess.check_output([sys.executable, "-c", code], text=True)
    assert out.strip().endswith("ok")
    from basilisk_env_b200.envs import leoPowerAttEnv
    assert "gym" in sys.modules or leoPowerAttEnv.__mro__[1] is object


def test_sb_adapter_has_the_vecenv_surface_and_no_cpu_path():
    import torch
    from basilisk_env_b200 import sb_vec_env
    from basilisk_env_b200.vec_env import BskEnvError
    for name in ("reset", "step_async", "step_wait", "step", "close", "get_attr", "set_attr", "env_method", "seed",
                 "env_is_wrapped"):
        assert callable(getattr(sb_vec_env.LeoPowerAttSBVecEnv, name))
    if not torch.cuda.is_available():
        with pytest.raises(BskEnvError):
            sb_vec_env.LeoPowerAttSBVecEnv(4)


def test_host_entry_points_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from basilisk_env_b200 import _native
    L = _native.lib()
    p = C.c_void_p()
    assert L.bskenv_alloc_host(4096, C.byref(p)) == -2 and not p.value        # BSKENV_ECUDA, never a malloc fallback
    assert L.bskenv_free_host(None) == 0
    assert L.bskenv_step_host_wait(None) == -1 and L.bskenv_step_host_async(None, *([None] * 8)) == -1
    assert L.bskenv_step_info(None, *([None] * 9)) == -1


def test_oracle_episode_record(orc):
    """orc_env_episode restates ENV:130-136: 'r' = reward_total, 'l' = curr_step before its increment; with max_length = L
    the (L + 1)-th call ends the episode (quirk Q9) with l = L."""
    L = 3
    row = orc.ic_to_row(orc.sample_ic_dict(np.random.RandomState(2)))
    env = orc.LeoEnv(orc.default_cfg(step_duration=10.0), max_length=L)
    env.reset(row)
    total = 0.0
    for t in range(L + 1):
        ob, rew, done, reason = env.step(0)
        total += rew
        r, l = env.episode()
        assert l == t and abs(r - total) <= 1e-15
        assert done == (t == L)
    assert reason == 1 and 0.0 < total <= (L + 1) / L
