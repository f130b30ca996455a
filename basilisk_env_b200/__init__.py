"""B200-native batched implementation of basilisk_env's LEO power/attitude environment step.

Public surface (mirrors /root/reference/basilisk_env/__init__.py and envs/__init__.py):
    make('leo_power_att_env-v0')     -> leoPowerAttEnv   (gym API, one spacecraft)
    LeoPowerAttVecEnv(num_envs, ...) -> N spacecraft per CUDA launch (device tensors or host buffers)
    LEOPowerAttitudeSimulator        -> the simulator object the env drives (run_sim / obs / ICs)
"""
import logging

from .registration import make, register, registered_ids
from . import spaces

logger = logging.getLogger(__name__)

register(id='leo_power_att_env-v0', entry_point='basilisk_env_b200.envs:leoPowerAttEnv')

__all__ = ["make", "register", "registered_ids", "spaces", "leoPowerAttEnv", "LEOPowerAttitudeSimulator",
           "LeoPowerAttVecEnv"]


def __getattr__(name):   # torch is only imported when an environment class is actually requested
    if name in ("leoPowerAttEnv", "LEOPowerAttitudeSimulator"):
        from . import envs
        return getattr(envs, name)
    if name in ("LeoPowerAttVecEnv", "shard_range", "fp64_peak_tflops", "BskEnvError"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
