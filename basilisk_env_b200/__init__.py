"""B200-native batched implementation of basilisk_env's LEO power/attitude environment step.

Public surface (mirrors /root/reference/basilisk_env/__init__.py and envs/__init__.py):
    make('leo_power_att_env-v0')     -> leoPowerAttEnv   (gym API, one spacecraft)
    LeoPowerAttVecEnv(num_envs, ...) -> N spacecraft per CUDA launch (device tensors or host buffers)
    LEOPowerAttitudeSimulator        -> the simulator object the env drives (run_sim / obs / ICs)
    opNavEnv / scenario_OpNav / OpNavVecEnv -> the opNav env (dynamics half + nav measurement model + OD filter)
"""
import logging

from .registration import make, register, registered_ids
from . import spaces

logger = logging.getLogger(__name__)

register(id='leo_power_att_env-v0', entry_point='basilisk_env_b200.envs:leoPowerAttEnv')

# 'opnav_env-v0' stays unregistered, as in the reference (/root/reference/basilisk_env/__init__.py:11-14 is commented
# out); the class itself is importable (envs/__init__.py:2): basilisk_env_b200.opNavEnv / basilisk_env_b200.envs.opNavEnv

__all__ = ["make", "register", "registered_ids", "spaces", "leoPowerAttEnv", "LEOPowerAttitudeSimulator",
           "LeoPowerAttVecEnv", "opNavEnv", "scenario_OpNav", "OpNavVecEnv"]


def __getattr__(name):   # torch is only imported when an environment class is actually requested
    if name in ("leoPowerAttEnv", "LEOPowerAttitudeSimulator"):
        from . import envs
        return getattr(envs, name)
    if name in ("opNavEnv", "scenario_OpNav", "OpNavVecEnv"):
        from . import opnav_env
        return getattr(opnav_env, name)
    if name in ("LeoPowerAttVecEnv", "shard_range", "fp64_peak_tflops", "BskEnvError"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
