"""`gym.make`-style registry.

The reference registers one id (/root/reference/basilisk_env/__init__.py:6-9):
    register(id='leo_power_att_env-v0', entry_point='basilisk_env.envs:leoPowerAttEnv')
and leaves 'opnav_env-v0' commented out (:11-14).  `gym`/`gymnasium` are not installed in the build
image, so `make` resolves ids from the in-package registry; if a real gym IS importable the same ids
are also registered there so `gym.make('leo_power_att_env-v0')` keeps working."""
import importlib

_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)
    for modname in ("gym", "gymnasium"):
        try:
            mod = importlib.import_module(modname + ".envs.registration")
            mod.register(id=id, entry_point=entry_point, **kwargs)
        except Exception:   # not installed, or already registered
            pass


def registered_ids():
    return sorted(_REGISTRY)


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f"No registered env with id: {id} (known: {registered_ids()})")
    entry_point, reg_kwargs = _REGISTRY[id]
    modname, _, attr = entry_point.partition(":")
    cls = getattr(importlib.import_module(modname), attr)
    kw = dict(reg_kwargs.get("kwargs", {}))
    kw.update(kwargs)
    return cls(**kw)
