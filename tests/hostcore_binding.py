"""TEST INFRASTRUCTURE: ctypes binding of tests/hostcore/hostcore.cpp, i.e. the device core
(basilisk_env_b200/csrc/leo_core.cuh) compiled for the host with g++ -ffp-contract=off.
Used by the CPU test-suite to check the fused schedule against the independent oracle; the product
never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

from basilisk_env_b200._native import Config

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcore")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libhostcore.so")
    deps = [os.path.join(_HERE, "hostcore.cpp")] + [os.path.join(_ROOT, "basilisk_env_b200", "csrc", f)
                                                     for f in ("leo_core.cuh", "leo_params.h", "leo_host.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, deps[0]])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.hc_default_config.argtypes = [C.POINTER(Config)]
        L.hc_create.restype = vp
        L.hc_create.argtypes = [C.POINTER(Config), C.c_int64]
        L.hc_destroy.argtypes = [vp]
        L.hc_reset_ics.argtypes = [vp, vp, vp]
        L.hc_reset_seeded.argtypes = [vp, C.c_uint64, C.c_int64, vp, vp]
        L.hc_step.argtypes = [vp] * 6
        L.hc_get_state.argtypes = [vp] * 3
        L.hc_dims.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hc_thr_force_mapping.argtypes = [vp, vp, vp]
        L.hc_force_general.argtypes = [vp, C.c_int]
        L.hc_eclipse.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, vp]
        _LIB = L
    return _LIB


def default_config(**kw):
    cfg = Config()
    lib().hc_default_config(C.byref(cfg))
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            v = (C.c_double * len(v))(*v)
        setattr(cfg, k, v)
    return cfg


class HostCore:
    def __init__(self, n, **kw):
        self.L = lib()
        self.n = n
        self.cfg = default_config(**kw)
        self.h = self.L.hc_create(C.byref(self.cfg), n)
        assert self.h, "hc_create rejected the configuration"
        nd, ni = C.c_int(), C.c_int()
        self.L.hc_dims(C.byref(nd), C.byref(ni))
        self.nd, self.ni = nd.value, ni.value

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hc_destroy(self.h)
            self.h = None

    def reset_ics(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        obs = np.zeros((self.n, 5))
        self.L.hc_reset_ics(self.h, rows.ctypes.data, obs.ctypes.data)
        return obs

    def reset_seeded(self, seed, first_env=0):
        ics = np.zeros((self.n, 19))
        obs = np.zeros((self.n, 5))
        self.L.hc_reset_seeded(self.h, seed, first_env, ics.ctypes.data, obs.ctypes.data)
        return ics, obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, 5)); rew = np.zeros(self.n)
        done = np.zeros(self.n, np.uint8); reason = np.zeros(self.n, np.uint8)
        self.L.hc_step(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, reason.ctypes.data)
        return obs, rew, done.astype(bool), reason.astype(np.int32)

    def state(self):
        S = np.zeros((self.nd, self.n)); I = np.zeros((self.ni, self.n), np.int64)
        self.L.hc_get_state(self.h, S.ctypes.data, I.ctypes.data)
        return S, I

    def force_general(self, on=True):
        """Run the general (non-DIAG) EOM path even for the reference configuration."""
        self.L.hc_force_general(self.h, int(on))

    def eclipse(self, msg_ns, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.zeros(len(r)); sun = np.zeros(3)
        self.L.hc_eclipse(self.h, int(msg_ns), r.ctypes.data, len(r), out.ctypes.data, sun.ctypes.data)
        return out, sun

    def thr_force_mapping(self, Lr):
        Lr = np.ascontiguousarray(Lr, dtype=np.float64); F = np.zeros(8)
        self.L.hc_thr_force_mapping(self.h, Lr.ctypes.data, F.ctypes.data)
        return F
