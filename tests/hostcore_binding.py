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
        L.hc_set_gravity_degree2.argtypes = [vp, C.c_int, vp]
        L.hc_set_ephemeris.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, vp]
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

    def set_gravity_degree2(self, enable=True, cbar=None):
        c = None if cbar is None else np.ascontiguousarray(cbar, dtype=np.float64)
        self.L.hc_set_gravity_degree2(self.h, int(enable), None if c is None else c.ctypes.data)

    def set_ephemeris(self, kind, table):
        if table is None:
            self.L.hc_set_ephemeris(self.h, kind, 0.0, 1.0, 0, 1, None)
            return
        coef = np.ascontiguousarray(table.coef, dtype=np.float64)
        self.L.hc_set_ephemeris(self.h, kind, float(table.t0), float(table.seg_len), coef.shape[0], coef.shape[2], coef.ctypes.data)

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


# --------------------------------------------------------------------------------------------------
# opNav device core (basilisk_env_b200/csrc/opnav_core.cuh) compiled for the host
# --------------------------------------------------------------------------------------------------
_LIB_ON = None


def build_opnav(force=False):
    so = os.path.join(_HERE, "libhostcore_opnav.so")
    deps = [os.path.join(_HERE, "hostcore_opnav.cpp")] + [os.path.join(_ROOT, "basilisk_env_b200", "csrc", f) for f in
                                                           ("opnav_core.cuh", "opnav_params.h", "opnav_host.h", "leo_core.cuh",
                                                            "leo_params.h", "leo_host.h")]
    deps.append(os.path.join(_ROOT, "include", "bskenv.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                               "-o", so, deps[0]])
    return so


def lib_opnav():
    global _LIB_ON
    if _LIB_ON is None:
        from basilisk_env_b200._native import OpNavConfig
        L = C.CDLL(build_opnav())
        vp = C.c_void_p
        L.hco_default_config.argtypes = [C.POINTER(OpNavConfig)]
        L.hco_create.restype = vp
        L.hco_create.argtypes = [C.POINTER(OpNavConfig), C.c_int64, C.c_int64]
        L.hco_destroy.argtypes = [vp]
        L.hco_reset_ics.argtypes = [vp, vp, vp]
        L.hco_reset_seeded.argtypes = [vp, C.c_uint64, vp, vp]
        L.hco_step.argtypes = [vp] * 7
        L.hco_get_state.argtypes = [vp] * 3
        L.hco_dims.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hco_set_ephemeris.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int, vp]
        L.hco_normals.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.hco_sun.argtypes = [vp, C.c_double, vp, vp]
        L.hco_eclipse.restype = C.c_double
        L.hco_eclipse.argtypes = [vp, vp, vp]
        L.hco_ukf_time_update.argtypes = [vp, vp, vp, vp, C.c_double]
        L.hco_ukf_meas_update.restype = C.c_int
        L.hco_ukf_meas_update.argtypes = [vp, vp, vp, vp, C.c_double, vp, vp]
        _LIB_ON = L
    return _LIB_ON


class HostCoreOpNav:
    def __init__(self, n, first_env=0, **kw):
        from basilisk_env_b200._native import OpNavConfig
        self.L = lib_opnav()
        self.n = n
        self.cfg = OpNavConfig()
        self.L.hco_default_config(C.byref(self.cfg))
        for k, v in kw.items():
            setattr(self.cfg, k, v)
        self.h = self.L.hco_create(C.byref(self.cfg), n, first_env)
        assert self.h, "hco_create rejected the configuration"
        nd, ni = C.c_int(), C.c_int()
        self.L.hco_dims(C.byref(nd), C.byref(ni))
        self.nd, self.ni = nd.value, ni.value

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hco_destroy(self.h)
            self.h = None

    def reset_ics(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        obs = np.zeros((self.n, 4))
        self.L.hco_reset_ics(self.h, rows.ctypes.data, obs.ctypes.data)
        return obs

    def set_ephemeris(self, table):
        if table is None:
            self.L.hco_set_ephemeris(self.h, 0.0, 1.0, 0, 1, None)
            return
        coef = np.ascontiguousarray(table.coef, dtype=np.float64)
        self.L.hco_set_ephemeris(self.h, float(table.t0), float(table.seg_len), coef.shape[0], coef.shape[2], coef.ctypes.data)

    def reset_seeded(self, seed):
        ics = np.zeros((self.n, 12)); obs = np.zeros((self.n, 4))
        self.L.hco_reset_seeded(self.h, seed, ics.ctypes.data, obs.ctypes.data)
        return ics, obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, 4)); rew = np.zeros(self.n); dbg = np.zeros((self.n, 12))
        done = np.zeros(self.n, np.uint8); reason = np.zeros(self.n, np.uint8)
        self.L.hco_step(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, reason.ctypes.data,
                        dbg.ctypes.data)
        return obs, rew, done.astype(bool), reason.astype(np.int32), dbg

    def state(self):
        S = np.zeros((self.nd, self.n)); I = np.zeros((self.ni, self.n), np.int64)
        self.L.hco_get_state(self.h, S.ctypes.data, I.ctypes.data)
        return S, I

    def normals(self, env, episode, tick, stream, block):
        out = np.zeros(4)
        self.L.hco_normals(self.h, env, episode, tick, stream, block, out.ctypes.data)
        return out

    def sun(self, t):
        r = np.zeros(3); v = np.zeros(3)
        self.L.hco_sun(self.h, float(t), r.ctypes.data, v.ctypes.data)
        return r, v

    def eclipse(self, sun, r):
        sun = np.ascontiguousarray(sun, dtype=np.float64); r = np.ascontiguousarray(r, dtype=np.float64)
        return float(self.L.hco_eclipse(self.h, sun.ctypes.data, r.ctypes.data))

    def ukf_time_update(self, x, S21, dt):
        x = np.array(x, dtype=np.float64); S21 = np.array(S21, dtype=np.float64); m = np.zeros(6)
        self.L.hco_ukf_time_update(self.h, x.ctypes.data, S21.ctypes.data, m.ctypes.data, float(dt))
        return x, S21, m

    def ukf_meas_update(self, x, S21, m, dt, obs, R6):
        x = np.array(x, dtype=np.float64); S21 = np.array(S21, dtype=np.float64)
        m = np.ascontiguousarray(m, dtype=np.float64); obs = np.ascontiguousarray(obs, dtype=np.float64)
        R6 = np.ascontiguousarray(R6, dtype=np.float64)
        ok = self.L.hco_ukf_meas_update(self.h, x.ctypes.data, S21.ctypes.data, m.ctypes.data, float(dt), obs.ctypes.data,
                                        R6.ctypes.data)
        return x, S21, bool(ok)
