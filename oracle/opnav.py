"""ctypes binding of the opNav CPU oracle (oracle/opnav_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (Basilisk is not available; see opnav_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C

import numpy as np

from . import oracle as _o

MU_MARS = 4.2828371901284001E+13
MU_MARS_FSW = 42828.314 * 1E9    # the filter's gravitational parameter ([BSK astroConstants MU_MARS] * 1e9)
IC_DIM = 12          # rN(3) vN(3) rError(3) vError(3): the C-ABI's opNav IC row


class OpNavIC(C.Structure):
    _fields_ = [("rN", C.c_double * 3), ("vN", C.c_double * 3), ("rError", C.c_double * 3), ("vError", C.c_double * 3)]


class OpNavCfg(C.Structure):
    _fields_ = [("dynRate", C.c_double), ("fswRate", C.c_double), ("step_duration_min", C.c_double),
                ("nav_noise", C.c_int), ("camera_reenable", C.c_int), ("pixel_noise_std", C.c_double),
                ("circle_unc", C.c_double), ("seed", C.c_uint64), ("numModes", C.c_int), ("reserved", C.c_int * 3)]


class OpNavState(C.Structure):
    _fields_ = [("r_BN_N", C.c_double * 3), ("v_BN_N", C.c_double * 3), ("sigma_BN", C.c_double * 3),
                ("omega_BN_B", C.c_double * 3), ("Omega", C.c_double * 4), ("u_current", C.c_double * 4),
                ("navErrors", C.c_double * 18),
                ("nav_r", C.c_double * 3), ("nav_v", C.c_double * 3), ("nav_sigma", C.c_double * 3),
                ("nav_omega", C.c_double * 3), ("nav_sun_B", C.c_double * 3),
                ("sigma_BR", C.c_double * 3), ("omega_BR_B", C.c_double * 3), ("Lr", C.c_double * 3), ("rwCmd", C.c_double * 4),
                ("css", C.c_double * 8), ("sun_point", C.c_double * 3), ("shadow", C.c_double),
                ("filt_state", C.c_double * 6), ("filt_covar", C.c_double * 36), ("filt_sBar", C.c_double * 36),
                ("filt_time", C.c_double),
                ("meas_r", C.c_double * 3), ("meas_covar", C.c_double * 9), ("circle", C.c_double * 3),
                ("n_meas", C.c_int64), ("n_bad", C.c_int64), ("n_images", C.c_int64), ("mrp_switch_count", C.c_int64),
                ("camera_on", C.c_int32), ("mode", C.c_int32), ("modeCounter", C.c_int32), ("pad", C.c_int32),
                ("sim_nanos", C.c_uint64)]


class OpNavOut(C.Structure):
    _fields_ = [("ob", C.c_double * 4), ("debug", C.c_double * 12), ("reward", C.c_double), ("done", C.c_int),
                ("reason", C.c_int)]


class Ukf(C.Structure):
    _fields_ = [("state", C.c_double * 6), ("sBar", C.c_double * 36), ("covar", C.c_double * 36), ("xBar", C.c_double * 6),
                ("SP", C.c_double * 78), ("timeTag", C.c_double), ("wM", C.c_double * 13), ("wC", C.c_double * 13),
                ("gamma", C.c_double), ("sQnoise", C.c_double * 36), ("mu", C.c_double), ("noiseSF", C.c_double),
                ("n_bad", C.c_int64)]


_BOUND = set()


def lib(fast=False):
    """fast=True: the -O3 -march=native host build of the same sources (bench.py's optimised CPU baseline)."""
    L = _o.fast_lib() if fast else _o.lib()
    if id(L) not in _BOUND:
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.orc_opnav_default_cfg.argtypes = [C.POINTER(OpNavCfg)]
        L.orc_opnav_reference_orbit.argtypes = [C.POINTER(OpNavIC)]
        L.orc_opnav_create.restype = vp
        L.orc_opnav_create.argtypes = [C.POINTER(OpNavIC), C.POINTER(OpNavCfg), C.c_uint64, C.c_uint64]
        L.orc_opnav_destroy.argtypes = [vp]
        L.orc_opnav_run_sim.restype = C.c_int
        L.orc_opnav_run_sim.argtypes = [vp, C.c_int, dp, dp]
        L.orc_opnav_get_state.argtypes = [vp, C.POINTER(OpNavState)]
        L.orc_opnav_env_create.restype = vp
        L.orc_opnav_env_create.argtypes = [C.POINTER(OpNavCfg)]
        L.orc_opnav_env_destroy.argtypes = [vp]
        L.orc_opnav_env_reset.argtypes = [vp, C.POINTER(OpNavIC), C.c_uint64, C.c_uint64, dp]
        L.orc_opnav_env_step.argtypes = [vp, C.c_int, C.POINTER(OpNavOut)]
        L.orc_opnav_env_sim.restype = vp
        L.orc_opnav_env_sim.argtypes = [vp]
        L.orc_opnav_env_set_max_length.argtypes = [vp, C.c_int]
        L.orc_opnav_env_episode.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_opnav_env_step_batch.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int), C.POINTER(OpNavOut), C.c_int]
        L.orc_opnav_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.orc_opnav_project_circle.restype = C.c_int
        L.orc_opnav_project_circle.argtypes = [dp, C.c_double, dp]
        L.orc_opnav_pixel_line.argtypes = [dp, C.c_double, dp, dp, dp]
        L.orc_ukf_qr_just_r.argtypes = [dp, C.c_int, C.c_int, dp]
        L.orc_ukf_chol_downdate.restype = C.c_int
        L.orc_ukf_chol_downdate.argtypes = [dp, dp, C.c_double, C.c_int, dp]
        L.orc_ukf_chol_decomp.restype = C.c_int
        L.orc_ukf_chol_decomp.argtypes = [dp, C.c_int, dp]
        L.orc_ukf_state_prop.argtypes = [dp, C.c_double, C.c_double]
        L.orc_ukf_init.argtypes = [C.POINTER(Ukf), dp, dp, dp, C.c_double, C.c_double]
        L.orc_ukf_time_update.argtypes = [C.POINTER(Ukf), C.c_double]
        L.orc_ukf_meas_update.argtypes = [C.POINTER(Ukf), dp, dp]
        L.orc_sun_from_mars.argtypes = [C.c_double, dp, dp, dp]
        _BOUND.add(id(L))
    return L


_p = _o._p


def default_cfg(**kw):
    cfg = OpNavCfg()
    lib().orc_opnav_default_cfg(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def reference_orbit():
    """rN, vN of the reference's fixed Mars orbit (simulators/opNavSimulator.py:173-181)."""
    ic = OpNavIC()
    lib().orc_opnav_reference_orbit(C.byref(ic))
    return np.array(ic.rN[:]), np.array(ic.vN[:])


def sample_ic_row(rng=None, sample_orbit=False):
    """One IC row [rN vN rError vError] in the reference's draw order (simulators/opNavSimulator.py:163-188):
    the orbit is fixed (the random element draws are commented out there, :166-171; `sample_orbit` enables them),
    then rError = uniform(100000,-100000,3), vError = uniform(1000,-1000,3) from numpy's legacy stream."""
    R = np.random if rng is None else rng
    if sample_orbit:
        a = float(R.uniform(17000 * 1E3, 22000 * 1E3, 1)[0]); e = float(R.uniform(0, 0.6, 1)[0])
        i = float(R.uniform(-20 * _o.D2R, 20 * _o.D2R, 1)[0]); Om = float(R.uniform(0, 360 * _o.D2R, 1)[0])
        om = float(R.uniform(0, 360 * _o.D2R, 1)[0]); f = float(R.uniform(0, 360 * _o.D2R, 1)[0])
        rN, vN = _o.elem2rv(MU_MARS, a, e, i, Om, om, f)
    else:
        rN, vN = reference_orbit()
    rErr = R.uniform(100000, -100000, 3)
    vErr = R.uniform(1000, -1000, 3)
    return np.concatenate([rN, vN, rErr, vErr])


def ic_from_row(row):
    row = np.asarray(row, dtype=np.float64)
    ic = OpNavIC()
    ic.rN = (C.c_double * 3)(*row[0:3]); ic.vN = (C.c_double * 3)(*row[3:6])
    ic.rError = (C.c_double * 3)(*row[6:9]); ic.vError = (C.c_double * 3)(*row[9:12])
    return ic


class OpNavSim:
    """One scalar oracle sim == one scenario_OpNav(1., 1., 50.)."""

    def __init__(self, row, cfg=None, env_index=0, episode=0):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_cfg()
        self.ic = ic_from_row(row)
        self._h = self._L.orc_opnav_create(C.byref(self.ic), C.byref(self.cfg), int(env_index), int(episode))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_opnav_destroy(self._h)
            self._h = None

    def run_sim(self, action):
        o, d = np.zeros(4), np.zeros(12)
        over = self._L.orc_opnav_run_sim(self._h, int(action), _p(o), _p(d))
        return o, d, bool(over)

    def state(self):
        st = OpNavState()
        self._L.orc_opnav_get_state(self._h, C.byref(st))
        return st


class OpNavEnv:
    """Oracle restatement of opNavEnv.reset/step for one env."""

    def __init__(self, cfg=None, L=None, max_length=None):
        self._L = L if L is not None else lib()
        self.cfg = cfg if cfg is not None else default_cfg()
        self._h = self._L.orc_opnav_env_create(C.byref(self.cfg))
        if max_length is not None:
            self._L.orc_opnav_env_set_max_length(self._h, int(max_length))

    def episode(self):
        """(reward_total, curr_step) as the reference puts them into info['episode'] (opNavEnvironment.py:106-109)."""
        r, l = C.c_double(0.0), C.c_int(0)
        self._L.orc_opnav_env_episode(self._h, C.byref(r), C.byref(l))
        return r.value, l.value

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_opnav_env_destroy(self._h)
            self._h = None

    def reset(self, row, env_index=0, episode=0):
        ic = ic_from_row(row)
        ob = np.zeros(4)
        self._L.orc_opnav_env_reset(self._h, C.byref(ic), int(env_index), int(episode), _p(ob))
        return ob

    def step(self, action):
        out = OpNavOut()
        self._L.orc_opnav_env_step(self._h, int(action), C.byref(out))
        return np.array(out.ob[:]), out.reward, bool(out.done), out.reason, np.array(out.debug[:])

    def state(self):
        st = OpNavState()
        self._L.orc_opnav_get_state(self._L.orc_opnav_env_sim(self._h), C.byref(st))
        return st


class OpNavEnvBatch:
    """n independent oracle envs stepped with OpenMP over envs (the CPU baseline)."""

    def __init__(self, ic_rows, cfg=None, first_env_index=0, L=None, max_length=None):
        self._L = L if L is not None else lib()
        self.cfg = cfg if cfg is not None else default_cfg()
        self.n = len(ic_rows)
        self.envs = [OpNavEnv(self.cfg, self._L, max_length) for _ in range(self.n)]
        self.obs0 = np.stack([e.reset(r, first_env_index + k, 0) for k, (e, r) in enumerate(zip(self.envs, ic_rows))])
        self._handles = (C.c_void_p * self.n)(*[e._h for e in self.envs])
        self._outs = (OpNavOut * self.n)()

    def step(self, actions, nthreads=0):
        acts = (C.c_int * self.n)(*[int(a) for a in actions])
        self._L.orc_opnav_env_step_batch(self._handles, self.n, acts, self._outs, int(nthreads))
        ob = np.array([o.ob[:] for o in self._outs])
        rew = np.array([o.reward for o in self._outs])
        done = np.array([o.done for o in self._outs], dtype=bool)
        reason = np.array([o.reason for o in self._outs], dtype=np.int32)
        dbg = np.array([o.debug[:] for o in self._outs])
        return ob, rew, done, reason, dbg

    def states(self):
        return [e.state() for e in self.envs]
