"""CPU suite: the oracle's physics blocks against closed forms / invariants.

PARITY UNPINNED: the reference ships no golden outputs and Basilisk cannot be built here, so the
oracle itself is pinned against first principles (two-body invariants, rigid-body
and wheel momentum exchange, MRP identities, eclipse geometry, battery clamp) instead."""
import ctypes as C

import numpy as np
import pytest

from tests import parity

P = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(P)


def test_elem2rv_against_vis_viva_and_package(orc):
    from basilisk_env_b200 import initial_conditions as icm
    rng = np.random.RandomState(0)
    for _ in range(50):
        a = 6871e3; e = rng.uniform(0, 0.05); i = rng.uniform(-np.pi / 2, np.pi / 2)
        Om, om, f = rng.uniform(0, 2 * np.pi, 3)
        r, v = orc.elem2rv(orc.MU_EARTH, a, e, i, Om, om, f)
        # vis-viva and angular momentum
        assert abs(np.dot(v, v) - orc.MU_EARTH * (2 / np.linalg.norm(r) - 1 / a)) < 1e-6
        h = np.cross(r, v)
        assert abs(np.linalg.norm(h) - np.sqrt(orc.MU_EARTH * a * (1 - e * e))) < 1e-3
        assert abs(h[2] / np.linalg.norm(h) - np.cos(i)) < 1e-12
        r2, v2 = icm.elem2rv(icm.MU_EARTH, icm.ClassicElements(a, e, i, Om, om, f))
        np.testing.assert_allclose(r2, r, rtol=1e-14, atol=1e-7)
        np.testing.assert_allclose(v2, v, rtol=1e-14, atol=1e-10)


def test_mrp_identities(orc):
    L = orc.lib()
    rng = np.random.RandomState(1)
    for _ in range(100):
        q = rng.uniform(-0.6, 0.6, 3)
        Cm = np.zeros((3, 3)); L.orc_MRP2C(_p(q), _p(Cm))
        np.testing.assert_allclose(Cm @ Cm.T, np.eye(3), atol=1e-14)
        assert abs(np.linalg.det(Cm) - 1) < 1e-14
        q2 = np.zeros(3); L.orc_C2MRP(_p(Cm), _p(q2))
        np.testing.assert_allclose(q2, q, atol=1e-14)
        # shadow set maps to the same DCM
        qs = -q / np.dot(q, q); Cs = np.zeros((3, 3)); L.orc_MRP2C(_p(qs), _p(Cs))
        np.testing.assert_allclose(Cs, Cm, atol=1e-13)
        # add / sub are inverse; composition matches DCM products
        p = rng.uniform(-0.5, 0.5, 3)
        s = np.zeros(3); L.orc_addMRP(_p(q), _p(p), _p(s))
        Cp = np.zeros((3, 3)); L.orc_MRP2C(_p(p), _p(Cp))
        Csum = np.zeros((3, 3)); L.orc_MRP2C(_p(s), _p(Csum))
        np.testing.assert_allclose(Csum, Cp @ Cm, atol=1e-13)
        back = np.zeros(3); L.orc_subMRP(_p(s), _p(q), _p(back))
        np.testing.assert_allclose(back, p, atol=1e-13)


def test_sun_ephemeris_is_plausible_for_the_epoch(orc):
    """2021 May 4: Sun ~1.0083 AU away, declination ~ +16 deg, right ascension ~ 41.5 deg; v = dr/dt."""
    r = np.zeros(3); v = np.zeros(3); et = C.c_double()
    orc.lib().orc_sun_ephemeris(0.0, _p(r), _p(v), C.byref(et))
    AU = 149597870.693e3
    d = np.linalg.norm(r)
    assert 1.0075 < d / AU < 1.0090
    dec = np.degrees(np.arcsin(r[2] / d)); ra = np.degrees(np.arctan2(r[1], r[0]))
    assert 15.5 < dec < 16.5 and 40.5 < ra < 42.5
    r2 = np.zeros(3); v2 = np.zeros(3)
    orc.lib().orc_sun_ephemeris(200.0, _p(r2), _p(v2), None)
    np.testing.assert_allclose((r2 - r) / 200.0, 0.5 * (v + v2), rtol=1e-6)
    assert 29.0e3 < np.linalg.norm(v) < 30.5e3
    assert abs(et.value - (7793.5 * 86400 + 28068.965 + 69.184)) < 1e-3


def test_eclipse_geometry_cases(orc):
    L = orc.lib()
    Rp = 6378136.6
    sun = np.array([1.5e11, 0.0, 0.0]); planet = np.zeros(3)

    def shadow(r):
        r = np.asarray(r, float)
        return L.orc_eclipse_shadow(_p(sun), _p(planet), _p(r), Rp)

    assert shadow([7e6, 0, 0]) == 1.0                    # sub-solar side
    assert shadow([0, 7e6, 0]) == 1.0                    # terminator, well outside the cone
    assert shadow([-7e6, 0, 0]) == 0.0                   # umbra axis
    assert shadow([-7e6, 0, 7.5e6]) == 1.0               # behind the planet but far off-axis
    # penumbra: monotone from 0 to 1 across the band
    zs = np.linspace(6.30e6, 6.46e6, 200)
    f = np.array([shadow([-7e6, 0, z]) for z in zs])
    assert f[0] == 0.0 and f[-1] == 1.0 and np.all(np.diff(f) >= -1e-15)
    assert np.any((f > 0.05) & (f < 0.95))


def test_thr_force_mapping_produces_requested_direction(orc):
    L = orc.lib()
    rng = np.random.RandomState(2)
    loc = np.array([[3.874945160902288e-2, -1.206182747348013, 0.85245], [3.874945160902288e-2, -1.206182747348013, -0.85245],
                    [-3.8749451609022656e-2, -1.206182747348013, 0.85245], [-3.8749451609022656e-2, -1.206182747348013, -0.85245],
                    [-3.874945160902288e-2, 1.206182747348013, 0.85245], [-3.874945160902288e-2, 1.206182747348013, -0.85245],
                    [3.8749451609022656e-2, 1.206182747348013, 0.85245], [3.8749451609022656e-2, 1.206182747348013, -0.85245]])
    s = 0.7071067811865476
    dirs = np.array([[-s, s, 0], [-s, s, 0], [s, s, 0], [s, s, 0], [s, -s, 0], [s, -s, 0], [-s, -s, 0], [-s, -s, 0]])
    D = np.cross(loc, dirs).T
    for _ in range(20):
        Lr = rng.normal(size=3) * 10
        F = np.zeros(8); ang = C.c_double()
        L.orc_thr_force_mapping(_p(Lr), _p(F), C.byref(ang))
        assert np.all(F >= 0) and F.max() <= 0.9 + 1e-12        # on-pulsing, saturation-scaled
        tau = D @ F
        c = np.dot(tau, Lr) / np.linalg.norm(tau) / np.linalg.norm(Lr)
        assert c > 1 - 1e-9                                    # net torque parallel to the request


def _free_sim(orc, row, **cfg):
    return orc.LeoSim(row, orc.default_cfg(**cfg))


def test_orbit_invariants_over_one_interval(orc):
    """Energy / angular momentum of the orbit change only by drag + Sun third body (tiny at 500 km)."""
    row = parity.sample_rows(orc, 1, seed=9)[0]
    # circular-ish 500 km orbit: drag is ~1e-9 relative per interval
    r0, v0 = row[0:3].copy(), row[3:6].copy()
    sim = _free_sim(orc, row)
    sim.run_sim(1)
    st = sim.state()
    r1, v1 = np.array(st.r_BN_N[:]), np.array(st.v_BN_N[:])
    E0 = 0.5 * v0 @ v0 - orc.MU_EARTH / np.linalg.norm(r0)
    E1 = 0.5 * v1 @ v1 - orc.MU_EARTH / np.linalg.norm(r1)
    assert abs(E1 - E0) / abs(E0) < 5e-6
    h0, h1 = np.cross(r0, v0), np.cross(r1, v1)
    assert np.linalg.norm(h1 - h0) / np.linalg.norm(h0) < 5e-6
    assert st.sim_nanos == 180 * 10**9


def test_kepler_propagation_matches_analytic_two_body(orc):
    """RK4 at h = 0.1 s against the closed-form two-body solution (high orbit: drag vanishes;
    the Sun third-body term bounds the difference)."""
    a = 42164e3
    r0, v0 = orc.elem2rv(orc.MU_EARTH, a, 0.0, 0.3, 1.0, 0.0, 0.5)
    row = np.zeros(19); row[0:3] = r0; row[3:6] = v0; row[18] = 12 * 3600.
    sim = _free_sim(orc, row)
    sim.run_sim(1)
    st = sim.state()
    n = np.sqrt(orc.MU_EARTH / a**3)
    r_an, _ = orc.elem2rv(orc.MU_EARTH, a, 0.0, 0.3, 1.0, 0.0, 0.5 + n * 180.0)
    # Sun tidal acceleration ~ 2 mu_s a / d^3 ~ 3e-6 m/s^2 -> <= 0.05 m over 180 s
    assert np.linalg.norm(np.array(st.r_BN_N[:]) - r_an) < 0.1


def test_total_angular_momentum_conserved_without_external_torque(orc):
    """H_N = [NB](I w + sum Js (Omega + g.w) g) is constant while wheels slew the hub (high orbit, zero disturbance)."""
    a = 42164e3
    r0, v0 = orc.elem2rv(orc.MU_EARTH, a, 0.0, 0.3, 1.0, 0.0, 0.5)
    row = np.zeros(19); row[0:3] = r0; row[3:6] = v0; row[6:9] = [0.3, -0.2, 0.1]; row[9:12] = [1e-3, -2e-3, 5e-4]
    row[15:18] = [500., -300., 200.]; row[18] = 12 * 3600.
    sim = _free_sim(orc, row, step_duration=20.0)
    I = np.diag([330 / 12 * (1.38**2 + 1.04**2), 330 / 12 * (1.04**2 + 1.58**2), 330 / 12 * (1.38**2 + 1.58**2)])
    Js = 50.0 / (6000 * 0.10471975511965977)

    def H_N(st):
        q = np.array(st.sigma_BN[:]); w = np.array(st.omega_BN_B[:]); W = np.array(st.Omega[:3])
        BN = np.zeros((3, 3)); orc.lib().orc_MRP2C(_p(q), _p(BN))
        return BN.T @ (I @ w + Js * W)     # balanced wheels: I already holds the wheel inertia about g (Basilisk I_sc)

    sim.run_sim(1)
    H0 = H_N(sim.state())
    for _ in range(5):
        sim.run_sim(1)
    H1 = H_N(sim.state())
    assert np.linalg.norm(H1 - H0) / np.linalg.norm(H0) < 1e-7


def test_controller_converges_to_inertial_reference(orc):
    row = parity.sample_rows(orc, 1, seed=6)[0]
    sim = _free_sim(orc, row)
    for _ in range(4):
        ob, _ = sim.run_sim(1)
    assert ob[0] < 1e-3 and ob[1] < 1e-4          # |sigma_BR| and body rate after 12 min of mode 1
    st = sim.state()
    np.testing.assert_allclose(st.sigma_RN[:], [1.0, 0.0, 0.0])


def test_battery_clamps(orc):
    row = parity.sample_rows(orc, 1, seed=6)[0]
    row[18] = 20.0 * 3600.
    sim = _free_sim(orc, row)
    for _ in range(3):
        ob, _ = sim.run_sim(1)
        assert 0.0 <= ob[3] <= 20.0
    row[18] = 50.0                                 # 50 J drains through the 5 W sink in 10 s unless sunlit
    sim = _free_sim(orc, row)
    obs = [sim.run_sim(0)[0][3] for _ in range(3)]
    assert min(obs) >= 0.0


def test_first_interval_runs_inclusive_stop_time(orc):
    """ConfigureStopTime is inclusive: the first interval executes ticks 0..1800, later ones 1800 each."""
    row = parity.sample_rows(orc, 1, seed=6)[0]
    sim = _free_sim(orc, row)
    sim.run_sim(0); assert sim.state().sim_nanos == 180 * 10**9
    sim.run_sim(0); assert sim.state().sim_nanos == 360 * 10**9
    assert sim.state().task_mask == (2 | 4)
    sim.run_sim(2); assert sim.state().task_mask == (1 | 4 | 8)
    sim.run_sim(7); assert sim.state().task_mask == (1 | 4 | 8)     # unknown action: enables untouched
    sim.run_sim(1); assert sim.state().task_mask == (1 | 4)
