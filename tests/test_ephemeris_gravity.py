"""SURVEY 8(f)-4 on the CPU: Chebyshev ephemeris tables and the planet-fixed degree-2 field.

 * the oracle's pieces against closed forms (table evaluation vs numpy's Chebyshev routines, the Pines recursion vs the
   analytic gradient of the degree-2 potential and vs finite differences of the potential, the orientation DCM vs
   orthogonality / finite differences / the Earth's rotation rate);
 * the host-compiled device core (csrc/leo_core.cuh) against the oracle per decision step with the tables and the field
   switched on -- the same comparison tests/test_gpu_parity.py::test_ephemeris_tables_and_degree2_field makes on the GPU.
"""
import numpy as np
import pytest

from basilisk_env_b200 import ephemeris as eph
from tests import parity

MU, REQ = 0.3986004415e15, 6378136.6


@pytest.fixture()
def clean_oracle(orc):
    yield orc
    orc.set_ephemeris(0, None); orc.set_ephemeris(1, None); orc.set_gravity_coeffs(None)


def potential2(cbar, r):
    """Degree-2 potential from the textbook associated-Legendre form (normalised coefficients)."""
    x, y, z = r
    rr = np.linalg.norm(r)
    sphi = z / rr; cphi = np.hypot(x, y) / rr; lam = np.arctan2(y, x)
    P20, P21, P22 = 0.5 * (3 * sphi ** 2 - 1), 3 * sphi * cphi, 3 * cphi ** 2
    N0, N1, N2 = np.sqrt(5.0), np.sqrt(5.0 / 3.0), np.sqrt(5.0 / 12.0)
    C20, C21, S21, C22, S22 = cbar
    return MU / rr * (REQ / rr) ** 2 * (N0 * C20 * P20 + N1 * P21 * (C21 * np.cos(lam) + S21 * np.sin(lam))
                                        + N2 * P22 * (C22 * np.cos(2 * lam) + S22 * np.sin(2 * lam)))


def test_pines_degree2_matches_gradient_of_the_potential(orc):
    rng = np.random.RandomState(0)
    for trial in range(6):
        cbar = orc.GGM03S_CBAR if trial == 0 else rng.randn(5) * 1e-4
        r = rng.randn(3) * 3e6 + np.array([5.5e6, 1e6, -2e6])
        a = orc.grav_degree2_pfix(cbar, r)
        h = 10.0
        fd = np.array([(potential2(cbar, r + h * e) - potential2(cbar, r - h * e)) / (2 * h) for e in np.eye(3)])
        assert np.linalg.norm(a - fd) <= 1e-7 * np.linalg.norm(a)            # central differences: O(h^2 / r^2) ~ 1e-11, rounding 1e-8
    # zonal-only coefficients reproduce the J2 closed form used by use_j2
    J2 = 1.08262668355e-3
    r = np.array([3e6, -4e6, 4.5e6]); r2 = r @ r
    k = -1.5 * J2 * MU * REQ ** 2 / (r2 * r2 * np.sqrt(r2)); z2 = 5 * r[2] ** 2 / r2
    np.testing.assert_allclose(orc.grav_degree2_pfix([-J2 / np.sqrt(5), 0, 0, 0, 0], r),
                               [k * r[0] * (1 - z2), k * r[1] * (1 - z2), k * r[2] * (3 - z2)], rtol=1e-14)
    assert abs(orc.GGM03S_CBAR[0] * np.sqrt(5) + J2) < 1e-18                   # the default C20 is the J2 of use_j2


def test_table_evaluation_matches_numpy_chebyshev(clean_oracle):
    orc = clean_oracle
    tab = eph.ChebTable.fit(eph.analytic_sun, -3600.0, 86400.0, 3, 9)
    orc.set_ephemeris(0, tab)
    for t in (0.0, 17.3, 86399.0, 90000.0, 200000.5):
        v, r = orc.eph_eval(0, t)
        v_np, r_np = tab(t)
        np.testing.assert_allclose(v, v_np, rtol=1e-14)
        np.testing.assert_allclose(r, r_np, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(v, eph.analytic_sun(t), rtol=1e-11)      # a 9-term fit over one day is exact to rounding
        fd = (eph.analytic_sun(t + 1.0) - eph.analytic_sun(t - 1.0)) / 2.0
        np.testing.assert_allclose(r, fd, rtol=1e-6)                         # rate = derivative of the polynomial
    orc.set_ephemeris(0, None)
    with pytest.raises(AssertionError):
        orc.eph_eval(0, 0.0)


def test_earth_orientation_model(clean_oracle):
    orc = clean_oracle
    P, Pd = orc.earth_orientation(1000.0)
    np.testing.assert_allclose(P @ P.T, np.eye(3), atol=1e-15)
    assert abs(np.linalg.det(P) - 1.0) < 1e-14
    # pole: third row = (cos DEC cos RA, cos DEC sin RA, sin DEC) of the IAU model
    ra, dec, w = eph.iau_earth_angles(1000.0)
    np.testing.assert_allclose(P[2], [np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)], atol=1e-15)
    # rate: finite differences, and the angular velocity -Pd P^T = [omega x] with |omega| = Earth's rotation rate about the pole
    P1, _ = orc.earth_orientation(1001.0); P0, _ = orc.earth_orientation(999.0)
    np.testing.assert_allclose(Pd, (P1 - P0) / 2.0, atol=1e-12)
    Wx = -Pd @ P.T
    np.testing.assert_allclose(Wx + Wx.T, 0.0, atol=1e-18)
    omega = np.array([Wx[2, 1], Wx[0, 2], Wx[1, 0]])
    np.testing.assert_allclose(omega, [0.0, 0.0, 7.2921158e-5], atol=5e-12)   # rad/s in the planet-fixed frame; the pole itself drifts by ~3e-12 rad/s
    # the same DCM from an orientation table fitted to the model
    tab = eph.ChebTable.fit(eph.iau_earth_angles, 0.0, 43200.0, 4, 4)
    orc.set_ephemeris(1, tab)
    Pt, Pdt = orc.earth_orientation(1000.0)
    np.testing.assert_allclose(Pt, P, atol=1e-10)       # W ~ 5e4 rad: the fit carries ~1e-11 rad of rounding
    np.testing.assert_allclose(Pdt, Pd, atol=1e-14)


def run_pair(orc, hostcore, rows, action_seq, sun=None, orient=None, cbar=None, degree2=True, **cfg):
    n = len(rows)
    hc = hostcore.HostCore(n, **cfg)
    ocfg = orc.default_cfg(grav_pfix=int(degree2), **{k: v for k, v in cfg.items() if k in ("step_duration", "rw_set", "use_j2")})
    orc.set_ephemeris(0, sun); orc.set_ephemeris(1, orient); orc.set_gravity_coeffs(cbar)
    if degree2:
        hc.set_gravity_degree2(True, cbar)
    hc.set_ephemeris(0, sun); hc.set_ephemeris(1, orient)
    envs = [orc.LeoEnv(ocfg) for _ in range(n)]
    np.testing.assert_array_equal(hc.reset_ics(rows), np.stack([e.reset(r) for e, r in zip(envs, rows)]))
    worst = {}
    for t, acts in enumerate(action_seq):
        obs, rew, done, reason = hc.step(acts)
        S, I = hc.state()
        for e in range(n):
            o_ob, o_rew, o_done, o_reason = envs[e].step(int(acts[e]))
            where = f"step {t} env {e} action {acts[e]}"
            parity.compare_obs(obs[e], o_ob, where)
            assert done[e] == o_done and reason[e] == o_reason, where
            for k, v in parity.compare_state(envs[e].state(), S[:, e], I[:, e], where).items():
                worst[k] = max(worst.get(k, 0.0), v)
    return worst, hc


def test_hostcore_degree2_field_vs_oracle(clean_oracle, hostcore):
    """Planet-fixed degree-2 field with the analytic orientation (Pines recursion in the oracle, closed-form gradient in
    the core), all three modes, several intervals so that the SPICE-message wrap (quirk Q18) is crossed."""
    orc = clean_oracle
    rows = parity.sample_rows(orc, 4, seed=77)
    acts = np.array([[0, 1, 2, 0], [1, 0, 2, 2], [0, 2, 1, 0]])
    worst, hc = run_pair(orc, hostcore, rows, acts, step_duration=60.0)
    assert max(worst.values()) <= parity.RTOL
    # the tesseral terms matter: with them switched off the trajectory moves by far more than the parity tolerance
    S1, _ = hc.state()
    worst0, hc0 = run_pair(orc, hostcore, rows, acts, cbar=np.array([orc.GGM03S_CBAR[0], 0, 0, 0, 0]), step_duration=60.0)
    S0, _ = hc0.state()
    assert np.linalg.norm(S1[0:3] - S0[0:3], axis=0).max() > 1e-3           # metres after three minutes


def test_hostcore_four_wheels_degree2_and_tables_vs_oracle(clean_oracle, hostcore):
    """Stress configuration + both tables: a perturbed Sun (so that the table is provably what is read) and an
    orientation table with a different prime meridian."""
    orc = clean_oracle
    sun = eph.ChebTable.fit(lambda t: eph.analytic_sun(t + 40 * 86400.0) * 1.01, 0.0, 7200.0, 3, 7)
    orient = eph.ChebTable.fit(lambda t: eph.iau_earth_angles(t) + np.array([0.01, -0.02, 0.5]), 0.0, 10800.0, 2, 5)
    rows = parity.sample_rows(orc, 3, seed=78)
    rows[0, 15:18] = [2800., -2600., 2900.]
    acts = np.array([[2, 0, 1], [2, 1, 0], [0, 2, 2]])
    worst, hc = run_pair(orc, hostcore, rows, acts, sun=sun, orient=orient, step_duration=60.0, rw_set=1)
    assert max(worst.values()) <= parity.RTOL
    # the oracle's Sun message is the table's value at the last SPICE tick (t = 180 s)
    env = orc.LeoEnv(orc.default_cfg(step_duration=60.0)); env.reset(rows[0]); env.step(0)
    np.testing.assert_allclose(env.state().sun_r[:], sun(60.0)[0], rtol=1e-13)


def test_hostcore_sun_table_only_reference_config(clean_oracle, hostcore):
    """Reference configuration (no harmonics) with the Sun from a table fitted to the analytic model: the result equals
    the table-free run to the fit error, and the table run matches the oracle at the parity tolerance."""
    orc = clean_oracle
    sun = eph.ChebTable.fit(eph.analytic_sun, 0.0, 3600.0, 2, 8)
    rows = parity.sample_rows(orc, 3, seed=79)
    acts = np.array([[1, 0, 1], [0, 1, 1]])
    worst, hc = run_pair(orc, hostcore, rows, acts, sun=sun, degree2=False, step_duration=90.0)
    assert max(worst.values()) <= parity.RTOL
    hc0 = hostcore.HostCore(3, step_duration=90.0); hc0.reset_ics(rows)
    for a in acts:
        hc0.step(a)
    S1, _ = hc.state(); S0, _ = hc0.state()
    np.testing.assert_allclose(S1[0:6], S0[0:6], rtol=1e-12)
    np.testing.assert_allclose(S1[parity.F("storedCharge")], S0[parity.F("storedCharge")], rtol=1e-9)
