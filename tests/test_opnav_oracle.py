"""CPU suite: the opNav oracle (oracle/opnav_oracle.c) against closed forms, numpy linear algebra and physical
invariants.  PARITY UNPINNED vs Basilisk (absent); these tests pin the restatement itself."""
import ctypes as C

import numpy as np
import pytest

from oracle import opnav as on
from oracle import oracle as orc

MU = on.MU_MARS
_p = on._p


def test_reference_orbit_elements():
    """The fixed orbit of opNavSimulator.py:173-181: a = 18000 km, e = 0.6 -> periapsis 7200 km, vis-viva holds."""
    rN, vN = on.reference_orbit()
    r, v = np.linalg.norm(rN), np.linalg.norm(vN)
    a = 1.0 / (2.0 / r - v * v / MU)
    assert abs(a - 18000e3) / 18000e3 < 1e-12
    h = np.cross(rN, vN)
    e_vec = np.cross(vN, h) / MU - rN / r
    assert abs(np.linalg.norm(e_vec) - 0.6) < 1e-12
    assert abs(np.degrees(np.arccos(h[2] / np.linalg.norm(h))) - 10.0) < 1e-9


def test_ic_stream_matches_reference_draw_order():
    """rError = uniform(100000,-100000,3), vError = uniform(1000,-1000,3) from numpy's legacy stream (:187-188)."""
    rng = np.random.RandomState(5)
    row = on.sample_ic_row(rng)
    rng2 = np.random.RandomState(5)
    np.testing.assert_array_equal(row[6:9], rng2.uniform(100000, -100000, 3))
    np.testing.assert_array_equal(row[9:12], rng2.uniform(1000, -1000, 3))
    assert np.all(np.abs(row[6:9]) <= 1e5) and np.all(np.abs(row[9:12]) <= 1e3)


def test_noise_stream_statistics_and_keys():
    L = on.lib()
    out = np.zeros(4)
    xs = []
    for tick in range(4000):
        L.orc_opnav_normals(11, 3, 0, tick, 1, 0, _p(out))
        xs.append(out.copy())
    xs = np.array(xs).ravel()
    assert abs(xs.mean()) < 0.03 and abs(xs.std() - 1.0) < 0.03
    assert abs(np.mean(xs ** 4) - 3.0) < 0.3
    a, b = np.zeros(4), np.zeros(4)
    L.orc_opnav_normals(11, 3, 0, 7, 1, 0, _p(a)); L.orc_opnav_normals(11, 3, 0, 7, 1, 0, _p(b))
    np.testing.assert_array_equal(a, b)                       # counter-based: reproducible
    for args in ((12, 3, 0, 7, 1, 0), (11, 4, 0, 7, 1, 0), (11, 3, 1, 7, 1, 0), (11, 3, 0, 8, 1, 0), (11, 3, 0, 7, 2, 0),
                 (11, 3, 0, 7, 1, 1)):
        L.orc_opnav_normals(*args, _p(b))
        assert not np.array_equal(a, b), args                 # every key component selects another stream


def test_circle_projection_inverts_pixel_line():
    """pixelLineConverter applied to the synthetic circle returns the true position (any offset inside the frame)."""
    L = on.lib()
    rng = np.random.RandomState(2)
    R = 3396.19e3
    for _ in range(200):
        d = rng.uniform(8000e3, 30000e3)
        off = rng.uniform(-0.2, 0.2, 2)
        r_C = d * np.array([off[0], off[1], 1.0]) / np.linalg.norm([off[0], off[1], 1.0])
        c = np.zeros(3)
        ok = L.orc_opnav_project_circle(_p(r_C), R, _p(c))
        if not ok:
            continue
        sigma = rng.uniform(-0.4, 0.4, 3)
        CN = np.zeros((3, 3)); L.orc_MRP2C(_p(sigma), _p(CN))
        r_meas, cov = np.zeros(3), np.zeros(9)
        L.orc_opnav_pixel_line(_p(c), 0.25, _p(CN), _p(r_meas), _p(cov))
        r_BN_N_true = -CN.T @ r_C                             # planet at r_C in the camera frame -> s/c position
        assert np.linalg.norm(r_meas - r_BN_N_true) / d < 1e-12
        cov = cov.reshape(3, 3)
        np.testing.assert_allclose(cov, cov.T, rtol=1e-12, atol=1e-6)
        assert np.all(np.linalg.eigvalsh(cov) > 0)
    # invalid: behind the camera, outside the frame, too small
    c = np.zeros(3)
    assert L.orc_opnav_project_circle(_p(np.array([0., 0., -1e7])), R, _p(c)) == 0
    assert L.orc_opnav_project_circle(_p(np.array([9e6, 0., 1e7])), R, _p(c)) == 0
    assert L.orc_opnav_project_circle(_p(np.array([0., 0., 1e9])), R, _p(c)) == 0


def test_ukf_utilities_against_numpy():
    L = on.lib()
    rng = np.random.RandomState(0)
    A = rng.randn(18, 6)
    R = np.zeros(36); L.orc_ukf_qr_just_r(_p(A.copy().ravel()), 18, 6, _p(R))
    R = R.reshape(6, 6)
    np.testing.assert_allclose(R.T @ R, A.T @ A, rtol=1e-12, atol=1e-12)
    assert np.allclose(np.tril(R, -1), 0.0)
    M = rng.randn(6, 6); Pm = M @ M.T + 6 * np.eye(6)
    Lc = np.zeros(36); assert L.orc_ukf_chol_decomp(_p(Pm.ravel()), 6, _p(Lc)) == 0
    np.testing.assert_allclose(Lc.reshape(6, 6), np.linalg.cholesky(Pm), rtol=1e-12, atol=1e-13)
    x = rng.randn(6)
    for beta in (0.7, -0.05):
        out = np.zeros(36)
        assert L.orc_ukf_chol_downdate(_p(Lc), _p(x), beta, 6, _p(out)) == 0
        out = out.reshape(6, 6)
        np.testing.assert_allclose(out @ out.T, Pm + beta * np.outer(x, x), rtol=1e-11, atol=1e-12)
    out = np.zeros(36)
    assert L.orc_ukf_chol_downdate(_p(Lc), _p(100 * x), -1.0, 6, _p(out)) == -1      # not positive definite


def _dense_ut(x, P, Q_sqrt_diag, mu, dt, alpha=0.02, beta=2.0):
    """Textbook scaled unscented transform with full covariances (no square roots)."""
    L = on.lib()
    n = 6
    lam = alpha * alpha * n - n
    S = np.linalg.cholesky(P) * np.sqrt(n + lam)
    pts = [x.copy()] + [x + S[:, i] for i in range(n)] + [x - S[:, i] for i in range(n)]
    Y = []
    for p in pts:
        q = p.copy(); L.orc_ukf_state_prop(_p(q), mu, dt); Y.append(q)
    Y = np.array(Y)
    wm = np.full(13, 0.5 / (n + lam)); wc = wm.copy()
    wm[0] = lam / (n + lam); wc[0] = wm[0] + (1 - alpha * alpha + beta)
    xbar = wm @ Y
    D = Y - xbar
    Pn = (D.T * wc) @ D + np.diag(Q_sqrt_diag ** 2)
    return Y, xbar, Pn, wc


def test_ukf_time_and_measurement_update_against_dense_unscented_transform():
    L = on.lib()
    rN, vN = on.reference_orbit()
    x0 = np.concatenate([rN, vN])
    rng = np.random.RandomState(4)
    M = rng.randn(6, 6) * np.array([3e3, 3e3, 3e3, 3., 3., 3.])[:, None]
    P0 = M @ M.T + np.diag([1e6] * 3 + [1.0] * 3)
    Q = np.diag([1e-6] * 3 + [1e-8] * 3)
    f = on.Ukf()
    L.orc_ukf_init(C.byref(f), _p(x0), _p(P0.ravel()), _p(Q.ravel()), on.MU_MARS, 5.0)
    dt = 1.0
    Y, xbar, Pn, wc = _dense_ut(x0, P0, np.array([1e-3 * dt * dt / 2] * 3 + [1e-4 * dt] * 3), on.MU_MARS, dt)
    L.orc_ukf_time_update(C.byref(f), dt)
    np.testing.assert_allclose(np.array(f.state[:]), Y[0], rtol=1e-15)
    np.testing.assert_allclose(np.array(f.xBar[:]), xbar, rtol=1e-12)
    P_sr = np.array(f.covar[:]).reshape(6, 6)
    sc = np.sqrt(np.outer(np.diag(Pn), np.diag(Pn)))
    assert np.max(np.abs(P_sr - Pn) / sc) < 1e-6       # dense form: cancellation against wc[0] = -2496 limits this check
    # measurement update == Kalman update with H = [I 0] on (state = Y0, xBar, Pn)
    Rm = np.diag([4e7, 9e7, 1e8]) + 1e6
    obs = Y[0][:3] + np.array([2e3, -1e3, 5e2])
    L.orc_ukf_meas_update(C.byref(f), _p(obs), _p((Rm / 5.0).ravel()))
    H = np.hstack([np.eye(3), np.zeros((3, 3))])
    Pxy = P_sr @ H.T - np.vstack([np.diag([(1e-3 * dt * dt / 2) ** 2] * 3), np.zeros((3, 3))])
    Pyy = Pxy[:3] + Rm
    K = Pxy @ np.linalg.inv(Pyy)
    x_new = Y[0] + K @ (obs - xbar[:3])
    P_new = P_sr - K @ Pyy @ K.T
    np.testing.assert_allclose(np.array(f.state[:]), x_new, rtol=1e-10)
    P_f = np.array(f.covar[:]).reshape(6, 6)
    assert np.max(np.abs(P_f - P_new) / sc) < 1e-8
    assert f.n_bad == 0


def test_sun_from_mars_is_a_mars_orbit():
    L = on.lib()
    r, v, et = np.zeros(3), np.zeros(3), np.zeros(1)
    AU = 149597870700.0
    for t in (0.0, 1e5, 3e7):
        L.orc_sun_from_mars(t, _p(r), _p(v), _p(et))
        assert 1.38 * AU < np.linalg.norm(r) < 1.67 * AU
        r2, v2 = np.zeros(3), np.zeros(3)
        L.orc_sun_from_mars(t + 10.0, _p(r2), _p(v2), _p(et))
        np.testing.assert_allclose((r2 - r) / 10.0, v, rtol=1e-4)    # v omits the secular element rates (3e-5)
        assert 21e3 < np.linalg.norm(v) < 27e3


@pytest.fixture(scope="module")
def run0():
    rng = np.random.RandomState(9)
    row = on.sample_ic_row(rng)
    row[6:9] = rng.randn(3) * 1e3; row[9:12] = rng.randn(3) * 141.0       # initial error consistent with covarInit
    sim = on.OpNavSim(row, on.default_cfg(seed=5), env_index=2)
    out = [sim.run_sim(0)]
    st = sim.state()
    return row, sim, out, st


def test_truth_orbit_energy_and_total_angular_momentum(run0):
    """No external torque acts on hub + wheels: the inertial total angular momentum is conserved while the controller
    slews through ~180 deg; the Mars two-body energy of the truth orbit is conserved by the RK4 at 1 s."""
    row, sim, out, st = run0
    r, v = np.array(st.r_BN_N[:]), np.array(st.v_BN_N[:])
    E0 = 0.5 * row[3:6] @ row[3:6] - MU / np.linalg.norm(row[0:3])
    E1 = 0.5 * v @ v - MU / np.linalg.norm(r)
    assert abs(E1 - E0) / abs(E0) < 1e-11
    I = np.diag([900., 800., 600.])
    Js = 50. / (6000. * 2 * np.pi / 60)
    el, az = np.radians(40.), np.radians([45., 135., 225., 315.])
    gs = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.full(4, np.sin(el))], axis=1)
    w, Om = np.array(st.omega_BN_B[:]), np.array(st.Omega[:])
    H_B = I @ w + (Js * Om) @ gs
    BN = np.zeros((3, 3)); on.lib().orc_MRP2C(_p(np.array(st.sigma_BN[:])), _p(BN))
    H_N = BN.T @ H_B
    assert np.linalg.norm(Js * Om) > 0.3                      # the wheels did take up the slew momentum
    assert np.linalg.norm(H_N) < 1e-9 * np.linalg.norm(Js * Om) * 1e3      # started from rest: H_N stays ~0


def test_controller_points_camera_at_mars_and_filter_converges(run0):
    row, sim, out, st = run0
    assert np.linalg.norm(st.sigma_BR[:]) < 1e-4
    assert abs(st.circle[0] - 256) < 3 and abs(st.circle[1] - 256) < 3      # disc centred in the 512 x 512 frame
    assert st.n_images == 51 and 45 <= st.n_meas <= 51 and st.n_bad == 0
    o, d, over = out[0]
    err = np.abs(d[0:3] - d[3:6])
    sig = np.sqrt(np.diag(np.array(st.filt_covar[:]).reshape(6, 6)))[:3]
    assert np.all(err < 4 * sig) and np.all(sig < 5e4)         # consistent estimate (noiseSF = 5 makes it conservative)
    assert np.all(o[1:] > 0) and np.all(o[1:] < 1e-2) and abs(o[0]) <= 1.0


def test_first_step_runs_opnav_mode_and_action1_switches_camera_off_for_good():
    """Quirks: the pending 'OpNavOD' event overrides the first action's task set (opNavSimulator.py:157 + BSK_OpNavFsw.py:
    219-224); action 1 clears cameraIsOn and nothing sets it again (opNavSimulator.py:239 is commented out)."""
    rng = np.random.RandomState(10)
    row = on.sample_ic_row(rng)
    sim = on.OpNavSim(row, on.default_cfg(seed=1))
    sim.run_sim(1)
    st = sim.state()
    assert st.mode == 0 and st.camera_on == 0 and st.n_images == 0 and st.n_meas == 0
    assert np.linalg.norm(st.sigma_BR[:]) < 1e-4              # Mars pointing although sun-safe was requested
    sim.run_sim(1)
    st = sim.state()
    assert st.mode == 1
    assert abs(np.array(st.sun_point[:])[2] - 1.0) < 1e-6     # sun-safe: the CSS estimate sits on the +z body axis
    assert np.linalg.norm(st.sigma_BR[:]) < 1e-3
    assert sum(1 for c in st.css[:] if c > 0) >= 3
    sim.run_sim(0)
    st = sim.state()
    assert st.mode == 0 and st.camera_on == 0 and st.n_images == 0
    sim2 = on.OpNavSim(row, on.default_cfg(seed=1, camera_reenable=1))
    sim2.run_sim(1); sim2.run_sim(0)
    assert sim2.state().camera_on == 1 and sim2.state().n_images == 50


def test_gym_layer_semantics():
    """40-step limit checked before the action (41 calls), reward only for action 1, initial obs zeros, done latches."""
    rng = np.random.RandomState(11)
    env = on.OpNavEnv(on.default_cfg(seed=3, step_duration_min=1.0))
    ob0 = env.reset(on.sample_ic_row(rng))
    np.testing.assert_array_equal(ob0, np.zeros(4))
    n = 0
    while True:
        a = n % 2
        ob, rew, done, reason, dbg = env.step(a)
        n += 1
        if a == 0:
            assert rew == 0
        else:
            nav = (dbg[0:3] - dbg[3:6]) / np.linalg.norm(dbg[3:6])
            assert abs(rew - 1.0 / (1.0 + nav @ nav)) < 1e-15 and 0 < rew <= 1
        if done:
            break
        assert n <= 41
    assert n == 41 and reason == 1
    ob, rew, done, reason, dbg = env.step(0)
    assert done
    env2 = on.OpNavEnv(on.default_cfg(seed=3, step_duration_min=1.0, numModes=5))
    env2.reset(on.sample_ic_row(rng))
    flags = [env2.step(0)[2:4] for _ in range(5)]
    assert [f[0] for f in flags] == [False] * 4 + [True] and flags[-1][1] == 2
