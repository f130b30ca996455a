"""CPU suite: the opNav device core (basilisk_env_b200/csrc/opnav_core.cuh) compiled for the host against the
independent oracle -- fused schedule, noise streams, streaming square-root filter, gym bookkeeping."""
import ctypes as C

import numpy as np
import pytest

from oracle import opnav as on
from tests import opnav_parity as par
from tests.hostcore_binding import HostCoreOpNav


def _run(rows, acts, first_env=0, sun_table=None, **kw):
    cfg = on.default_cfg(seed=77, **kw)
    hc = HostCoreOpNav(len(rows), first_env=first_env, noise_seed=77, **kw)
    if sun_table is not None:
        hc.set_ephemeris(sun_table)
    ob0 = hc.reset_ics(rows)
    np.testing.assert_array_equal(ob0, np.zeros_like(ob0))
    batch = on.OpNavEnvBatch(rows, cfg, first_env_index=first_env)
    for t in range(len(acts)):
        obs, rew, done, reason, dbg = hc.step(acts[t])
        o_ob, o_rew, o_done, o_reason, o_dbg = batch.step(acts[t])
        S, I = hc.state()
        for e, st in enumerate(batch.states()):
            where = f"{kw} step {t} env {e}"
            par.compare_state(st, S[:, e], I[:, e], where)
            par.compare_obs(obs[e], o_ob[e], where)
            par.compare_debug(dbg[e], o_dbg[e], where)
        np.testing.assert_array_equal(done, o_done)
        np.testing.assert_array_equal(reason, o_reason)
        np.testing.assert_allclose(rew, o_rew, rtol=1e-12, atol=1e-14)
    return hc, batch


@pytest.mark.parametrize("kw", [dict(), dict(camera_reenable=1), dict(nav_noise=0, pixel_noise_std=0.0, camera_reenable=1)],
                         ids=["reference", "camera_reenable", "noise_free"])
def test_fused_schedule_matches_oracle(kw):
    rows = par.sample_rows(on, 4, seed=3)
    acts = np.array([[0, 0, 1, 1], [0, 1, 1, 0], [1, 1, 0, 0], [0, 0, 0, 1]])
    _run(rows, acts, first_env=10, **kw)


def test_invalid_action_keeps_the_task_set():
    rows = par.sample_rows(on, 2, seed=4)
    acts = np.array([[0, 1], [7, -1], [1, 0]])
    _run(rows, acts, camera_reenable=1)


def test_noise_streams_are_bit_identical_to_the_oracle():
    """Same Philox counters and Box-Muller on both sides; libm vs libm here, so exactly equal on the CPU."""
    hc = HostCoreOpNav(1, noise_seed=0x1234567890ABCDEF)
    L = on.lib()
    out = np.zeros(4)
    for env, ep, tick, stream, block in ((0, 0, 0, 1, 0), (5, 2, 1234, 1, 3), (2**33 + 1, 7, 119999, 2, 0)):
        L.orc_opnav_normals(0x1234567890ABCDEF, env, ep, tick, stream, block, on._p(out))
        np.testing.assert_array_equal(hc.normals(env, ep, tick, stream, block), out)


def test_sun_ephemeris_matches_oracle():
    hc = HostCoreOpNav(1)
    L = on.lib()
    r, v, et = np.zeros(3), np.zeros(3), np.zeros(1)
    for t in (0.0, 3000.0, 123456.0):
        L.orc_sun_from_mars(t, on._p(r), on._p(v), on._p(et))
        rk, vk = hc.sun(t)
        np.testing.assert_allclose(rk, r, rtol=1e-15); np.testing.assert_allclose(vk, v, rtol=1e-15)


def test_mars_eclipse_fast_path_matches_conical_model():
    """Squared-cone tests + regrouped penumbra form (kernel) vs the literal eclipse.cpp restatement (oracle): exactly 0 / 1
    outside the penumbra, <= 1e-7 inside it (the literal form is ill-conditioned there: tests/parity.py)."""
    from oracle import oracle as orc
    hc = HostCoreOpNav(1)
    L = on.lib()
    sun, v, et = np.zeros(3), np.zeros(3), np.zeros(1)
    L.orc_sun_from_mars(1000.0, on._p(sun), on._p(v), on._p(et))
    s_hat = sun / np.linalg.norm(sun)
    e1 = np.cross(s_hat, [0, 0, 1.0]); e1 /= np.linalg.norm(e1)
    Rm = 3396.19e3
    rng = np.random.RandomState(0)
    zero = np.zeros(3)
    seen = set()
    for _ in range(4000):
        depth = rng.uniform(-2e7, 3e7)                        # along the anti-sun line (negative: sunny side)
        lat = Rm * (1 + rng.uniform(-0.02, 0.02)) if rng.rand() < 0.7 else rng.uniform(0.2 * Rm, 3 * Rm)
        r = -s_hat * depth + e1 * lat
        if np.linalg.norm(r) <= Rm:
            continue
        want = orc.lib().orc_eclipse_shadow(on._p(sun), on._p(zero), on._p(r), Rm)
        got = hc.eclipse(sun, r)
        if want in (0.0, 1.0) and abs(got - want) > 0:
            assert min(got, 1 - got) < 1e-7
        assert abs(got - want) <= 1e-7, (depth, lat, got, want)
        seen.add("umbra" if want == 0.0 else ("sun" if want == 1.0 else "penumbra"))
    assert seen == {"umbra", "sun", "penumbra"}


def test_streaming_filter_matches_basilisk_form():
    """Givens / hyperbolic sweeps (kernel) vs Householder QR + Gill-Murray down-dates (oracle): one time update and
    one measurement update from a dense, correlated covariance."""
    L = on.lib()
    hc = HostCoreOpNav(1)
    rN, vN = on.reference_orbit()
    x0 = np.concatenate([rN, vN])
    rng = np.random.RandomState(8)
    M = rng.randn(6, 6) * np.array([3e3, 3e3, 3e3, 3., 3., 3.])[:, None]
    P0 = M @ M.T + np.diag([1e6] * 3 + [1.0] * 3)
    Q = np.diag([1e-6] * 3 + [1e-8] * 3)
    f = on.Ukf()
    L.orc_ukf_init(C.byref(f), on._p(x0), on._p(P0.ravel()), on._p(Q.ravel()), on.MU_MARS_FSW, 5.0)
    S0 = np.linalg.cholesky(P0)
    S21 = np.array([S0[i, j] for i in range(6) for j in range(i + 1)])
    for dt in (1.0, 0.0, 60.0):
        L.orc_ukf_time_update(C.byref(f), f.timeTag + dt)
        x0k, S21, m = hc.ukf_time_update(x0, S21, dt)
        np.testing.assert_allclose(x0k, np.array(f.state[:]), rtol=1e-15)
        Lk = par.tri_to_full(S21)
        P_o = np.array(f.covar[:]).reshape(6, 6)
        sc = np.sqrt(np.outer(np.diag(P_o), np.diag(P_o)))
        assert np.max(np.abs(Lk @ Lk.T - P_o) / sc) < 1e-9
        np.testing.assert_allclose(x0k + m, np.array(f.xBar[:]), rtol=1e-11)   # the oracle sums with wM[0] = -2499
        x0 = x0k
    Rm = np.array([[4e7, 1e6, -2e6], [1e6, 9e7, 3e6], [-2e6, 3e6, 1e8]])
    obs = x0[:3] + np.array([2e3, -1e3, 5e2])
    L.orc_ukf_meas_update(C.byref(f), on._p(obs), on._p((Rm / 5.0).ravel()))
    R6 = np.array([Rm[0, 0], Rm[1, 0], Rm[2, 0], Rm[1, 1], Rm[2, 1], Rm[2, 2]])
    xk, S21, ok = hc.ukf_meas_update(x0, S21, m, 60.0, obs, R6)
    assert ok and f.n_bad == 0
    np.testing.assert_allclose(xk, np.array(f.state[:]), rtol=1e-11)
    Lk = par.tri_to_full(S21)
    P_o = np.array(f.covar[:]).reshape(6, 6)
    sc = np.sqrt(np.outer(np.diag(P_o), np.diag(P_o)))
    assert np.max(np.abs(Lk @ Lk.T - P_o) / sc) < 1e-9
    assert np.all(np.diag(Lk) > 0)


def test_device_sampler_ranges_and_sharding_invariance():
    """Per-env streams are keyed by the GLOBAL env index: 6 envs in one batch == two shards of 3, bit for bit."""
    whole = HostCoreOpNav(6, first_env=100, noise_seed=5, sample_orbit=1, camera_reenable=1)
    ics, ob0 = whole.reset_seeded(42)
    assert np.all(np.abs(ics[:, 6:9]) <= 1e5) and np.all(np.abs(ics[:, 9:12]) <= 1e3)
    r = np.linalg.norm(ics[:, 0:3], axis=1); v = np.linalg.norm(ics[:, 3:6], axis=1)
    a = 1.0 / (2.0 / r - v * v / on.MU_MARS)
    assert np.all(a > 17000e3 * (1 - 1e-12)) and np.all(a < 22000e3 * (1 + 1e-12))
    fixed = HostCoreOpNav(2, noise_seed=5)
    ics_f, _ = fixed.reset_seeded(42)
    rN, vN = on.reference_orbit()
    np.testing.assert_allclose(ics_f[:, 0:3], np.tile(rN, (2, 1)), rtol=1e-15)
    acts = np.array([0, 1, 0, 0, 1, 1])
    shards = [HostCoreOpNav(3, first_env=100, noise_seed=5, sample_orbit=1, camera_reenable=1),
              HostCoreOpNav(3, first_env=103, noise_seed=5, sample_orbit=1, camera_reenable=1)]
    ics_s = np.concatenate([s.reset_seeded(42)[0] for s in shards])
    np.testing.assert_array_equal(ics, ics_s)
    for t in range(2):
        ow = whole.step(acts)
        os_ = [s.step(acts[3 * k:3 * k + 3]) for k, s in enumerate(shards)]
        for j in range(5):
            np.testing.assert_array_equal(ow[j], np.concatenate([o[j] for o in os_]))


def test_episode_bookkeeping_matches_oracle_to_the_end():
    """A full 41-call episode at a short decision interval: done / reason / reward sequence identical."""
    rows = par.sample_rows(on, 2, seed=6)
    acts = np.random.RandomState(1).randint(0, 2, size=(42, 2))
    hc, batch = _run(rows, acts, step_duration_min=1.0, camera_reenable=1)
    S, I = hc.state()
    assert list(I[par.F("curr_step")]) == [42, 42] and list(I[par.F("episode_over")]) == [1, 1]


def mars_sun_table(scale=1.0, shift_days=0.0, n_seg=3, n_coef=8, seg_len=8 * 3600.0):
    """Chebyshev table fitted to the oracle's analytic Sun-from-Mars series (optionally perturbed, so that a test can tell
    the table from the built-in model)."""
    from basilisk_env_b200 import ephemeris as eph
    L = on.lib()

    def fn(t):
        r, v, et = np.zeros(3), np.zeros(3), np.zeros(1)
        L.orc_sun_from_mars(t + shift_days * 86400.0, on._p(r), on._p(v), on._p(et))
        return r * scale
    return eph.ChebTable.fit(fn, 0.0, seg_len, n_seg, n_coef)


def test_sun_table_replaces_the_analytic_series():
    """SURVEY 8(f)-4 for the opNav env: with a (perturbed) Sun table loaded on both sides the fused schedule still matches
    the oracle, the table is what the core evaluates, and the sun-safe observation really changes."""
    from oracle import oracle as orc
    tab = mars_sun_table(scale=0.97, shift_days=60.0)
    rows = par.sample_rows(on, 3, seed=9)
    acts = np.array([[0, 1, 1], [1, 1, 0]])
    kw = dict(step_duration_min=10.0, camera_reenable=1)
    try:
        orc.set_ephemeris(2, tab)
        hc, batch = _run(rows, acts, sun_table=tab, **kw)
        rk, vk = hc.sun(700.0)
        np.testing.assert_allclose(rk, tab(700.0)[0], rtol=1e-14)
        np.testing.assert_allclose(vk, tab(700.0)[1], rtol=1e-10)
    finally:
        orc.set_ephemeris(2, None)
    hc0, _ = _run(rows, acts, **kw)
    S1, _ = hc.state(); S0, _ = hc0.state()
    assert np.abs(S1[par.F("sigma_BN"):par.F("sigma_BN") + 3] - S0[par.F("sigma_BN"):par.F("sigma_BN") + 3]).max() > 1e-3


def test_three_pass_interval_is_bit_identical_to_the_fused_first_pass(monkeypatch):
    """The opt-in three-kernel form of the interval (opnav_core.cuh: opnav_pass0 -> noise buffer -> opnav_pass1_fed -> opnav_pass2;
    GPU: tests/test_opnav_gpu.py::test_opnav_three_kernel_interval_is_bit_identical) run back to back on the host: same bits as
    the default two-pass interval, observations, rewards, flags and the whole state, over three intervals incl. the first tick."""
    rows = par.sample_rows(on, 6, seed=5)
    acts = np.array([[0, 1, 0, 1, 0, 0], [1, 1, 0, 0, 0, 1], [0, 0, 1, 1, 1, 0]], np.int32)
    res = {}
    for name in ("fused", "split"):
        monkeypatch.setenv("HCO_NOISE_SPLIT", "1" if name == "split" else "0")
        hc = HostCoreOpNav(len(rows), noise_seed=9, camera_reenable=1)
        hc.reset_ics(rows)
        outs = [hc.step(a) for a in acts]
        res[name] = (outs, hc.state())
    for a, b in zip(res["fused"][0], res["split"][0]):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(res["fused"][1][0], res["split"][1][0])
    np.testing.assert_array_equal(res["fused"][1][1], res["split"][1][1])
    assert np.abs(res["split"][1][0]).sum() > 0
