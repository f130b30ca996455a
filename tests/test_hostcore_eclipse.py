"""CPU suite: the step core's eclipse fast path (squared cone radii + guard band, no transcendentals
outside the penumbra) returns what the oracle's eclipse.cpp restatement returns -- exactly 0.0 / 1.0
outside the penumbra, and the reference's disk-overlap formula (regrouped for a small Sun disk) inside it."""
import ctypes as C

import numpy as np
import pytest

P = C.POINTER(C.c_double)


def test_fast_path_equals_reference_formula(orc, hostcore):
    hc = hostcore.HostCore(1)
    rng = np.random.RandomState(0)
    Rp = 6378136.6
    n = 200000
    # positions all over LEO shells ...
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    r = u * rng.uniform(Rp + 150e3, Rp + 900e3, size=(n, 1))
    _, sun = hc.eclipse(0, r[:1])
    s_hat = sun / np.linalg.norm(sun)
    # ... plus a dense set straddling the umbra and penumbra cone surfaces behind the planet
    m = 100000
    depth = rng.uniform(1e5, 7.2e6, m)
    ang = rng.uniform(0, 2 * np.pi, m)
    e1 = np.cross(s_hat, [0, 0, 1.0]); e1 /= np.linalg.norm(e1); e2 = np.cross(s_hat, e1)
    f = 4.65e-3
    rad_u = Rp - depth * f; rad_p = Rp + depth * f
    rad = np.where(rng.rand(m) < 0.5, rad_u, rad_p) + rng.normal(scale=np.where(rng.rand(m) < 0.5, 5.0, 2e4), size=m)
    r2 = -depth[:, None] * s_hat + rad[:, None] * (np.cos(ang)[:, None] * e1 + np.sin(ang)[:, None] * e2)
    r2 = r2[np.linalg.norm(r2, axis=1) > Rp + 100e3]
    pts = np.vstack([r, r2])
    fast, sun = hc.eclipse(0, pts)
    L = orc.lib()
    planet = np.zeros(3)
    ref = np.array([L.orc_eclipse_shadow(sun.ctypes.data_as(P), planet.ctypes.data_as(P), p.ctypes.data_as(P), Rp)
                    for p in np.ascontiguousarray(pts)])
    # umbra / full sun: exact.  Penumbra: the regrouped disk-overlap formula is the same function, but the literal
    # double-precision evaluation is ill-conditioned (b^2 acos((c-x)/b) with an argument within 1e-5 of 1): it
    # carries up to ~6e-8 of rounding error itself (test_penumbra_fraction_against_exact_arithmetic below), so the
    # two agree to parity.SHADOW_ATOL only.
    clear = (ref == 0) | (ref == 1)
    assert np.array_equal(fast[clear], ref[clear])
    assert np.abs(fast - ref).max() <= 1e-7
    assert np.abs(fast - ref)[~clear].mean() <= 2e-9
    assert (ref == 0).sum() > 1000 and (ref == 1).sum() > 1000 and ((ref > 0) & (ref < 1)).sum() > 1000


def test_penumbra_fraction_against_exact_arithmetic(orc, hostcore):
    """eclipse.cpp computePercentShadow evaluated in 60-digit arithmetic is the truth: the step core's regrouped
    form is within 1e-12 of it, the literal double-precision form (the oracle) within 1e-7."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    hc = hostcore.HostCore(1)
    rng = np.random.RandomState(1)
    Rp, RS = 6378136.6, 695000.0 * 1000
    _, sun = hc.eclipse(0, np.array([[7e6, 0, 0.]]))
    s_hat = sun / np.linalg.norm(sun)
    m = 1500
    depth = rng.uniform(1e5, 7.2e6, m); ang = rng.uniform(0, 2 * np.pi, m)
    e1 = np.cross(s_hat, [0, 0, 1.0]); e1 /= np.linalg.norm(e1); e2 = np.cross(s_hat, e1)
    rad = Rp + depth * 4.65e-3 * rng.uniform(-1.05, 1.05, m)
    pts = -depth[:, None] * s_hat + rad[:, None] * (np.cos(ang)[:, None] * e1 + np.sin(ang)[:, None] * e2)
    pts = np.ascontiguousarray(pts[np.linalg.norm(pts, axis=1) > Rp + 100e3])
    fast, sun = hc.eclipse(0, pts)
    L = orc.lib()
    planet = np.zeros(3)
    ref = np.array([L.orc_eclipse_shadow(sun.ctypes.data_as(P), planet.ctypes.data_as(P), p.ctypes.data_as(P), Rp) for p in pts])

    def exact(r):
        S = [mp.mpf(float(v)) for v in sun]; R = [mp.mpf(float(v)) for v in r]
        hb = [S[i] - R[i] for i in range(3)]
        nH = mp.sqrt(sum(v * v for v in hb)); nB = mp.sqrt(sum(v * v for v in R))
        a = mp.asin(RS / nH); b = mp.asin(mp.mpf(Rp) / nB); c = mp.acos(-sum(R[i] * hb[i] for i in range(3)) / (nB * nH))
        if c < b - a:
            return mp.mpf(0)
        if c < a + b:
            x = (c * c + a * a - b * b) / (2 * c); y = mp.sqrt(a * a - x * x)
            return 1 - (a * a * mp.acos(x / a) + b * b * mp.acos((c - x) / b) - c * y) / (mp.pi * a * a)
        return mp.mpf(1)

    pen = np.nonzero((ref > 0) & (ref < 1))[0]
    assert len(pen) > 800
    err_fast = max(abs(float(fast[i] - exact(pts[i]))) for i in pen)
    err_ref = max(abs(float(ref[i] - exact(pts[i]))) for i in pen)
    assert err_fast <= 1e-12, err_fast
    assert err_ref <= 1e-7, err_ref


def test_reference_penumbra_formula_is_ill_conditioned(orc, hostcore):
    """Why shadowFactor / obs[4] carries an absolute tolerance of 1e-7 (tests/parity.py): inside the penumbra
    the reference's computePercentShadow, evaluated literally in double precision (the oracle), changes by up to ~1e-8 when
    the spacecraft position moves by ONE unit in the last place (1e-9 m) -- an amplification of ~1e7 -- while the regrouped
    form of the step core moves by < 1e-13.  No evaluation of the literal formula can agree with another one to 1e-9."""
    hc = hostcore.HostCore(1)
    rng = np.random.RandomState(3)
    Rp = 6378136.6
    _, sun = hc.eclipse(0, np.array([[7e6, 0, 0.]]))
    s_hat = sun / np.linalg.norm(sun)
    m = 4000
    depth = rng.uniform(1e5, 7.2e6, m); ang = rng.uniform(0, 2 * np.pi, m)
    e1 = np.cross(s_hat, [0, 0, 1.0]); e1 /= np.linalg.norm(e1); e2 = np.cross(s_hat, e1)
    rad = Rp + depth * 4.65e-3 * rng.uniform(-0.9, 0.9, m)
    pts = -depth[:, None] * s_hat + rad[:, None] * (np.cos(ang)[:, None] * e1 + np.sin(ang)[:, None] * e2)
    pts = np.ascontiguousarray(pts[np.linalg.norm(pts, axis=1) > Rp + 100e3])
    pts2 = np.nextafter(pts, np.inf)                          # every coordinate one ulp up (~1e-9 m)
    L = orc.lib()
    planet = np.zeros(3)
    ref = lambda q: np.array([L.orc_eclipse_shadow(sun.ctypes.data_as(P), planet.ctypes.data_as(P), p.ctypes.data_as(P), Rp) for p in q])   # noqa: E731
    r1, r2 = ref(pts), ref(pts2)
    f1, _ = hc.eclipse(0, pts); f2, _ = hc.eclipse(0, pts2)
    pen = (r1 > 0) & (r1 < 1)
    assert pen.sum() > 2000
    d_ref, d_fast = np.abs(r1 - r2)[pen].max(), np.abs(f1 - f2)[pen].max()
    print(f"one-ulp position change: literal formula moves by up to {d_ref:.2e}, regrouped form by {d_fast:.2e}")
    assert 1e-10 < d_ref < 1e-6 and d_fast < 1e-12


def test_general_eom_path_equals_fast_path(hostcore, orc):
    """DIAG fast path (structural zeros dropped, wheel speeds from the per-axis momentum invariant) == general
    3x3 path (per-wheel invariants) up to rounding: discrete outputs exact, continuous ones to 1e-10."""
    from tests import parity
    rows = parity.sample_rows(orc, 4, seed=77)
    a = hostcore.HostCore(4); b = hostcore.HostCore(4); b.force_general()
    a.reset_ics(rows); b.reset_ics(rows)
    for acts in ([0, 1, 2, 0], [2, 2, 1, 0], [1, 0, 2, 2]):
        oa = a.step(acts); ob = b.step(acts)
        np.testing.assert_allclose(oa[0], ob[0], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(oa[1], ob[1], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(oa[2], ob[2]); np.testing.assert_array_equal(oa[3], ob[3])
    Sa, Ia = a.state(); Sb, Ib = b.state()
    np.testing.assert_allclose(Sa, Sb, rtol=1e-10, atol=1e-13)
    np.testing.assert_array_equal(Ia, Ib)
