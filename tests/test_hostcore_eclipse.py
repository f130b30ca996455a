"""CPU suite: the step core's eclipse fast path (squared cone radii + guard band, no transcendentals
outside the penumbra) returns what the oracle's eclipse.cpp restatement returns -- exactly 0.0 / 1.0
outside the penumbra, and the reference formula inside it."""
import ctypes as C

import numpy as np

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
    assert np.array_equal(fast, ref)
    assert (ref == 0).sum() > 1000 and (ref == 1).sum() > 1000 and ((ref > 0) & (ref < 1)).sum() > 1000


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
