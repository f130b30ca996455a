"""Chebyshev ephemeris tables for the batched step.

The reference reads the Sun's position from SPICE (de430.bsp through Basilisk's spice_interface,
/root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py:219-225) once per decision interval.
SPICE itself is not part of this repository; what the kernel consumes is a table in the layout of an SPK type 2 /
binary PCK type 2 record -- equal-length segments, three components, Chebyshev coefficients per component, the rate
being the derivative of the same polynomial -- uploaded with `bskenv_set_ephemeris` (include/bskenv.h).

A table can come from
  * `ChebTable.fit(fn, ...)`: a fit of any callable t -> (3,) on Chebyshev nodes (e.g. `spiceypy.spkpos` on a machine
    that has the kernels; the analytic models below in this image), or
  * `ChebTable(t0, seg_len, coef)`: coefficients taken from elsewhere, e.g. copied out of an SPK type 2 segment
    (their MID / RADIUS / coefficient records map one to one after shifting the epoch to sim time 0 and km -> m).

`analytic_sun` / `iau_earth_angles` restate the closed-form models the kernel falls back to without a table (DESIGN D1);
they exist so that a fitted table can be checked against the built-in model."""
import numpy as np

EPOCH_DAYS_TT_FROM_J2000 = 7793.5 + (28068.965 + 69.184) / 86400.0   # '2021 MAY 04 07:47:48.965 (UTC)' (SIM:219)
D2R = np.pi / 180.0
AU_M = 149597870.693 * 1000.0


class ChebTable:
    """coef[n_seg, 3, n_coef]; segment i covers [t0 + i*seg_len, t0 + (i+1)*seg_len] seconds of sim time."""

    def __init__(self, t0, seg_len, coef):
        self.t0 = float(t0)
        self.seg_len = float(seg_len)
        self.coef = np.ascontiguousarray(coef, dtype=np.float64)
        if self.coef.ndim != 3 or self.coef.shape[1] != 3 or not self.seg_len > 0:
            raise ValueError("coef must be [n_seg, 3, n_coef] and seg_len positive")

    @classmethod
    def fit(cls, fn, t0, seg_len, n_seg, n_coef):
        """Interpolate fn(t) -> (3,) at the n_coef Chebyshev nodes of every segment (discrete orthogonality)."""
        k = np.arange(n_coef)
        nodes = np.cos(np.pi * (k + 0.5) / n_coef)                      # roots of T_n
        Tk = np.cos(np.outer(k, np.arccos(nodes)))                      # [k, node]
        coef = np.zeros((n_seg, 3, n_coef))
        for i in range(n_seg):
            mid = t0 + (i + 0.5) * seg_len
            y = np.stack([np.asarray(fn(mid + 0.5 * seg_len * s), dtype=np.float64).reshape(3) for s in nodes])   # [node, 3]
            c = 2.0 / n_coef * Tk @ y                                   # [k, 3]
            c[0] *= 0.5
            coef[i] = c.T
        return cls(t0, seg_len, coef)

    @classmethod
    def from_nodes(cls, t0, seg_len, pos, vel):
        """Table through recorded states: pos[k], vel[k] (each [n_seg + 1, 3]) at t0 + k*seg_len -- e.g. the Sun messages a
        Basilisk run logged once per decision interval.  One cubic per segment that reproduces position AND velocity at
        both ends (Hermite data written in the Chebyshev basis), so a look-up at a node returns what was recorded."""
        pos = np.asarray(pos, dtype=np.float64); vel = np.asarray(vel, dtype=np.float64)
        if pos.ndim != 2 or pos.shape[1] != 3 or pos.shape != vel.shape or pos.shape[0] < 2:
            raise ValueError("pos and vel must be [n_seg + 1, 3]")
        half = 0.5 * float(seg_len)
        p0, p1, m0, m1 = pos[:-1], pos[1:], vel[:-1] * half, vel[1:] * half
        a2 = (m1 - m0) / 8.0
        a3 = ((m1 + m0) / 2.0 - (p1 - p0) / 2.0) / 8.0
        a1 = (p1 - p0) / 2.0 - a3
        a0 = (p1 + p0) / 2.0 - a2
        return cls(t0, seg_len, np.stack([a0, a1, a2, a3], axis=2))

    def extended(self, n_seg_total):
        """The same table padded with straight-line segments (last position, last velocity) up to n_seg_total segments:
        `bskenv_set_ephemeris` wants a table that covers a whole episode even when the recorded trace is shorter."""
        n = self.coef.shape[0]
        if n_seg_total <= n:
            return self
        val, rate = self(self.t0 + n * self.seg_len)
        pad = np.zeros((n_seg_total - n, 3, self.coef.shape[2]))
        for k in range(n_seg_total - n):
            pad[k, :, 0] = val + rate * self.seg_len * (k + 0.5)
            if self.coef.shape[2] > 1:
                pad[k, :, 1] = rate * 0.5 * self.seg_len
        return ChebTable(self.t0, self.seg_len, np.concatenate([self.coef, pad], axis=0))

    def __call__(self, t):
        """value (3,), rate (3,) at sim time t [s] -- numpy's own Chebyshev evaluation (not the kernel's recurrence)."""
        from numpy.polynomial import chebyshev as ch
        i = int(np.clip(np.floor((t - self.t0) / self.seg_len), 0, self.coef.shape[0] - 1))
        half = 0.5 * self.seg_len
        s = (t - (self.t0 + (i + 0.5) * self.seg_len)) / half
        val = np.array([ch.chebval(s, self.coef[i, c]) for c in range(3)])
        rate = np.array([ch.chebval(s, ch.chebder(self.coef[i, c])) if self.coef.shape[2] > 1 else 0.0 for c in range(3)]) / half
        return val, rate


def analytic_sun(t):
    """Sun relative to Earth [m], mean equator / equinox treated as J2000: the Astronomical-Almanac low-precision series
    the kernel uses without a table (csrc/leo_core.cuh: sun_latch)."""
    n = EPOCH_DAYS_TT_FROM_J2000 + t / 86400.0
    L = (280.460 + 0.9856474 * n) * D2R
    g = (357.528 + 0.9856003 * n) * D2R
    lam = L + (1.915 * np.sin(g) + 0.020 * np.sin(2 * g)) * D2R
    eps = (23.439 - 4e-7 * n) * D2R
    R = (1.00014 - 0.01671 * np.cos(g) - 0.00014 * np.cos(2 * g)) * AU_M
    return R * np.array([np.cos(lam), np.cos(eps) * np.sin(lam), np.sin(eps) * np.sin(lam)])


def iau_earth_angles(t):
    """RA, DEC of the pole and prime-meridian angle W [rad] of the IAU Earth rotation model (SPICE pck00010.tpc)."""
    d = EPOCH_DAYS_TT_FROM_J2000 + t / 86400.0
    T = d / 36525.0
    return np.array([(0.0 - 0.641 * T) * D2R, (90.0 - 0.557 * T) * D2R, (190.147 + 360.9856235 * d) * D2R])
