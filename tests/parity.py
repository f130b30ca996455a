"""TEST INFRASTRUCTURE: comparison of the SoA state of the CUDA path (or its host-compiled core)
with the oracle's state at a decision boundary.

Tolerances (north_star: discrete bit-exact, continuous <= 1e-9 relative per decision interval, FP64):
  r, v, Omega : |delta| / |value|                      <= RTOL
  sigma_BN    : |delta| (MRP, canonical |sigma| <= 1)  <= RTOL * max(1, |sigma|)
  omega_BN_B  : |delta| <= RTOL * |omega| + OMEGA_ATOL, OMEGA_ATOL = 1e-12 rad/s.  The body rate is not one of the
                north-star's 1e-9 quantities (position, velocity, MRP, wheel speeds): under control it settles to
                ~1e-8 rad/s, where a relative measure of a 1e-17 rad/s difference is meaningless, and in the violent
                regime near a low perigee (drag torque ~1 N m against saturated wheels) rounding differences grow
                by ~100x per interval.  1e-12 rad/s over a 180 s interval is 2e-10 rad of attitude, i.e. still 20x
                tighter than what the 1e-9 MRP tolerance implies for the rate.
  storedCharge: relative to max(|E|, 1 Wh = 3600 J; capacity 20 Wh).  The battery integrates panel power x shadow factor every
                tick, so it inherits the 1e-8-level evaluation differences of the shadow factor inside the penumbra (below):
                up to ~2e-6 J per penumbra transit -- 3e-11 of the capacity, but 1e-9 of a battery that is down to its last
                half watt-hour (seen once in the long-horizon runs: 1.07e-9 relative to a 50 J charge).
  shadowFactor / obs[4]: absolute SHADOW_ATOL = 1e-7.  Inside the penumbra the Basilisk formula
                (eclipse.computePercentShadow) is ill-conditioned: the term b^2*acos((c-x)/b) has its
                argument within ~1e-5 of 1 (b = apparent Earth radius ~1.2 rad, a = apparent Sun radius
                4.65e-3 rad), so the literal double-precision evaluation (the oracle) carries ~3e-10 of
                rounding error on average and up to ~6e-8 next to the cone surfaces, while the step core
                evaluates the same function in a regrouped, well-conditioned form (within 1e-12 of 60-digit
                arithmetic: tests/test_hostcore_eclipse.py); and its position gradient (1/penumbra width ~ 1/35 km) turns the 1e-9 relative position
                tolerance (7 mm) into 2e-7.  Outside the penumbra the factor is exactly 0.0 or 1.0 and is
                compared exactly through the done/obs checks.
"""
import numpy as np

from basilisk_env_b200 import _native

RTOL = 1e-9
OMEGA_ATOL = 1e-12
SHADOW_ATOL = 1e-7
CHARGE_FLOOR = 3600.0      # J (1 Wh of the 20 Wh battery)


def F(name):
    return _native.state_field(name)[0]


def vec_err(a, b, floor=0.0):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor, 1e-300))


def compare_state(st, S, I, where="", check_continuous=True, dyn_ns=100000000):
    """st: oracle LeoState; S, I: one env's column of the double / int64 state blocks.  Returns the continuous deviations;
    with check_continuous=False only the discrete quantities are asserted here (the caller applies its own bounds)."""
    errs = {}
    errs["r"] = vec_err(S[F("r_BN_N"):F("r_BN_N") + 3], st.r_BN_N[:])
    errs["v"] = vec_err(S[F("v_BN_N"):F("v_BN_N") + 3], st.v_BN_N[:])
    errs["sigma"] = vec_err(S[F("sigma_BN"):F("sigma_BN") + 3], st.sigma_BN[:], floor=1.0)
    w_o = np.array(st.omega_BN_B[:]); w_k = S[F("omega_BN_B"):F("omega_BN_B") + 3]
    errs["omega"] = float(np.linalg.norm(w_k - w_o) / (np.linalg.norm(w_o) + OMEGA_ATOL / RTOL))
    errs["Omega"] = vec_err(S[F("Omega"):F("Omega") + 4], st.Omega[:4], floor=1.0)
    errs["charge"] = abs(S[F("storedCharge")] - st.storedCharge) / max(abs(st.storedCharge), CHARGE_FLOOR)
    errs["shadow"] = abs(S[F("shadowFactor")] - st.shadowFactor)
    errs["sigma_BR"] = vec_err(S[F("att_guidance"):F("att_guidance") + 3], st.sigma_BR[:], floor=1.0)
    errs["u"] = vec_err(S[F("u_current"):F("u_current") + 4], st.u_current[:4], floor=1e-3)
    for k, v in errs.items():
        tol = SHADOW_ATOL if k == "shadow" else RTOL
        assert not check_continuous or v <= tol, f"{where}: {k} differs by {v:.3e} (> {tol})"
    # discrete quantities: bit-exact
    assert int(I[F("MRPSwitchCount")]) == st.mrp_switch_count, f"{where}: MRP switch count"
    assert int(I[F("task_mask")]) == st.task_mask, f"{where}: task mask"
    assert int(I[F("thrFactorMask")]) == st.thr_factor_mask, f"{where}: thruster firing mask"
    assert int(I[F("initRequest")]) == st.init_request, f"{where}: initRequest"
    assert int(I[F("thrDumpingCounter")]) == st.dump_counter, f"{where}: dumping counter"
    fc = [int(x) for x in I[F("fireCounter"):F("fireCounter") + 8]]
    assert fc == list(st.thr_fire_count[:]), f"{where}: thruster fire counters {fc} vs {list(st.thr_fire_count[:])}"
    assert int(I[F("tick")]) * dyn_ns == st.sim_nanos, f"{where}: sim clock"
    # commanded on-times are compared exactly as well: they gate discrete firing decisions
    on_k = S[F("ThrustOnCmd"):F("ThrustOnCmd") + 8]
    if check_continuous:
        np.testing.assert_allclose(on_k, np.array(st.thrOnCmd[:]), rtol=1e-9, atol=1e-12, err_msg=f"{where}: ThrustOnCmd")
    return errs


def compare_obs(ob_k, ob_o, where=""):
    ob_k = np.asarray(ob_k, float); ob_o = np.asarray(ob_o, float)
    assert abs(ob_k[0] - ob_o[0]) <= RTOL * max(1.0, abs(ob_o[0])), f"{where}: obs[0] {ob_k[0]} vs {ob_o[0]}"
    assert abs(ob_k[1] - ob_o[1]) <= RTOL * abs(ob_o[1]) + OMEGA_ATOL, f"{where}: obs[1] {ob_k[1]} vs {ob_o[1]}"
    assert abs(ob_k[2] - ob_o[2]) <= RTOL * max(abs(ob_o[2]), 1e-3), f"{where}: obs[2] {ob_k[2]} vs {ob_o[2]}"
    assert abs(ob_k[3] - ob_o[3]) <= RTOL * max(abs(ob_o[3]), 1e-3), f"{where}: obs[3] {ob_k[3]} vs {ob_o[3]}"
    assert abs(ob_k[4] - ob_o[4]) <= SHADOW_ATOL, f"{where}: obs[4] {ob_k[4]} vs {ob_o[4]}"
    if ob_o[4] in (0.0, 1.0) and abs(ob_k[4] - ob_o[4]) > 0:
        # umbra / full sun are discrete outcomes; a mismatch is only legitimate on the cone boundary itself
        assert min(ob_k[4], 1.0 - ob_k[4]) < SHADOW_ATOL, f"{where}: eclipse state differs ({ob_k[4]} vs {ob_o[4]})"


def sample_rows(orc, n, seed):
    """n IC rows drawn from a seeded legacy numpy stream in the reference's draw order."""
    rng = np.random.RandomState(seed)
    return np.stack([orc.ic_to_row(orc.sample_ic_dict(rng)) for _ in range(n)])
