"""CPU suite: the fused schedule of csrc/leo_core.cuh (compiled for the host) against the
independent oracle, per decision step -- the same comparison tests/test_gpu_parity.py makes on the
GPU through the C ABI."""
import numpy as np
import pytest

from tests import parity


def run_pair(orc, hostcore, rows, action_seq, **cfg):
    n = len(rows)
    hc = hostcore.HostCore(n, **cfg)
    ocfg = orc.default_cfg(**{k: v for k, v in cfg.items() if k in ("dynRate", "fswRate", "step_duration", "use_j2", "hill_cel_pun", "rw_set")})
    envs = [orc.LeoEnv(ocfg) for _ in range(n)]
    ob_h = hc.reset_ics(rows)
    ob_o = np.stack([e.reset(r) for e, r in zip(envs, rows)])
    np.testing.assert_array_equal(ob_h, ob_o)
    worst = {}
    for t, acts in enumerate(action_seq):
        obs, rew, done, reason = hc.step(acts)
        S, I = hc.state()
        for e in range(n):
            o_ob, o_rew, o_done, o_reason = envs[e].step(int(acts[e]))
            where = f"step {t} env {e} action {acts[e]}"
            parity.compare_obs(obs[e], o_ob, where)
            assert done[e] == o_done and reason[e] == o_reason, where
            assert abs(rew[e] - o_rew) <= 1e-12, where
            errs = parity.compare_state(envs[e].state(), S[:, e], I[:, e], where)
            for k, v in errs.items():
                worst[k] = max(worst.get(k, 0.0), v)
    return worst


def test_all_modes_random_actions(orc, hostcore):
    rows = parity.sample_rows(orc, 6, seed=11)
    rng = np.random.RandomState(5)
    acts = rng.randint(0, 3, size=(10, 6))
    worst = run_pair(orc, hostcore, rows, acts)
    assert max(worst.values()) <= parity.RTOL


@pytest.mark.parametrize("action", [0, 1, 2])
def test_constant_action(orc, hostcore, action):
    rows = parity.sample_rows(orc, 3, seed=100 + action)
    run_pair(orc, hostcore, rows, np.full((6, 3), action))


def test_desat_with_fast_wheels(orc, hostcore):
    """Wheel momentum above hs_min so that mode 2 really fires thrusters (fire counters > 0)."""
    rows = parity.sample_rows(orc, 4, seed=3)
    rows[:, 15:18] = np.array([[2500., -2300., 2700.], [-2900., 2000., 1500.], [200., -150., 100.], [2999., 2999., -2999.]])
    acts = np.array([[2] * 4, [2] * 4, [1] * 4, [2] * 4, [0] * 4, [2] * 4])
    hc_n = 4
    run_pair(orc, hostcore, rows, acts)
    hc = hostcore.HostCore(hc_n); hc.reset_ics(rows)
    hc.step([2] * 4)      # first interval: the RW speed message is still unwritten at t=0 -> Delta H = 0
    _, I = hc.state()
    assert I[parity.F("fireCounter"):parity.F("fireCounter") + 8].sum() == 0
    hc.step([2] * 4)
    _, I = hc.state()
    fire = I[parity.F("fireCounter"):parity.F("fireCounter") + 8]
    assert fire[:, 0].sum() > 0 and fire[:, 2].sum() == 0     # |h| > 4 Nms fires; |h| ~ 2.2 Nms < hs_min does not


def test_unknown_action_keeps_mode(orc, hostcore):
    rows = parity.sample_rows(orc, 2, seed=8)
    acts = np.array([[0, 1], [-1, 7], [2, -1], [5, 5]])
    run_pair(orc, hostcore, rows, acts)


def test_short_step_duration_and_j2(orc, hostcore):
    rows = parity.sample_rows(orc, 2, seed=21)
    run_pair(orc, hostcore, rows, np.array([[0, 2], [1, 0], [2, 1]]), step_duration=60.0)
    run_pair(orc, hostcore, rows, np.array([[0, 2], [1, 0]]), use_j2=1)


def test_stress_config_j2_four_wheel_pyramid(orc, hostcore):
    """BASELINE configs[4]: J2 on, four reaction wheels in the opNav pyramid, momentum dumping; general EOM path."""
    rows = parity.sample_rows(orc, 5, seed=55)
    rows[0, 15:18] = [2800., -2600., 2900.]          # enough momentum for the thrusters to fire in mode 2
    acts = np.array([[2, 0, 1, 2, 0], [2, 1, 0, 2, 1], [0, 2, 2, 1, 0], [1, 0, 2, 0, 2]])
    worst = run_pair(orc, hostcore, rows, acts, use_j2=1, rw_set=1)
    assert max(worst.values()) <= parity.RTOL
    hc = hostcore.HostCore(5, use_j2=1, rw_set=1); hc.reset_ics(rows)
    hc.step([2] * 5); hc.step([2] * 5)
    S, I = hc.state()
    assert I[parity.F("fireCounter"):parity.F("fireCounter") + 8, 0].sum() > 0
    assert np.all(S[parity.F("Omega") + 3] != 0.0)    # the fourth wheel is alive


def test_hill_cel_pun_switch(orc, hostcore):
    rows = parity.sample_rows(orc, 2, seed=33)
    run_pair(orc, hostcore, rows, np.array([[0, 0], [0, 1]]), hill_cel_pun=1)


def test_episode_termination_flags(orc, hostcore):
    """Power failure (battery clamps at exactly 0.0) and wheel over-speed end the episode with -1 each."""
    rows = parity.sample_rows(orc, 2, seed=4)
    rows[0, 18] = 100.0                  # 100 J: the 5 W sink drains it within the first interval unless sunlit
    rows[0, 6:9] = [0.0, 0.0, 0.0]
    rows[1, 15:18] = [2000., 2000., 2000.]   # |Omega| = 3464 rpm > 3000 rpm limit
    hc = hostcore.HostCore(2); hc.reset_ics(rows)
    envs = [orc.LeoEnv() for _ in range(2)]
    for e, r in zip(envs, rows):
        e.reset(r)
    obs, rew, done, reason = hc.step([1, 1])
    outs = [e.step(1) for e in envs]
    assert [o[2] for o in outs] == list(done) and [o[3] for o in outs] == list(reason)
    assert done[1] and (reason[1] & 2) and rew[1] == -1.0
    if obs[0, 3] == 0.0:
        assert done[0] and (reason[0] & 4) and rew[0] == -1.0


def test_max_length_rule(orc, hostcore):
    """Quirk Q9: episode_over is raised by the (max_length+1)-th step call."""
    rows = parity.sample_rows(orc, 1, seed=2)
    hc = hostcore.HostCore(1, max_length=3, step_duration=10.0); hc.reset_ics(rows)
    dones = [bool(hc.step([1])[2][0]) for _ in range(4)]
    assert dones == [False, False, False, True]
    assert hc.step([1])[3][0] & 1


def test_mixed_precision_variant_tracks_fp64(hostcore, orc):
    """precision = 1 (leo_f32.cuh: FP32 stage arithmetic, FP64 accumulation / clocks / flight software / events) against the
    FP64 core on the same inputs: discrete outcomes equal, continuous states within the FP32 increment error per interval
    (position ~0.1 m, attitude ~1e-7).  This is the accuracy side of BASELINE config 5; parity claims are about FP64 only."""
    rows = parity.sample_rows(orc, 12, seed=5)
    for kw in (dict(), dict(use_j2=1, rw_set=1)):
        h64 = hostcore.HostCore(12, **kw); h32 = hostcore.HostCore(12, precision=1, **kw)
        np.testing.assert_array_equal(h64.reset_ics(rows), h32.reset_ics(rows))
        rng = np.random.RandomState(1)
        for t in range(3):
            a = rng.randint(0, 2, 12)                       # modes 0/1 (thruster timing is exercised on the GPU)
            o64, o32 = h64.step(a), h32.step(a)
            S64, I64 = h64.state(); S32, I32 = h32.state()
            assert np.linalg.norm(S64[0:3] - S32[0:3], axis=0).max() < 1.0          # m
            assert np.linalg.norm(S64[3:6] - S32[3:6], axis=0).max() < 5e-3         # m/s
            assert np.abs(S64[6:9] - S32[6:9]).max() < 1e-6 and np.abs(S64[9:12] - S32[9:12]).max() < 1e-7
            np.testing.assert_allclose(o32[0], o64[0], atol=2e-6, rtol=0)
            np.testing.assert_array_equal(o32[2], o64[2]); np.testing.assert_array_equal(o32[3], o64[3])
            np.testing.assert_array_equal(I64[parity.F("MRPSwitchCount")], I32[parity.F("MRPSwitchCount")])
    with pytest.raises(AssertionError):
        hostcore.HostCore(2, precision=2)


def test_full_episodes_without_resynchronisation(orc, hostcore):
    """CPU twin of tests/test_gpu_round2.py::test_long_horizon_batch_parity_all_done_reasons (reduced): 36 envs x complete
    episodes of up to 25 decision intervals of 180 s, random actions, the host-compiled core never re-synchronised with the
    oracle; every env is compared at every step until its episode ends.  All termination reasons occur and match."""
    from tests.test_gpu_round2 import _failing_rows, _perigee_altitude, LOW_PERIGEE_ATT_TOL
    n, L = 36, 24
    rows = _failing_rows(orc, n, seed=41)
    benign = _perigee_altitude(rows) >= 200e3
    hc = hostcore.HostCore(n, max_length=L)
    hc.reset_ics(rows)
    batch = orc.LeoEnvBatch(rows, max_length=L)
    acts = np.random.RandomState(43).randint(0, 3, size=(L + 1, n)).astype(np.int32)
    live = np.ones(n, bool)
    seen = 0
    att = ("sigma", "omega", "Omega", "sigma_BR", "u")
    for t in range(L + 1):
        obs, rew, done, reason = hc.step(acts[t])
        S, I = hc.state()
        o_ob, o_rew, o_done, o_reason = batch.step(acts[t])
        np.testing.assert_array_equal(np.asarray(done, bool)[live], o_done[live])
        np.testing.assert_array_equal(np.asarray(reason)[live], o_reason[live])
        for e in np.flatnonzero(live):
            errs = parity.compare_state(batch.envs[e].state(), S[:, e], I[:, e], f"step {t} env {e}", check_continuous=False)
            for k, v in errs.items():
                tol = parity.SHADOW_ATOL if k == "shadow" else (parity.RTOL if (benign[e] or k not in att) else LOW_PERIGEE_ATT_TOL)
                assert v <= tol, (t, e, k, v)
            if o_done[e]:
                seen |= int(o_reason[e])
        live &= ~o_done
    assert not live.any() and seen & 1 and seen & 2 and seen & 4, seen
