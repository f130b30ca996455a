"""GPU suite (`-m gpu`): the CUDA path, called through the C ABI (include/bskenv.h), against the
CPU oracle on the same seeded inputs, against the committed golden fixtures, and -- at the sizes
BASELINE.json names -- through size-independent properties.

"Oracle" = the in-repo FP64 restatement of the Basilisk 1.x algorithms (PARITY UNPINNED: Basilisk
itself cannot be built or imported in this image and the reference ships no golden outputs).
Bar: discrete quantities bit-exact, continuous states <= 1e-9 relative (tests/parity.py)."""
import os

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _vec(bsk, n, **kw):
    from basilisk_env_b200.vec_env import LeoPowerAttVecEnv
    return LeoPowerAttVecEnv(n, device=0, **kw)


def _state_np(env):
    d, i = env.get_state()
    return d.cpu().numpy(), i.cpu().numpy()


def _run_against_oracle(bsk, orc, rows, action_seq, host_path=False, **cfg):
    n = len(rows)
    env = _vec(bsk, n, **cfg)
    ocfg = orc.default_cfg(**{k: v for k, v in cfg.items() if k in ("dynRate", "fswRate", "step_duration", "use_j2", "hill_cel_pun", "rw_set")})
    batch = orc.LeoEnvBatch(rows, ocfg)
    ob0 = env.reset_ics(rows).cpu().numpy()
    np.testing.assert_allclose(ob0, batch.obs0, rtol=1e-14, atol=0)   # reset obs: norms, FMA-contracted on the GPU
    for t, acts in enumerate(action_seq):
        if host_path:
            obs, rew, done, reason = env.step_host(np.asarray(acts, np.int32))
        else:
            o, r, d, info = env.step(acts)
            obs, rew, done, reason = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), info["done_reason"].cpu().numpy()
        S, I = _state_np(env)
        o_ob, o_rew, o_done, o_reason = batch.step(acts)
        for e in range(n):
            where = f"step {t} env {e} action {acts[e]}"
            parity.compare_obs(obs[e], o_ob[e], where)
            assert bool(done[e]) == bool(o_done[e]) and int(reason[e]) == int(o_reason[e]), where
            assert abs(rew[e] - o_rew[e]) <= 1e-12, where
            parity.compare_state(batch.envs[e].state(), S[:, e], I[:, e], where, dyn_ns=int(round(cfg.get("dynRate", 0.1) * 1e9)))
    env.close()


def test_random_actions_64_envs(bsk, orc):
    rows = parity.sample_rows(orc, 64, seed=1)
    acts = np.random.RandomState(2).randint(0, 3, size=(8, 64))
    _run_against_oracle(bsk, orc, rows, acts)


def test_host_buffer_entry_point(bsk, orc):
    rows = parity.sample_rows(orc, 33, seed=3)          # ragged: not a multiple of the warp / block size
    acts = np.random.RandomState(4).randint(0, 3, size=(4, 33))
    _run_against_oracle(bsk, orc, rows, acts, host_path=True)


def test_desat_fires_thrusters(bsk, orc):
    rows = parity.sample_rows(orc, 8, seed=5)
    rows[:, 15:18] = np.random.RandomState(6).uniform(1500, 2900, size=(8, 3)) * np.array([1, -1, 1])
    acts = np.array([[2] * 8, [2] * 8, [2] * 8, [1] * 8, [2] * 8, [0] * 8])
    _run_against_oracle(bsk, orc, rows, acts)


def test_unknown_actions_and_short_interval(bsk, orc):
    rows = parity.sample_rows(orc, 5, seed=7)
    acts = np.array([[0, 1, 2, -1, 9], [-1, -1, -1, 0, 1], [2, 0, 1, 5, 2]])
    _run_against_oracle(bsk, orc, rows, acts, step_duration=60.0)


def test_stress_config_j2_four_wheels(bsk, orc):
    """BASELINE configs[4] (FP64 leg): J2 + drag + eclipse + four wheels in the opNav pyramid with momentum dumping."""
    rows = parity.sample_rows(orc, 40, seed=12)
    rows[:8, 15:18] = np.random.RandomState(13).uniform(1500, 2900, size=(8, 3)) * np.array([1, -1, 1])
    acts = np.random.RandomState(14).randint(0, 3, size=(5, 40))
    acts[:2, :8] = 2
    _run_against_oracle(bsk, orc, rows, acts, use_j2=1, rw_set=1)


def test_single_env(bsk, orc):
    rows = parity.sample_rows(orc, 1, seed=8)
    _run_against_oracle(bsk, orc, rows, [[0], [2], [1]])


def test_4096_envs_one_decision_step(bsk, orc):
    """BASELINE config 2: 4096 envs, parity per decision step (two steps: first = 1801 ticks, second = 1800)."""
    n = 4096
    rows = parity.sample_rows(orc, n, seed=10)
    acts = np.random.RandomState(11).randint(0, 3, size=(2, n))
    env = _vec(bsk, n)
    batch = orc.LeoEnvBatch(rows)
    env.reset_ics(rows)
    for t in range(2):
        o, r, d, info = env.step(acts[t])
        obs = o.cpu().numpy()
        S, I = _state_np(env)
        o_ob, o_rew, o_done, o_reason = batch.step(acts[t])
        np.testing.assert_array_equal(d.cpu().numpy().astype(bool), o_done)
        np.testing.assert_array_equal(info["done_reason"].cpu().numpy().astype(np.int32), o_reason)
        np.testing.assert_allclose(r.cpu().numpy(), o_rew, rtol=0, atol=1e-12)
        for e in range(n):
            parity.compare_obs(obs[e], o_ob[e], f"step {t} env {e}")
        for e in range(0, n, 7):
            parity.compare_state(batch.envs[e].state(), S[:, e], I[:, e], f"step {t} env {e}")
    env.close()


def test_golden_episode_fixture(bsk):
    """BASELINE config 1: single instance, fixed seed, recorded action sequences (constant-0 = the
    reference's own demo, ENV:226-227; and a recorded random sequence), full episode."""
    for name in ("leo_episode_const0.npz", "leo_episode_random.npz"):
        g = np.load(os.path.join(GOLDEN, name))
        env = _vec(bsk, 1)
        ob0 = env.reset_ics(g["ic"][None, :]).cpu().numpy()[0]
        np.testing.assert_allclose(ob0, g["ob0"], rtol=1e-14, atol=0)
        for t, a in enumerate(g["actions"]):
            obs, rew, done, reason = env.step_host(np.array([a], np.int32))
            parity.compare_obs(obs[0], g["obs"][t], f"{name} step {t}")
            assert abs(rew[0] - g["reward"][t]) <= 1e-12
            assert bool(done[0]) == bool(g["done"][t]) and int(reason[0]) == int(g["reason"][t]), f"{name} step {t}"
        S, I = _state_np(env)
        np.testing.assert_allclose(S[0:3, 0], g["final_r"], rtol=1e-9)
        np.testing.assert_allclose(S[3:6, 0], g["final_v"], rtol=1e-9)
        assert int(I[parity.F("MRPSwitchCount"), 0]) == int(g["final_switch"])
        env.close()


def test_gym_api_single_env(bsk, orc):
    """make('leo_power_att_env-v0'): (5,1) float64 obs, 4-tuple, info keys, reset_init replay, legacy RNG stream."""
    np.random.seed(2024)
    env = bsk.make('leo_power_att_env-v0')
    ob = env.reset()
    assert ob.shape == (5, 1) and ob.dtype == np.float64 and ob[4, 0] == 0.0
    d = orc.sample_ic_dict(np.random.RandomState(2024))
    oenv = orc.LeoEnv(); o_ob0 = oenv.reset(d)
    np.testing.assert_allclose(ob[:, 0], o_ob0, rtol=1e-15)
    traj = []
    for a in (0, 1, 2, 0):
        ob, rew, over, info = env.step(a)
        o_ob, o_rew, o_done, _ = oenv.step(a)
        assert ob.shape == (5, 1) and set(info) == {"full_states", "obs"} and info["full_states"] == [] and not over
        parity.compare_obs(ob[:, 0], o_ob, f"gym action {a}")
        assert abs(rew - o_rew) < 1e-12
        traj.append(ob.copy())
    sim_ob, sim_states, sim_over = env.simulator.run_sim(1)
    assert sim_ob.shape == (5, 1) and sim_states == [] and sim_over is False
    ob_r = env.reset_init()
    np.testing.assert_allclose(ob_r[:, 0], o_ob0, rtol=1e-15)
    ob1, *_ = env.step(0)
    np.testing.assert_array_equal(ob1, traj[0])          # deterministic replay from the same ICs
    env.close()


def test_max_length_and_episode_info(bsk):
    from basilisk_env_b200.envs import leoPowerAttEnv
    env = leoPowerAttEnv()
    env.max_length = 2; env.step_duration = 10.0
    np.random.seed(1)
    env.reset()
    outs = [env.step(1) for _ in range(3)]
    assert [o[2] for o in outs] == [False, False, True]
    assert outs[2][3]["episode"]["l"] == 2 and "r" in outs[2][3]["episode"]
    env.close()


def test_determinism_and_shard_invariance(bsk):
    """Same seed -> bit-identical results, and N envs on one handle == the same envs split over two handles."""
    import torch
    n = 1000
    acts = torch.randint(0, 3, (3, n), dtype=torch.int32, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    def run(lo, hi):
        env = _vec(bsk, hi - lo, first_env_index=lo, seed=42, step_duration=30.0)
        ob0 = env.reset().clone()
        outs = []
        for t in range(3):
            o, r, d, _ = env.step(acts[t, lo:hi])
            outs.append((o.clone(), r.clone(), d.clone()))
        st = env.get_state()
        env.close()
        return ob0, outs, st
    full = run(0, n); again = run(0, n)
    assert torch.equal(full[0], again[0]) and torch.equal(full[2][0], again[2][0]) and torch.equal(full[2][1], again[2][1])
    a = run(0, 350); b = run(350, n)
    assert torch.equal(torch.cat([a[0], b[0]]), full[0])
    for t in range(3):
        for k in range(3):
            assert torch.equal(torch.cat([a[1][t][k], b[1][t][k]]), full[1][t][k])
    epi = parity.F("episode")
    keep = [k for k in range(full[2][1].shape[0]) if k != epi]
    assert torch.equal(torch.cat([a[2][0], b[2][0]], dim=1), full[2][0])
    assert torch.equal(torch.cat([a[2][1], b[2][1]], dim=1)[keep], full[2][1][keep])


def test_state_roundtrip_is_a_checkpoint(bsk):
    import torch
    n = 257
    env = _vec(bsk, n, seed=5, step_duration=20.0)
    env.reset()
    acts = torch.randint(0, 3, (4, n), dtype=torch.int32, device="cuda")
    env.step(acts[0]); env.step(acts[1])
    ckpt = tuple(t.clone() for t in env.get_state())
    o1 = [tuple(x.clone() for x in env.step(acts[t])[:3]) for t in (2, 3)]
    env2 = _vec(bsk, n, seed=5, step_duration=20.0)
    env2.set_state(*ckpt)
    o2 = [tuple(x.clone() for x in env2.step(acts[t])[:3]) for t in (2, 3)]
    for x, y in zip(o1, o2):
        assert all(torch.equal(p, q) for p, q in zip(x, y))
    env.close(); env2.close()


def test_auto_reset_and_episode_stats(bsk):
    import torch
    n = 512
    env = _vec(bsk, n, seed=9, auto_reset=True, max_length=3, step_duration=10.0)
    ob0 = env.reset().clone()
    ics0 = env.initial_conditions().clone()
    a = torch.ones(n, dtype=torch.int32, device="cuda")
    dones = []
    for t in range(8):
        o, r, d, info = env.step(a)
        dones.append(d.clone())
        if t == 3:
            assert bool(d.all())                           # the 4th call ends the episode (quirk Q9)
            assert bool((info["done_reason"] & 1).all())
            assert bool((o[:, 4] == 0).all())              # obs of the NEW episode (eclipse entry 0 at reset)
            assert not torch.equal(env.initial_conditions(), ics0)
    assert [int(x.sum()) for x in dones] == [0, 0, 0, n, 0, 0, 0, n]
    st = env.episode_stats()
    assert st["episodes"] == 2 * n and st["max_length_ends"] == 2 * n and st["length_sum"] == 2 * n * 4 and st["env_steps"] == 8 * n
    assert env.launch_count() == 8
    env.close()


def test_mask_reset_and_error_paths(bsk):
    import torch
    from basilisk_env_b200.vec_env import BskEnvError
    env = _vec(bsk, 16, seed=1, step_duration=10.0)
    env.reset()
    env.step(torch.zeros(16, dtype=torch.int32, device="cuda"))
    tick = env.field("tick").clone()
    mask = torch.zeros(16, dtype=torch.uint8); mask[::2] = 1
    env.reset(mask=mask)
    tick2 = env.field("tick")
    assert bool((tick2[0, ::2] == -1).all()) and torch.equal(tick2[0, 1::2], tick[0, 1::2])
    with pytest.raises(ValueError):
        env.step(torch.zeros(15, dtype=torch.int32, device="cuda"))
    with pytest.raises(BskEnvError):
        _vec(bsk, 4, Ki=1.0)                                # integral feedback is not on the reference path
    with pytest.raises(BskEnvError):
        _vec(bsk, 4, dynRate=0.3)                           # fswRate must be a multiple of dynRate
    env.close()


def test_full_size_properties_65536_envs(bsk):
    """BASELINE config sizes (65536 envs/GPU): properties that need no oracle --
    orbit energy drift bounded by drag, |sigma| <= 1 after the switch, battery within [0, capacity],
    wheel momentum exchange: total inertial angular momentum changes only by the external torques."""
    import torch
    n = 65536
    env = _vec(bsk, n, seed=123)
    env.reset()
    d0, _ = env.get_state()
    r0, v0 = d0[0:3].clone(), d0[3:6].clone()
    mu = 0.3986004415e15
    E0 = 0.5 * (v0 * v0).sum(0) - mu / r0.norm(dim=0)
    acts = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda")
    o, r, d, info = env.step(acts)
    d1, i1 = env.get_state()
    r1, v1 = d1[0:3], d1[3:6]
    E1 = 0.5 * (v1 * v1).sum(0) - mu / r1.norm(dim=0)
    rel = ((E1 - E0) / E0.abs()).abs()
    perigee_alt = r0.norm(dim=0).minimum(r1.norm(dim=0)) - 6378136.6
    assert float(rel[perigee_alt > 300e3].max()) < 1e-5           # drag is the only dissipation
    # drag only removes two-body energy; what can add some is the Sun's tide: <= 2 mu_sun r / d^3 * v * 180 s = 0.8 J/kg,
    # i.e. 3e-8 of |E0| = mu / 2a = 2.9e7 J/kg
    assert bool((E1 <= E0 + 1e-7 * E0.abs()).all())
    assert float(d1[6:9].norm(dim=0).max()) <= 1.0 + 1e-12        # MRP switched to the inner set
    assert float(d1[parity.F("storedCharge")].min()) >= 0.0 and float(d1[parity.F("storedCharge")].max()) <= 72000.0
    assert bool((o[:, 4] >= 0).all()) and bool((o[:, 4] <= 1).all())
    assert bool((i1[parity.F("tick")] == 1800).all())
    assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(d1).all())
    # reward only for action 0, within (0, 1/540]
    rw = r[acts == 0]
    assert bool((r[acts != 0] <= 0).all()) and float(rw[rw > 0].max()) <= 1.0 / 540 + 1e-15
    env.close()


def test_step_and_step_host_can_be_mixed(bsk):
    """`step` queues on the caller's stream, `step_host` on the handle's own: the host entry point waits for the
    device first, so an un-synchronised mix gives the same trajectory as a serial one (and counts every episode once)."""
    import torch
    n = 4096
    acts = np.random.RandomState(3).randint(0, 3, size=(6, n)).astype(np.int32)
    outs = []
    for mixed in (False, True):
        env = _vec(bsk, n, seed=11, auto_reset=True, step_duration=20.0, max_length=2)
        env.reset()
        for t in range(6):
            if mixed and t % 2 == 1:
                ob = env.step_host(acts[t])[0].copy()
            else:
                ob = env.step(torch.as_tensor(acts[t], device="cuda"))[0]
                if not mixed:
                    torch.cuda.synchronize()
                    ob = ob.cpu().numpy()
        torch.cuda.synchronize()
        ob = ob if isinstance(ob, np.ndarray) else ob.cpu().numpy()
        outs.append((ob, env.episode_stats()))
        env.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    a, b = outs
    assert abs(a[1].pop("return_sum") - b[1].pop("return_sum")) <= 1e-9       # a floating-point atomic sum: order-dependent rounding
    assert a[1] == b[1] and a[1]["episodes"] == 2 * n


def test_mixed_precision_variant_on_the_stress_config(bsk):
    """BASELINE config 5, accuracy leg: precision = 1 (FP32 stage arithmetic, FP64 accumulation / clocks / FSW / events) against the
    FP64 kernel on the stress configuration, same initial conditions and actions (all three modes).  Not a parity claim: it pins
    the size of the deviation -- sub-metre positions and 1e-6 attitudes after three intervals for 99 % of the envs, identical
    episode-termination flags."""
    import torch
    n = 8192
    envs = [_vec(bsk, n, seed=3, use_j2=1, rw_set=1, precision=p) for p in (0, 1)]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 3, (3, n), dtype=torch.int32, device="cuda", generator=g)
    for t in range(3):
        outs = [[x.clone() for x in e.step(acts[t])[:3]] for e in envs]
    (d0, i0), (d1, i1) = envs[0].get_state(), envs[1].get_state()
    F = parity.F
    dr = (d0[F("r_BN_N"):F("r_BN_N") + 3] - d1[F("r_BN_N"):F("r_BN_N") + 3]).norm(dim=0)
    ds = (d0[F("sigma_BN"):F("sigma_BN") + 3] - d1[F("sigma_BN"):F("sigma_BN") + 3]).abs().max(dim=0).values
    assert float(dr.median()) < 0.5 and float(torch.quantile(dr, 0.99)) < 2.0
    assert float(ds.median()) < 1e-8 and float(torch.quantile(ds, 0.99)) < 1e-5
    assert float((outs[0][2] == outs[1][2]).double().mean()) > 0.999
    assert float((outs[0][0] - outs[1][0]).abs().median()) < 1e-7
    for e in envs:
        e.close()
    with pytest.raises(Exception):
        _vec(bsk, 8, use_j2=1, precision=1)         # built for the reference and the stress configuration only


def test_ephemeris_tables_and_degree2_field(bsk, orc):
    """SURVEY 8(f)-4 through the C ABI: Sun and Earth-orientation Chebyshev tables (bskenv_set_ephemeris) and the
    planet-fixed degree-2 field (bskenv_set_gravity_degree2) against the oracle (Pines recursion, forward Chebyshev
    recurrences), reference wheel set and the four-wheel stress set, all three modes, full 180 s intervals."""
    from basilisk_env_b200 import ephemeris as eph
    from basilisk_env_b200.vec_env import BskEnvError
    sun = eph.ChebTable.fit(lambda t: eph.analytic_sun(t + 40 * 86400.0) * 1.01, 0.0, 16 * 3600.0, 2, 9)
    orient = eph.ChebTable.fit(lambda t: eph.iau_earth_angles(t) + np.array([0.01, -0.02, 0.5]), 0.0, 16 * 3600.0, 2, 5)
    try:
        for case, (kw, tables) in enumerate([(dict(), False), (dict(rw_set=1), True)]):
            n = 40
            rows = parity.sample_rows(orc, n, seed=90 + case)
            rows[:6, 15:18] = np.random.RandomState(91).uniform(1500, 2900, size=(6, 3)) * np.array([1, -1, 1])
            acts = np.random.RandomState(92 + case).randint(0, 3, size=(3, n))
            acts[:2, :6] = 2
            env = _vec(bsk, n, max_length=100, **kw)
            env.set_gravity_degree2(True)
            orc.set_gravity_coeffs(None)
            orc.set_ephemeris(0, sun if tables else None); orc.set_ephemeris(1, orient if tables else None)
            env.set_ephemeris("sun", sun if tables else None); env.set_ephemeris("orientation", orient if tables else None)
            batch = orc.LeoEnvBatch(rows, orc.default_cfg(grav_pfix=1, **kw))
            np.testing.assert_allclose(env.reset_ics(rows).cpu().numpy(), batch.obs0, rtol=1e-14, atol=0)
            for t in range(len(acts)):
                o, r, d, info = env.step(acts[t])
                obs = o.cpu().numpy()
                S, I = _state_np(env)
                o_ob, o_rew, o_done, o_reason = batch.step(acts[t])
                np.testing.assert_array_equal(d.cpu().numpy().astype(bool), o_done)
                for e in range(n):
                    where = f"case {case} step {t} env {e} action {acts[t][e]}"
                    parity.compare_obs(obs[e], o_ob[e], where)
                    parity.compare_state(batch.envs[e].state(), S[:, e], I[:, e], where)
            if case == 0:
                # the tesseral terms are really in the kernel: the zonal-only inertial J2 build ends somewhere else
                env2 = _vec(bsk, n, max_length=100, use_j2=1)
                env2.reset_ics(rows)
                for t in range(len(acts)):
                    env2.step(acts[t])
                S2, _ = _state_np(env2)
                d = np.linalg.norm(S[0:3] - S2[0:3], axis=0)
                assert d.max() > 0.05 and d.max() < 500.0           # metres after nine minutes
                env2.close()
            # a table that does not cover an episode is refused; unloading goes back to the analytic model
            with pytest.raises(BskEnvError):
                env.set_ephemeris("sun", eph.ChebTable.fit(eph.analytic_sun, 0.0, 600.0, 2, 4))
            env.close()
    finally:
        orc.set_ephemeris(0, None); orc.set_ephemeris(1, None); orc.set_gravity_coeffs(None)


def test_sun_table_fitted_to_the_analytic_model_changes_nothing(bsk):
    """A Sun table fitted to the built-in series reproduces the table-free run to the fit error (1e-11 relative)."""
    import torch
    from basilisk_env_b200 import ephemeris as eph
    n = 256
    a = _vec(bsk, n, seed=5); b = _vec(bsk, n, seed=5)
    b.set_ephemeris("sun", eph.ChebTable.fit(eph.analytic_sun, 0.0, 86400.0, 2, 9))
    a.reset(); b.reset()
    acts = torch.randint(0, 3, (3, n), dtype=torch.int32, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    for t in range(3):
        oa = a.step(acts[t])[0]; ob = b.step(acts[t])[0]
    Sa, Ia = _state_np(a); Sb, Ib = _state_np(b)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_allclose(Sb[0:6], Sa[0:6], rtol=1e-12)
    np.testing.assert_allclose(ob.cpu().numpy(), oa.cpu().numpy(), atol=1e-7)
    a.close(); b.close()


def test_one_million_envs_sharded_eight_ways_equals_one_handle(bsk):
    """BASELINE configs[2] at its full size: 2^20 envs with device-sampled random orbits.  One handle on one GPU and the
    eight index shards a box of 8 GPUs would own ([g * 131072, (g + 1) * 131072), here run back to back on the same device)
    give bit-identical observations, rewards, flags and state after a full 180 s decision interval with auto-reset."""
    import torch
    n, shards = 1 << 20, 8
    per = n // shards
    acts = torch.randint(0, 3, (n,), dtype=torch.int32, device="cuda", generator=torch.Generator("cuda").manual_seed(11))
    env = _vec(bsk, n, seed=77, auto_reset=True)
    ob0 = env.reset()
    o, r, d, info = env.step(acts)
    S, I = env.get_state()
    tick = I[parity.F("tick")]
    assert bool(torch.isfinite(o).all()) and bool(((tick == 1800) | ((tick == -1) & d.bool())).all())     # -1: re-initialised in the launch
    ref = (ob0.clone(), o.clone(), r.clone(), d.clone(), S[:12].clone(), S[parity.F("storedCharge")].clone())
    stats = env.episode_stats()
    env.close()
    del env, S, I
    torch.cuda.empty_cache()
    steps = 0.0
    for g in range(shards):
        lo = g * per
        sh = _vec(bsk, per, first_env_index=lo, seed=77, auto_reset=True)
        s_ob0 = sh.reset()
        so, sr, sd, _ = sh.step(acts[lo:lo + per])
        sS, _ = sh.get_state()
        assert torch.equal(s_ob0, ref[0][lo:lo + per]) and torch.equal(so, ref[1][lo:lo + per])
        assert torch.equal(sr, ref[2][lo:lo + per]) and torch.equal(sd, ref[3][lo:lo + per])
        assert torch.equal(sS[:12], ref[4][:, lo:lo + per]) and torch.equal(sS[parity.F("storedCharge")], ref[5][lo:lo + per])
        steps += sh.episode_stats()["env_steps"]
        sh.close()
    assert steps == stats["env_steps"] == float(n)
