"""GPU suite (`-m gpu`) of the opNav path: the CUDA kernel, called through the C ABI (bskenv_opnav_*), against the
opNav oracle on the same seeded inputs, against the committed golden fixture, and -- at BASELINE's batch sizes --
through size-independent properties (sharding invariance, checkpoint round trip, filter consistency).

"Oracle" = the in-repo FP64 restatement of the Basilisk 1.x algorithms (PARITY UNPINNED).  Tolerances:
tests/opnav_parity.py."""
import os

import numpy as np
import pytest

from tests import opnav_parity as par

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORC_KEYS = ("dynRate", "fswRate", "step_duration_min", "nav_noise", "camera_reenable", "pixel_noise_std", "circle_unc", "numModes")


def _vec(n, **kw):
    from basilisk_env_b200.opnav_env import OpNavVecEnv
    return OpNavVecEnv(n, device=0, **kw)


def _state_np(env):
    d, i = env.get_state()
    return d.cpu().numpy(), i.cpu().numpy()


def _run_against_oracle(bsk, rows, action_seq, host_path=False, first_env=0, seed=77, sun_table=None, **cfg):
    from oracle import opnav as on
    n = len(rows)
    env = _vec(n, first_env_index=first_env, noise_seed=seed, **cfg)
    if sun_table is not None:
        env.set_ephemeris(sun_table)
    batch = on.OpNavEnvBatch(rows, on.default_cfg(seed=seed, **{k: v for k, v in cfg.items() if k in ORC_KEYS}), first_env_index=first_env)
    ob0 = env.reset_ics(rows).cpu().numpy()
    np.testing.assert_array_equal(ob0, np.zeros((n, 4)))
    for t, acts in enumerate(action_seq):
        if host_path:
            obs, rew, done, reason, dbg = env.step_host(np.asarray(acts, np.int32))
        else:
            o, r, d, info = env.step(acts)
            obs, rew, done, reason = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), info["done_reason"].cpu().numpy()
            dbg = info["full_states"].cpu().numpy()
        S, I = _state_np(env)
        o_ob, o_rew, o_done, o_reason, o_dbg = batch.step(acts)
        for e, st in enumerate(batch.states()):
            where = f"step {t} env {e} action {acts[e]}"
            par.compare_obs(obs[e], o_ob[e], where)
            par.compare_debug(dbg[e], o_dbg[e], where)
            assert bool(done[e]) == bool(o_done[e]) and int(reason[e]) == int(o_reason[e]), where
            assert abs(rew[e] - o_rew[e]) <= 1e-12, where
            par.compare_state(st, S[:, e], I[:, e], where)
    assert env.launch_count() == 2 * len(action_seq)        # two kernels per decision interval (pass 1, pass 2), nothing else
    env.close()


def test_opnav_random_actions_64_envs(bsk):
    from oracle import opnav as on
    rows = par.sample_rows(on, 64, seed=1)
    acts = np.random.RandomState(2).randint(0, 2, size=(4, 64))
    _run_against_oracle(bsk, rows, acts, first_env=1000, camera_reenable=1)


def test_opnav_1024_envs_one_decision_step(bsk):
    """BASELINE configs[3] at batch scale: 1024 envs (orbits over the reference's element ranges), two full 3000-tick decision
    intervals (the first has 3001 ticks), every env against its own oracle run."""
    from oracle import opnav as on
    rows = par.sample_rows(on, 1024, seed=21)
    acts = np.random.RandomState(22).randint(0, 2, size=(2, 1024))
    _run_against_oracle(bsk, rows, acts, first_env=5000, seed=123, camera_reenable=1)


def test_opnav_reference_semantics_and_host_entry_point(bsk):
    """Reference camera behaviour (never re-enabled), ragged batch, host-buffer entry point."""
    from oracle import opnav as on
    rows = par.sample_rows(on, 33, seed=3)
    acts = np.random.RandomState(4).randint(0, 2, size=(3, 33))
    acts[0, :16] = 0
    _run_against_oracle(bsk, rows, acts, host_path=True)


def test_opnav_noise_free_and_unknown_actions(bsk):
    from oracle import opnav as on
    rows = par.sample_rows(on, 6, seed=5)
    acts = np.array([[0, 1, 0, 0, 1, 7], [-1, 1, 1, 0, 0, 0], [1, 0, 5, 1, 0, 1]])
    _run_against_oracle(bsk, rows, acts, nav_noise=0, pixel_noise_std=0.0, camera_reenable=1)


def test_opnav_golden_fixture(bsk):
    g = np.load(os.path.join(GOLDEN, "opnav_batch8.npz"))
    n = len(g["ics"])
    for tag, kw in (("ref", dict()), ("cam", dict(camera_reenable=1))):
        env = _vec(n, first_env_index=int(g["first_env"]), noise_seed=int(g["seed"]), **kw)
        env.reset_ics(g["ics"])
        for t in range(len(g["actions"])):
            o, r, d, info = env.step(g["actions"][t])
            obs, rew, done, dbg = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), info["full_states"].cpu().numpy()
            for e in range(n):
                par.compare_obs(obs[e], g[f"{tag}_obs"][t, e], f"{tag} step {t} env {e}")
                par.compare_debug(dbg[e], g[f"{tag}_debug"][t, e], f"{tag} step {t} env {e}")
            np.testing.assert_allclose(rew, g[f"{tag}_reward"][t], rtol=1e-12, atol=1e-14)
            np.testing.assert_array_equal(done.astype(bool), g[f"{tag}_done"][t])
            np.testing.assert_array_equal(env.field("n_meas")[0].cpu().numpy(), g[f"{tag}_n_meas"][t])
        env.close()


def test_opnav_full_episode_bookkeeping(bsk):
    """41 calls end the episode (40-step limit checked before the action); short interval to keep the oracle quick."""
    from oracle import opnav as on
    rows = par.sample_rows(on, 4, seed=6)
    acts = np.random.RandomState(1).randint(0, 2, size=(42, 4))
    _run_against_oracle(bsk, rows, acts, step_duration_min=1.0, camera_reenable=1)


def test_opnav_sharding_checkpoint_and_auto_reset_at_4096(bsk):
    """BASELINE batch size: (i) 4096 envs in one handle == two handles of 2048 (streams keyed by the global index);
    (ii) get_state -> set_state into a fresh handle continues bit-identically; (iii) the filter stays consistent over
    the batch (normalised position error of order one); (iv) auto-reset re-samples finished envs and the episode
    statistics add up."""
    import torch
    n = 4096
    kw = dict(noise_seed=9, sample_orbit=1, camera_reenable=1, step_duration_min=10.0)
    whole = _vec(n, first_env_index=0, **kw)
    halves = [_vec(n // 2, first_env_index=0, **kw), _vec(n // 2, first_env_index=n // 2, **kw)]
    whole.reset(seed=9)
    for h in halves:
        h.reset(seed=9)
    torch.manual_seed(0)
    for t in range(3):
        a = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda")
        ow = [x.clone() for x in whole.step(a)[:3]]
        oh = [h.step(a[k * n // 2:(k + 1) * n // 2])[:3] for k, h in enumerate(halves)]
        for j in range(3):
            assert torch.equal(ow[j], torch.cat([o[j] for o in oh])), f"step {t} output {j}"
    d, i = whole.get_state()
    clone = _vec(n, first_env_index=0, **kw)
    clone.set_state(d, i)
    a = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda")
    o1 = [x.clone() for x in whole.step(a)[:3]]
    o2 = clone.step(a)[:3]
    for j in range(3):
        assert torch.equal(o1[j], o2[j])
    # filter consistency: |r_nav - r_true| against the filter's own sigma, envs that have measured at least 20 times
    d, i = whole.get_state()
    F = par.F
    meas = i[F("n_meas")] >= 20
    assert int(meas.sum()) > n // 4
    err = (d[F("filter_state"):F("filter_state") + 3] - d[F("r_BN_N"):F("r_BN_N") + 3])[:, meas]
    S = d[F("filter_sBar"):F("filter_sBar") + 21][:, meas]
    sig = torch.stack([S[0].abs(), torch.sqrt(S[1] ** 2 + S[2] ** 2), torch.sqrt(S[3] ** 2 + S[4] ** 2 + S[5] ** 2)])
    z = (err / sig).abs()
    assert float(z.median()) < 3.0 and int(i[F("n_bad")].sum()) == 0
    for e in (whole, clone, *halves):
        e.close()
    # auto-reset
    env = _vec(256, auto_reset=True, max_length=2, step_duration_min=1.0, noise_seed=3)
    env.reset(seed=3)
    zeros = torch.zeros(256, dtype=torch.int32, device="cuda")
    dones = []
    for t in range(6):
        obs, rew, done, info = env.step(zeros + (t % 2))
        dones.append(int(done.sum()))
    assert dones == [0, 0, 256, 0, 0, 256]
    st = env.episode_stats()
    assert st["episodes"] == 512 and st["length_sum"] == 512 * 3 and st["max_length_ends"] == 512 and st["env_steps"] == 6 * 256
    assert torch.equal(obs, torch.zeros_like(obs))                  # first observation of the new episodes
    assert int(env.field("episode")[0].min()) == 3
    env.close()


def test_opnav_gym_surface(bsk):
    """`opNavEnv` / `scenario_OpNav` keep the reference's shapes, keys and step semantics."""
    import basilisk_env_b200 as b
    env = b.opNavEnv()
    assert env.observation_space.shape == (4, 1) and env.action_space.n == 2 and env.max_length == 40
    ob = env.reset()
    assert ob.shape == (4, 1) and not ob.any()
    ob, reward, over, info = env.step(1)
    assert ob.shape == (4, 1) and set(info) == {"full_states", "obs"} and info["full_states"].shape == (12, 1)
    assert 0 < reward <= 1 and not over
    ob, reward, over, info = env.step(0)
    assert reward == 0
    sim = env.simulator
    assert sim.modeCounter == 2 and sim.simTime == 100.0
    env.close()


def test_opnav_sun_ephemeris_table(bsk):
    """SURVEY 8(f)-4 for the opNav env through the C ABI: a perturbed Sun-from-Mars Chebyshev table on both sides."""
    from oracle import opnav as on
    from oracle import oracle as orc
    from basilisk_env_b200.vec_env import BskEnvError
    from tests.test_opnav_hostcore import mars_sun_table
    tab = mars_sun_table(scale=0.97, shift_days=60.0, n_seg=5, seg_len=8 * 3600.0)      # covers the 41 x 50 min episode
    rows = par.sample_rows(on, 48, seed=21)
    acts = np.random.RandomState(22).randint(0, 2, size=(2, 48))
    try:
        orc.set_ephemeris(2, tab)
        _run_against_oracle(bsk, rows, acts, first_env=5, camera_reenable=1, sun_table=tab)
    finally:
        orc.set_ephemeris(2, None)
    env = _vec(8)
    with pytest.raises(BskEnvError):
        env.set_ephemeris(mars_sun_table(n_seg=1, seg_len=3600.0))       # does not cover an episode
    env.set_ephemeris(None)
    env.close()


def test_opnav_per_env_episode_record_matches_oracle(bsk):
    """`info['episode'] = {'r': reward_total, 'l': curr_step}` (opNavEnvironment.py:106-109) per env from the kernel
    (bskenv_opnav_step_info, taken before the in-kernel auto-reset) against the oracle's restatement of the env."""
    import torch
    from oracle import opnav as on
    n, L = 48, 3
    rows = par.sample_rows(on, n, seed=31)
    env = _vec(n, noise_seed=5, auto_reset=True, max_length=L, step_duration_min=2.0, camera_reenable=1)
    batch = on.OpNavEnvBatch(rows, on.default_cfg(seed=5, step_duration_min=2.0, camera_reenable=1), max_length=L)
    env.reset_ics(rows)
    acts = np.random.RandomState(3).randint(0, 2, size=(L + 1, n)).astype(np.int32)
    for t in range(L + 1):
        o, r, d, info = env.step(torch.as_tensor(acts[t], device="cuda"))
        _, o_rew, o_done, _, _ = batch.step(acts[t])
        ep_r, ep_l = info["episode_r"].cpu().numpy(), info["episode_l"].cpu().numpy()
        np.testing.assert_array_equal(d.cpu().numpy().astype(bool), o_done)
        for e in range(n):
            want_r, want_l = batch.envs[e].episode()
            assert int(ep_l[e]) == want_l == t and abs(ep_r[e] - want_r) <= 1e-9 * max(1.0, abs(want_r)), (t, e, ep_r[e], want_r)
    assert o_done.all()                                    # the (L + 1)-th call ends every episode (opNavEnvironment.py:91-92)
    env.close()


def test_opnav_pinned_host_buffers_equal_device_step(bsk):
    """`OpNavVecEnv.host_buffers()`: page-locked, device-mapped buffers that the second kernel of the step writes in place
    (zero-copy); same bytes as the device-buffer entry point on a twin env."""
    import torch
    n = 96
    a, b = _vec(n, noise_seed=9, camera_reenable=1, step_duration_min=2.0), _vec(n, noise_seed=9, camera_reenable=1, step_duration_min=2.0)
    a.reset(seed=4); b.reset(seed=4)
    act_pinned, outs = b.host_buffers()
    acts = np.random.RandomState(6).randint(0, 2, size=(3, n)).astype(np.int32)
    for t in range(3):
        o, r, d, info = a.step(torch.as_tensor(acts[t], device="cuda"))
        act_pinned[:] = acts[t]
        ob, rw, dn, rs, dbg = b.step_host(act_pinned, outs)
        np.testing.assert_array_equal(o.cpu().numpy(), ob); np.testing.assert_array_equal(r.cpu().numpy(), rw)
        np.testing.assert_array_equal(d.cpu().numpy(), dn); np.testing.assert_array_equal(info["done_reason"].cpu().numpy(), rs)
        np.testing.assert_array_equal(info["full_states"].cpu().numpy(), dbg)
    a.close(); b.close()


def test_opnav_queued_three_block_builds_equal_static_two_block_shards(bsk):
    """113664 envs on one handle take the three-blocks-per-SM build of the first pass and the atomic work queues of both
    passes (two resident sets of 56832 envs); three shards of 37888 envs take its two-block build on the static path (each
    warp its own group).  Same global env indices, same actions: obs, reward, done, the per-env episode record and the whole
    state have to agree bit for bit over three decision steps with auto-reset (short intervals: 2 min = 120 ticks, two camera
    frames each), and the measurement counters have to add up."""
    import torch
    n, shards = 113664, 3
    per = n // shards
    kw = dict(noise_seed=21, sample_orbit=1, camera_reenable=1, step_duration_min=2.0, auto_reset=True, max_length=2)
    big = _vec(n, first_env_index=0, **kw)
    parts = [_vec(per, first_env_index=k * per, **kw) for k in range(shards)]
    big.reset(seed=5)
    for p in parts:
        p.reset(seed=5)
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    ended = 0
    for t in range(3):
        a = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda", generator=g)
        o, r, d, info = big.step(a)
        ob = [x.clone() for x in (o, r, d, info["done_reason"], info["episode_r"], info["episode_l"])]
        outs = [p.step(a[k * per:(k + 1) * per]) for k, p in enumerate(parts)]
        cat = [torch.cat([x[0] for x in outs]), torch.cat([x[1] for x in outs]), torch.cat([x[2] for x in outs]),
               torch.cat([x[3]["done_reason"] for x in outs]), torch.cat([x[3]["episode_r"] for x in outs]),
               torch.cat([x[3]["episode_l"] for x in outs])]
        for j, (p, q) in enumerate(zip(ob, cat)):
            assert torch.equal(p, q), f"step {t} output {j}"
        ended += int(d.sum())
    Sb, Ib = big.get_state()
    Sp = torch.cat([p.get_state()[0] for p in parts], dim=1); Ip = torch.cat([p.get_state()[1] for p in parts], dim=1)
    assert torch.equal(Sb, Sp) and torch.equal(Ib, Ip)
    assert ended >= n                                          # max_length = 2: every env finished an episode and was re-sampled
    sb = big.episode_stats()
    sp = [p.episode_stats() for p in parts]
    for k in ("episodes", "env_steps", "measurements"):
        if k in sb:
            assert sb[k] == sum(s[k] for s in sp), (k, sb[k], [s[k] for s in sp])
    big.close()
    for p in parts:
        p.close()


def test_opnav_three_kernel_interval_is_bit_identical(bsk, monkeypatch):
    """Opt-in three-kernel form of the interval (BSKENV_OPNAV_NOISE_SPLIT=1; opnav_core.cuh: opnav_pass0 / opnav_pass1_fed): the
    noise walk of the whole interval first, into a slot-major buffer (120 B per env-tick), then the dynamics / flight-software
    pass fed from it one tick ahead through shared memory (cp.async), then the filter.  Same arithmetic per role as the default
    two-kernel interval: outputs, episode records and state are bit-identical over four intervals with auto-reset, a ragged
    batch and the queued work distribution (60001 envs; 21.6 GB of noise buffer)."""
    import torch
    from basilisk_env_b200.opnav_env import OpNavVecEnv
    n, steps = 60001, 4
    g = torch.Generator("cuda").manual_seed(4)
    acts = torch.randint(0, 2, (steps, n), dtype=torch.int32, device="cuda", generator=g)
    res = {}
    for name in ("default", "split"):
        monkeypatch.setenv("BSKENV_OPNAV_NOISE_SPLIT", "1" if name == "split" else "0")
        env = OpNavVecEnv(n, device=0, auto_reset=True, sample_orbit=1, camera_reenable=1, noise_seed=7, max_length=2)
        env.reset(seed=3)
        outs = []
        l0 = env.launch_count()
        for t in range(steps):
            o, r, d, info = env.step(acts[t])
            outs.append([x.clone() for x in (o, r, d, info["done_reason"], info["full_states"], info["episode_r"], info["episode_l"])])
        S, I = env.get_state()
        res[name] = (outs, S.clone(), I.clone(), (env.launch_count() - l0) // steps)
        env.close()
    a, b = res["default"], res["split"]
    assert a[3] == 2 and b[3] == 3                        # kernels per decision interval
    for t in range(steps):
        for x, y in zip(a[0][t], b[0][t]):
            assert torch.equal(x, y), t
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert int(sum(o[2].sum() for o in a[0])) >= n        # max_length = 2: every env finished an episode and was reset in the kernel
