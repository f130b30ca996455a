"""CPU suite: host-side logic -- spaces, registry, IC draw order, sharding, action decoding."""
import numpy as np
import pytest


def test_registry_and_spaces():
    import basilisk_env_b200 as b
    assert 'leo_power_att_env-v0' in b.registered_ids()
    with pytest.raises(KeyError):
        b.make('opnav_env-v0')                       # commented out in the reference as well
    box = b.spaces.Box(-1e16, 1e16, shape=(5, 1))
    assert box.shape == (5, 1) and box.contains(np.zeros((5, 1))) and not box.contains(np.zeros(5))
    d = b.spaces.Discrete(3)
    assert d.n == 3 and d.contains(2) and not d.contains(3) and not d.contains(1.5)
    d.seed(0)
    assert all(0 <= d.sample() < 3 for _ in range(20))


def test_opnav_env_attributes_without_gpu():
    """opNavEnv mirrors opNavEnvironment.py:17-49 and is importable from the package and from `envs` (envs/__init__.py:2)."""
    import basilisk_env_b200 as b
    from basilisk_env_b200.envs import opNavEnv
    from basilisk_env_b200.opnav_env import configure_initial_conditions, elem2rv, MU_MARS
    from basilisk_env_b200.vec_env import BskEnvError
    assert b.opNavEnv is opNavEnv
    env = opNavEnv()
    assert env.max_length == 40 and env.step_duration == 50. and env.reward_mult == 1.
    assert env.observation_space.shape == (4, 1) and env.action_space.n == 2 and env.obs.shape == (4,)
    with pytest.raises(BskEnvError):
        env.step(0)
    np.random.seed(3)
    row = configure_initial_conditions()
    np.random.seed(3)
    np.testing.assert_array_equal(row[6:9], np.random.uniform(100000, -100000, 3))
    np.testing.assert_array_equal(row[9:12], np.random.uniform(1000, -1000, 3))
    from oracle import opnav as on
    rN, vN = on.reference_orbit()
    np.testing.assert_allclose(row[:3], rN, rtol=1e-15); np.testing.assert_allclose(row[3:6], vN, rtol=1e-15)


def test_env_attributes_without_gpu():
    from basilisk_env_b200.envs import leoPowerAttEnv, _decode_action
    from basilisk_env_b200.vec_env import BskEnvError
    env = leoPowerAttEnv()
    assert env.max_length == 540 and env.step_duration == 180. and env.power_max == 20.0
    assert abs(env.wheel_limit - 3000 * 2 * np.pi / 60) < 1e-12
    assert env.reward_mult == 1. / 540 and env.failure_penalty == 1
    assert env.observation_space.shape == (5, 1) and env.action_space.n == 3
    with pytest.raises(BskEnvError):
        env.step(0)                                  # quirk Q7: step before reset raises
    assert [_decode_action(a) for a in (0, 1, 2, np.int64(2), "1", 3, -1, 0.0, None)] == [0, 1, 2, 2, 1, -1, -1, -1, -1]


def test_ic_draw_order_matches_reference_stream(orc):
    """np.random.seed(s) + set_ICs consumes the global legacy stream in the reference's order:
    5 orbit draws, 3+3 attitude, 3 normals, 3 wheel speeds, 1 charge (+3 discarded by the wheel factory)."""
    from basilisk_env_b200 import initial_conditions as icm
    np.random.seed(12345)
    ic = icm.set_ICs()
    icm.consume_wheel_factory_draws()
    after = np.random.uniform()
    np.random.seed(12345)
    e = np.random.uniform(0, 0.05, 1); i = np.random.uniform(-np.pi / 2, np.pi / 2, 1)
    Om = np.random.uniform(0, 2 * np.pi, 1); om = np.random.uniform(0, 2 * np.pi, 1); f = np.random.uniform(0, 2 * np.pi, 1)
    sig = np.random.uniform(0, 1.0, 3); w = np.random.uniform(-1e-5, 1e-5, 3)
    dist = np.random.standard_normal(3); wheels = np.random.uniform(-800, 800, 3)
    charge = np.random.uniform(8 * 3600., 20 * 3600., 1)[0]
    np.random.uniform(-800, 800, 3)
    assert after == np.random.uniform()
    assert ic["oe"].e[0] == e[0] and ic["oe"].i[0] == i[0] and ic["oe"].Omega[0] == Om[0] and ic["oe"].omega[0] == om[0] and ic["oe"].f[0] == f[0]
    np.testing.assert_array_equal(ic["sigma_init"], sig); np.testing.assert_array_equal(ic["omega_init"], w)
    np.testing.assert_array_equal(ic["disturbance_vector"], dist); np.testing.assert_array_equal(ic["wheelSpeeds"], wheels)
    assert ic["storedCharge_Init"] == charge
    # and the oracle-side sampler (used to make fixtures) draws the same numbers
    d = orc.sample_ic_dict(np.random.RandomState(12345))
    np.testing.assert_allclose(icm.ic_row(ic), orc.ic_to_row(d), rtol=1e-15, atol=1e-9)
    assert set(icm.config_overrides(ic)) == set(icm.CONFIG_KEYS)
    assert abs(np.linalg.norm(ic["rN"]) - 6871e3) < 0.05 * 6871e3 + 1


def test_shard_range_partitions_every_env_once():
    from basilisk_env_b200.vec_env import shard_range
    for total in (1, 7, 4096, 2**20, 2**20 + 5):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[k][1] == spans[k + 1][0] for k in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_device_sampler_distributions(hostcore):
    """The counter-based device sampler (host-compiled here) draws from the reference's distributions
    and is keyed by (seed, global env index, episode): shard-invariant and reproducible."""
    hc = hostcore.HostCore(2000)
    ics, obs = hc.reset_seeded(seed=7, first_env=0)
    r = np.linalg.norm(ics[:, 0:3], axis=1); a = 6871e3
    assert np.all(r > a * (1 - 0.05) - 1) and np.all(r < a * (1 + 0.05) + 1)
    v = np.linalg.norm(ics[:, 3:6], axis=1)
    energy = 0.5 * v**2 - 0.3986004415e15 / r
    np.testing.assert_allclose(energy, -0.3986004415e15 / (2 * a), rtol=1e-12)
    assert np.all((ics[:, 6:9] >= 0) & (ics[:, 6:9] < 1)) and abs(ics[:, 6:9].mean() - 0.5) < 0.02
    assert np.all(np.abs(ics[:, 9:12]) <= 1e-5)
    assert abs(ics[:, 12:15].mean()) < 0.05 and abs(ics[:, 12:15].std() - 1) < 0.05
    assert np.all(np.abs(ics[:, 15:18]) <= 800) and abs(ics[:, 15:18].std() - 1600 / np.sqrt(12)) < 20
    assert np.all((ics[:, 18] >= 8 * 3600) & (ics[:, 18] <= 20 * 3600))
    inc_z = np.cross(ics[:, 0:3], ics[:, 3:6])[:, 2]
    assert np.all(inc_z >= -1e-3 * np.abs(inc_z).max())       # |i| <= 90 deg: prograde
    # shard invariance: envs [500, 600) sampled as their own shard are the same numbers
    hc2 = hostcore.HostCore(100)
    ics2, _ = hc2.reset_seeded(seed=7, first_env=500)
    np.testing.assert_array_equal(ics2, ics[500:600])
    ics3, _ = hostcore.HostCore(100).reset_seeded(seed=8, first_env=500)
    assert not np.array_equal(ics3, ics2)
    # initial observation: ENV:188-190 normalisation of SIM:347-351 (wheel norm in RPM over rad/s limit: quirk kept)
    np.testing.assert_allclose(obs[:, 2], np.linalg.norm(ics[:, 15:18], axis=1) / (3000 * 2 * np.pi / 60), rtol=1e-15)
    np.testing.assert_allclose(obs[:, 3], ics[:, 18] / 3600. / 20., rtol=1e-15)
    assert np.all(obs[:, 4] == 0)


def test_unused_reference_ic_helpers():
    """leo_orbit.inclined_circular_300km (leo_orbit.py:6-23) and sc_attitudes.static_inertial (sc_attitudes.py:15-23)."""
    from basilisk_env_b200 import initial_conditions as icm
    oe, r, v = icm.inclined_circular_300km()
    a = 6671e3
    assert oe.e == 0.0 and abs(np.linalg.norm(r) - a) < 1e-6
    np.testing.assert_allclose(np.linalg.norm(v), np.sqrt(icm.MU_EARTH / a), rtol=1e-14)       # circular speed
    np.testing.assert_allclose(np.cross(r, v) / np.linalg.norm(np.cross(r, v)), [0.0, -np.sin(np.pi / 4), np.cos(np.pi / 4)], atol=1e-15)
    s, w = icm.static_inertial()
    assert s.shape == (3,) and w.shape == (3,) and not s.any() and not w.any()
