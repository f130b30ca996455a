"""CPU suite: the committed golden fixtures (tests/golden/, generated from the oracle by
make_golden.py) against (a) the oracle as built now and (b) the host-compiled step core.
The GPU suite replays the same fixtures through the C ABI."""
import os

import numpy as np

from tests import parity

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_its_fixtures(orc):
    g = np.load(os.path.join(GOLDEN, "leo_batch16.npz"))
    batch = orc.LeoEnvBatch(g["ics"])
    np.testing.assert_array_equal(batch.obs0, g["ob0"])
    for t in range(len(g["actions"])):
        o, r, d, w = batch.step(g["actions"][t], nthreads=2)
        np.testing.assert_allclose(o, g["obs"][t], rtol=1e-12, atol=1e-15)
        np.testing.assert_array_equal(d, g["done"][t]); np.testing.assert_array_equal(w, g["reason"][t])
    fire = np.array([e.state().thr_fire_count[:] for e in batch.envs])
    np.testing.assert_array_equal(fire, g["fire"])
    assert fire.sum() > 0


def test_fixture_ics_come_from_the_reference_seed(orc):
    """ICs of the episode fixtures = numpy legacy stream seeded with 12345 (the seed of the reference demo, ENV:225)."""
    rng = np.random.RandomState(12345)
    ic = orc.ic_to_row(orc.sample_ic_dict(rng))
    g = np.load(os.path.join(GOLDEN, "leo_episode_const0.npz"))
    np.testing.assert_array_equal(ic, g["ic"])
    assert len(g["actions"]) == len(g["obs"]) and bool(g["done"][-1]) and not g["done"][:-1].any()


def _replay(hostcore, name):
    g = np.load(os.path.join(GOLDEN, name))
    hc = hostcore.HostCore(1)
    ob0 = hc.reset_ics(g["ic"][None, :])[0]
    np.testing.assert_array_equal(ob0, g["ob0"])
    for t, a in enumerate(g["actions"]):
        obs, rew, done, reason = hc.step([a])
        parity.compare_obs(obs[0], g["obs"][t], f"{name} step {t}")
        assert abs(rew[0] - g["reward"][t]) <= 1e-12
        assert bool(done[0]) == bool(g["done"][t]) and int(reason[0]) == int(g["reason"][t]), f"{name} step {t}"
    S, I = hc.state()
    np.testing.assert_allclose(S[0:3, 0], g["final_r"], rtol=1e-9)
    np.testing.assert_allclose(S[3:6, 0], g["final_v"], rtol=1e-9)
    assert int(I[parity.F("MRPSwitchCount"), 0]) == int(g["final_switch"])
    np.testing.assert_array_equal(I[parity.F("fireCounter"):parity.F("fireCounter") + 8, 0], g["final_fire"])
    return g


def test_hostcore_replays_constant_zero_episode(hostcore):
    g = _replay(hostcore, "leo_episode_const0.npz")
    assert g["reason"][-1] & 4                       # always-nadir drains the battery: "ran out of power"


def test_hostcore_replays_random_full_episode(hostcore):
    g = _replay(hostcore, "leo_episode_random.npz")
    assert len(g["obs"]) == 541 and g["reason"][-1] == 1     # quirk Q9: the 541st call ends the episode
