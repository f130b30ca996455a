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


# --------------------------------------------------------------------------------------------------
# opNav fixture (tests/golden/opnav_batch8.npz, generator make_golden_opnav.py)
# --------------------------------------------------------------------------------------------------
def test_opnav_oracle_reproduces_its_fixture():
    from oracle import opnav as on
    g = np.load(os.path.join(GOLDEN, "opnav_batch8.npz"))
    assert list(g["actions"][:, 0]) == [1, 1, 0, 0, 1, 1, 1, 0, 0, 1]      # the reference's recorded actHist (ONS:327)
    rN, vN = on.reference_orbit()
    np.testing.assert_array_equal(g["ics"][0, :6], np.concatenate([rN, vN]))
    for tag, kw in (("ref", dict()), ("cam", dict(camera_reenable=1))):
        batch = on.OpNavEnvBatch(g["ics"], on.default_cfg(seed=int(g["seed"]), **kw), first_env_index=int(g["first_env"]))
        for t in range(len(g["actions"])):
            o, r, d, w, dbg = batch.step(g["actions"][t], nthreads=2)
            np.testing.assert_allclose(o, g[f"{tag}_obs"][t], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(r, g[f"{tag}_reward"][t], rtol=1e-12)
            np.testing.assert_array_equal(d, g[f"{tag}_done"][t]); np.testing.assert_array_equal(w, g[f"{tag}_reason"][t])
            np.testing.assert_array_equal([s.n_meas for s in batch.states()], g[f"{tag}_n_meas"][t])
    # reference semantics: the camera never comes back after the first action 1 -> env 0 never measures
    assert g["ref_n_meas"][-1, 0] == 0 and g["cam_n_meas"][-1, 0] > 100


def test_opnav_hostcore_replays_fixture():
    from tests import opnav_parity as par
    from tests.hostcore_binding import HostCoreOpNav
    g = np.load(os.path.join(GOLDEN, "opnav_batch8.npz"))
    n = len(g["ics"])
    for tag, kw in (("ref", dict()), ("cam", dict(camera_reenable=1))):
        hc = HostCoreOpNav(n, first_env=int(g["first_env"]), noise_seed=int(g["seed"]), **kw)
        hc.reset_ics(g["ics"])
        for t in range(len(g["actions"])):
            obs, rew, done, reason, dbg = hc.step(g["actions"][t])
            for e in range(n):
                par.compare_obs(obs[e], g[f"{tag}_obs"][t, e], f"{tag} step {t} env {e}")
                par.compare_debug(dbg[e], g[f"{tag}_debug"][t, e], f"{tag} step {t} env {e}")
            np.testing.assert_allclose(rew, g[f"{tag}_reward"][t], rtol=1e-12, atol=1e-14)
            np.testing.assert_array_equal(done, g[f"{tag}_done"][t])
            S, I = hc.state()
            np.testing.assert_array_equal(I[par.F("n_meas")], g[f"{tag}_n_meas"][t])
        np.testing.assert_allclose(S[par.F("filter_state"):par.F("filter_state") + 3].T, g[f"{tag}_filt_state"][:, :3], rtol=1e-9)
        np.testing.assert_allclose(S[par.F("Omega"):par.F("Omega") + 4].T, g[f"{tag}_Omega"], rtol=1e-9, atol=1e-9)
