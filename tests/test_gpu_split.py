"""GPU suite for the two-warp ("split") organisation of the LEO step kernel (csrc/leo_split.cuh; include/bskenv.h:
bskenv_set_organisation): a dynamics warp and a companion warp (flight software + EnvTask) per group of 32 envs, the
small-batch organisation behind BASELINE configs[1] (4096 envs on one B200).

Both organisations run the same arithmetic, so the bar is (i) the oracle, at the same tolerances as tests/test_gpu_parity.py
(discrete exact, continuous <= 1e-9), for BOTH organisations explicitly -- batches of a few groups are routed to the split
kernel automatically, so the one-thread kernel is pinned here as well -- and (ii) each other, BIT FOR BIT, at 4096 / 8192 /
16384 envs (one, two and four groups per block) with every mode, in-kernel auto-reset and the per-env episode record.
Reference for what a step is: LEOPowerAttitudeSimulator.run_sim (leoPowerAttitudeSimulator.py:535-644) +
leoPowerAttEnv.step (leoPowerAttitudeEnvironment.py:65-145)."""
import numpy as np
import pytest

from tests import parity
from tests.test_gpu_parity import _run_against_oracle, _state_np, _vec

pytestmark = pytest.mark.gpu
ORGS = ("thread", "split")


@pytest.mark.parametrize("org", ORGS)
def test_both_organisations_match_the_oracle(bsk, orc, org):
    """64 envs x 8 decision steps, all three modes, eight envs with wheels fast enough for the desat chain to fire thrusters."""
    rows = parity.sample_rows(orc, 64, seed=21)
    rows[:8, 15:18] = np.random.RandomState(22).uniform(1500, 2900, size=(8, 3)) * np.array([1, -1, 1])
    acts = np.random.RandomState(23).randint(0, 3, size=(8, 64))
    acts[:3, :8] = 2
    _run_against_oracle(bsk, orc, rows, acts, organisation=org)


@pytest.mark.parametrize("org", ORGS)
def test_both_organisations_match_the_oracle_on_the_stress_config(bsk, orc, org):
    rows = parity.sample_rows(orc, 32, seed=24)
    rows[:6, 15:18] = np.random.RandomState(25).uniform(1500, 2900, size=(6, 3)) * np.array([1, -1, 1])
    acts = np.random.RandomState(26).randint(0, 3, size=(4, 32))
    acts[:2, :6] = 2
    _run_against_oracle(bsk, orc, rows, acts, use_j2=1, rw_set=1, organisation=org)


@pytest.mark.parametrize("org", ORGS)
def test_both_organisations_other_rates(bsk, orc, org):
    """Unchunked interval (100 ticks), flight software EVERY tick (dynRate = fswRate: the FSW hand-off runs each tick) and a
    coarse dynamics rate, unknown actions included."""
    rows = parity.sample_rows(orc, 32, seed=27)
    acts = np.random.RandomState(28).randint(-1, 4, size=(3, 32))
    _run_against_oracle(bsk, orc, rows, acts, step_duration=10.0, organisation=org)
    _run_against_oracle(bsk, orc, rows, acts, step_duration=20.0, dynRate=1.0, fswRate=1.0, organisation=org)
    _run_against_oracle(bsk, orc, rows, acts, step_duration=30.0, dynRate=0.5, fswRate=1.0, organisation=org)


def _rollout(bsk, n, org, steps, **kw):
    import torch
    env = _vec(bsk, n, seed=7, auto_reset=True, max_length=3, organisation=org, **kw)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = torch.randint(0, 3, (steps, n), dtype=torch.int32, device="cuda", generator=g)
    out = []
    for t in range(steps):
        o, r, d, info = env.step(acts[t])
        out.append([x.clone() for x in (o, r, d, info["done_reason"], info["terminal_obs"], info["episode_r"], info["episode_l"])])
    S, I = env.get_state()
    name, stats = env.kernel_name(), env.episode_stats()
    env.close()
    return out, S, I, name, stats


@pytest.mark.parametrize("n,kw", [(4096, {}), (8192, {}), (16384, {}), (4096, dict(use_j2=1, rw_set=1)), (2048, dict(step_duration=10.0)),
                                  (4000, {}), (33, dict(step_duration=30.0)), (1, dict(step_duration=30.0)), (95, dict(use_j2=1, rw_set=1, step_duration=30.0))])
def test_split_equals_thread_bit_for_bit(bsk, n, kw):
    """BASELINE configs[1] size and its multiples (1, 2, 4 groups per block): random actions over all modes, episodes of at
    most four steps so that freshly reset envs (tick 0 runs) and running ones share warps, auto-reset inside the launch; and
    ragged batches (4000, 95, 33, 1 envs: the spare lanes of the last group step a duplicate of the last env)."""
    import torch
    a, Sa, Ia, ka, sta = _rollout(bsk, n, "thread", 6, **kw)
    b, Sb, Ib, kb, stb = _rollout(bsk, n, "split", 6, **kw)
    assert ka.startswith("leo_step_kernel") and kb.startswith("leo_split_kernel"), (ka, kb)
    ended = 0
    for t, (x, y) in enumerate(zip(a, b)):
        for k, (p, q) in enumerate(zip(x, y)):
            assert torch.equal(p, q), f"step {t} output {k}: max |diff| {(p.double() - q.double()).abs().max().item():.3e}"
        ended += int(x[2].sum())
    assert torch.equal(Sa, Sb) and torch.equal(Ia, Ib)
    assert ended >= n, ended
    for k in sta:                   # counts are exact; the two sums are accumulated with atomics (order differs between launches)
        assert sta[k] == stb[k] if k not in ("return_sum", "length_sum") else abs(sta[k] - stb[k]) <= 1e-9 * max(1.0, abs(sta[k])), (k, sta, stb)


def test_automatic_selection_and_ragged_batches(bsk):
    """`auto` picks the split kernel for batches that fit four groups of 32 per SM -- ragged ones included: the spare lanes of
    the last group step a duplicate of the last env (leo_pad_kernel) --, the one-thread kernel otherwise; asking for `split` on
    a batch that is too large is an error of the step call (no silent fallback)."""
    import torch
    from basilisk_env_b200.vec_env import BskEnvError
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    big = 32 * (4 * sms + 1)
    for n, want in ((4096, "leo_split_kernel"), (64, "leo_split_kernel"), (33, "leo_split_kernel"), (1, "leo_split_kernel"), (big, "leo_step_kernel")):
        env = _vec(bsk, n, seed=1, step_duration=10.0)
        env.reset()
        env.step(torch.zeros(n, dtype=torch.int32, device="cuda"))
        assert env.kernel_name().startswith(want), (n, env.kernel_name())
        env.close()
    env = _vec(bsk, big, seed=1, step_duration=10.0, organisation="split")
    env.reset()
    with pytest.raises(BskEnvError, match="at most"):
        env.step(torch.zeros(big, dtype=torch.int32, device="cuda"))
    env.set_organisation("thread")
    env.step(torch.zeros(big, dtype=torch.int32, device="cuda"))
    with pytest.raises(BskEnvError):
        env.set_organisation(7)
    env.close()


def test_checkpoint_crosses_organisations(bsk):
    """A state saved under one organisation continues bit-identically under the other (the persistent state is the same)."""
    import torch
    n = 128
    acts = torch.randint(0, 3, (4, n), dtype=torch.int32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    ref = _vec(bsk, n, seed=9, step_duration=30.0, organisation="thread")
    ref.reset()
    a = _vec(bsk, n, seed=9, step_duration=30.0, organisation="split")
    a.reset()
    for t in range(2):
        ref.step(acts[t]); a.step(acts[t])
    b = _vec(bsk, n, seed=9, step_duration=30.0, organisation="thread")
    b.set_state(*a.get_state())
    for t in range(2, 4):
        o0, r0, d0, _ = ref.step(acts[t]); o1, r1, d1, _ = b.step(acts[t])
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1)
    S0, I0 = _state_np(ref); S1, I1 = _state_np(b)
    np.testing.assert_array_equal(S0, S1); np.testing.assert_array_equal(I0, I1)
    for e in (ref, a, b):
        e.close()


def test_foreign_tick_counts_take_the_voted_schedule(bsk):
    """Envs of one warp whose tick counts are NOT congruent modulo the flight-software period (only reachable by injecting a
    foreign state with bskenv_set_state): the flight-software passes of a group then fall on different ticks per lane, the
    split kernel's barrier schedule switches from the uniform rule to warp votes, and the result is still what the
    one-thread kernel computes, bit for bit."""
    import torch
    from basilisk_env_b200 import _native
    n = 96
    acts = torch.randint(0, 3, (3, n), dtype=torch.int32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))
    src = _vec(bsk, n, seed=13, step_duration=30.0, organisation="thread")
    src.reset()
    src.step(acts[0])
    S, I = src.get_state()
    I = I.clone()
    I[_native.state_field("tick")[0]] += torch.arange(n, device=I.device, dtype=I.dtype) % 7      # 0 .. 6 ticks ahead, per lane
    out = {}
    for org in ORGS:
        env = _vec(bsk, n, seed=13, step_duration=30.0, organisation=org)
        env.reset()
        env.set_state(S, I)
        res = []
        for t in (1, 2):
            o, r, d, info = env.step(acts[t])
            res += [o.clone(), r.clone(), d.clone(), info["done_reason"].clone()]
        out[org] = res + list(env.get_state()) + [env.kernel_name()]
        env.close()
    assert out["thread"][-1].startswith("leo_step_kernel") and out["split"][-1].startswith("leo_split_kernel")
    for k, (p, q) in enumerate(zip(out["thread"][:-1], out["split"][:-1])):
        assert torch.equal(p, q) or bool((torch.isnan(p.double()) == torch.isnan(q.double())).all() and torch.equal(torch.nan_to_num(p.double()), torch.nan_to_num(q.double()))), f"output {k}"
    src.close()
