"""CPU suite: the N>1 host path with world_size-2 gloo -- env sharding by global index and the one
optional collective (all-reduce of the eight episode statistics)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from basilisk_env_b200.vec_env import shard_range, all_reduce_stats
    from tests import hostcore_binding as hb
    lo, hi = shard_range(total, rank, world)
    hc = hb.HostCore(hi - lo, step_duration=5.0, max_length=2)
    ics, _ = hc.reset_seeded(seed=99, first_env=lo)
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 3, size=(3, total))[:, lo:hi]
    rets = np.zeros(hi - lo); finished = 0
    for t in range(3):
        obs, rew, done, reason = hc.step(acts[t])
        rets += rew; finished += int(done.sum())
    local = np.array([rets.sum(), 3.0 * (hi - lo), finished, 0, 0, 0, finished, 3.0 * (hi - lo)])
    glob = all_reduce_stats(local)
    q.put((rank, lo, hi, ics, obs, rets, glob))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_shard():
    total, world = 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from tests import hostcore_binding as hb
    hc = hb.HostCore(total, step_duration=5.0, max_length=2)
    ics, _ = hc.reset_seeded(seed=99, first_env=0)
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 3, size=(3, total))
    rets = np.zeros(total); finished = 0
    for t in range(3):
        obs, rew, done, reason = hc.step(acts[t])
        rets += rew; finished += int(done.sum())
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]
    np.testing.assert_array_equal(np.concatenate([r[3] for r in res]), ics)       # same ICs whatever the sharding
    np.testing.assert_array_equal(np.concatenate([r[4] for r in res]), obs)       # bit-identical trajectories
    np.testing.assert_array_equal(np.concatenate([r[5] for r in res]), rets)
    for r in res:                                                                   # the all-reduced statistics
        assert r[6][1] == 3.0 * total and r[6][2] == finished and abs(r[6][0] - rets.sum()) < 1e-12
