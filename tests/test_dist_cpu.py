"""world_size-2 gloo run of the multi-rank plumbing (shard ranges, statistics all-reduce, policy broadcast)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openroborl_b200 import dist as odist


def test_shard_range_partitions_exactly():
    for total, world in ((65536, 8), (4096, 1), (10, 4), (3, 8)):
        rng = [odist.shard_range(total, r, world) for r in range(world)]
        assert rng[0][0] == 0 and rng[-1][1] == total
        assert all(rng[i][1] == rng[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in rng]
        assert max(sizes) - min(sizes) <= 1
    assert odist.shard_range(65536, 3, 8) == (3 * 8192, 4 * 8192)
    with pytest.raises(ValueError):
        odist.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stats = odist.allreduce_stats([1.0 + rank, 10.0 * (rank + 1), 2.0, 100.0])
        params = [torch.full((3, 2), float(rank)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
        synced_before = odist.check_synced(params)
        odist.broadcast_policy(params, src=0)
        synced_after = odist.check_synced(params)
        # per-shard oracle envs keyed by GLOBAL robot ids reproduce the unsharded reset
        from tests import common
        a, b = odist.shard_range(6, rank, world)
        env = common.make_oracle("laikago", ["laikago_pace"], b - a, seed=5, gid0=a)
        obs = env.reset()
        q.put((rank, stats.tolist(), [p.tolist() for p in params], synced_before, synced_after, obs))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_collectives_and_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, stats, params, before, after, _ in res:
        assert stats == [3.0, 30.0, 4.0, 200.0]
        assert params[0] == [[0.0, 0.0]] * 3 and params[1] == [0.0, 1.0, 2.0, 3.0]       # rank 0's weights everywhere
        assert before is False and after is True
    from tests import common
    full = common.make_oracle("laikago", ["laikago_pace"], 6, seed=5).reset()
    assert np.array_equal(np.concatenate([res[0][5], res[1][5]]), full)
