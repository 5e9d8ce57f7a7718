"""CPU (gloo, world_size 2): the N>1 host logic -- sharding covers every user exactly once and the
counter all-reduce sums rank-local totals.  The per-rank 'kernel' is replaced by the oracle, which
shares the user-keyed RNG property the sharding relies on."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from recogym_b200.parallel import shard_range, ctr_quantiles


def test_shard_ranges_partition_users():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            got = [shard_range(n, r, world, user_begin=11) for r in range(world)]
            assert got[0][0] == 11 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    q = ctr_quantiles(10, 990)
    assert q[1] < q[0] < q[2]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from conftest import load_golden, oracle_env_from_golden
    from oracle import oracle as O
    from recogym_b200.parallel import allreduce_counters, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("t2_seed42_P10_K5_u100")
    env = oracle_env_from_golden(g)
    first, count = shard_range(100, rank, world)
    r = O.rollout(env, first, count, hist=True)
    tot = torch.from_numpy(r["totals"].copy())
    vh = torch.from_numpy(r["view_hist"].copy())
    allreduce_counters(tot, vh)
    q.put((rank, tot.numpy().tolist(), vh.numpy().tolist(), r["hash"].tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_run():
    import torch.multiprocessing as mp
    from conftest import load_golden, oracle_env_from_golden
    from oracle import oracle as O
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = load_golden("t2_seed42_P10_K5_u100")
    full = O.rollout(oracle_env_from_golden(g), 0, 100, hist=True)
    for rank, tot, vh, _ in res:
        assert tot == full["totals"].tolist() and vh == full["view_hist"].tolist()
    assert res[0][3] + res[1][3] == full["hash"].tolist()     # concatenated shards == unsharded run
