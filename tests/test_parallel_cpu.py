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


def _chain_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from recogym_b200.parallel import allgather_ints, resolve_carry, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for case, (views, initial) in enumerate(_chain_cases()):
        first, count = shard_range(len(views), rank, world)
        mine = views[first:first + count]
        # the shard alone, derived with "no previous view" (-1), as every rank derives its lvp column
        lvp, cur = [], -1
        for v in mine:
            lvp.append(cur)
            if v >= 0:
                cur = v
        n_before = sum(c[0] for c in allgather_ints([len(mine)], "cpu")[:rank])      # the sliding windows' sample offset
        carries = [c[0] for c in allgather_ints([cur], "cpu")]
        prev, final = resolve_carry(carries, rank, initial)
        lvp = [prev if x == -1 else x for x in lvp]
        out.append((case, first, n_before, lvp, final))
    q.put((rank, out))
    dist.destroy_process_group()


def _chain_cases():
    """train calls of one log in user order: the view each call's session ends with, -1 = a call without a session (or a
    call a selective TrainingApproach dropped); and the agent's last viewed product before the log"""
    rs = np.random.RandomState(5)
    cases = [([3, -1, -1, 4, -1, 2, -1, -1], -1), ([-1, -1, -1, -1, -1, 7], 9), ([-1] * 6, 4), ([-1] * 5, -1)]
    for n in (2, 3, 17, 64):
        v = rs.randint(0, 10, n)
        v[rs.rand(n) < 0.6] = -1
        cases.append((v.tolist(), int(rs.randint(-1, 10))))
    return cases


def test_last_view_chain_over_shards_equals_sequential_training():
    """BanditCount.last_product_viewed before every train call when the calls are sharded over two ranks (the
    multi-GPU evolution loop: evaluate_agent._evaluate_agent_device) == the sequential loop's; same for the state the
    log leaves behind and for the running sample count."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for case, (views, initial) in enumerate(_chain_cases()):
        want, cur = [], initial
        for v in views:
            want.append(cur)
            if v >= 0:
                cur = v
        got = []
        for r in range(2):
            c, first, n_before, lvp, final = res[r][case]
            assert c == case and first == n_before == len(got) and final == cur
            got += lvp
        assert got == want, (case, got, want)


def _dense_reduce_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from recogym_b200.logs import DenseCountStats
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = 6

    def stats(seed):
        s = DenseCountStats.__new__(DenseCountStats)      # CPU tensors: the reduction logic does not care where they live
        rs = np.random.RandomState(seed)
        s.P, s.device, s.last_view = P, 0, -1
        s.co_counts = None
        s.pulls = torch.from_numpy(rs.randint(0, 9, (P, P)).astype(np.int64))
        s.clicks = torch.from_numpy(rs.randint(0, 3, (P, P)).astype(np.int64))
        return s

    mine, delta = stats(100), stats(rank)                 # replicated statistics; this rank's shard of the step
    mine.add_reduced(delta)
    q.put((rank, mine.pulls.numpy().tolist(), mine.clicks.numpy().tolist()))
    dist.destroy_process_group()


def test_dense_statistics_reduce_to_the_sum_of_the_shards():
    """DenseCountStats.add_reduced (the step-boundary all-reduce of the multi-GPU evolution loop): every rank ends with
    the replicated statistics plus the SUM of all ranks' shard statistics."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dense_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (a, b)) for r, a, b in (q.get(timeout=120) for _ in procs))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    P = 6
    def both(seed):
        rs = np.random.RandomState(seed)
        return rs.randint(0, 9, (P, P)).astype(np.int64), rs.randint(0, 3, (P, P)).astype(np.int64)
    base, d0, d1 = both(100), both(0), both(1)
    for r in range(2):
        assert np.array_equal(np.array(res[r][0]), base[0] + d0[0] + d1[0])
        assert np.array_equal(np.array(res[r][1]), base[1] + d0[1] + d1[1])
