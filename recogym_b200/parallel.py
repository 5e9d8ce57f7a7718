"""Multi-GPU plumbing: users shard embarrassingly.

Rank r of R simulates the contiguous user-id range shard_range(n, r, R); the RNG is keyed by the
global user id, so the union of the shards is bit-identical to a single-GPU run.  The only
collective is one all-reduce(sum) of the int64 counters (and optional histograms) at the end of a
rollout -- NCCL on GPUs, gloo in the CPU tests.
"""
from scipy.stats.distributions import beta


def shard_range(n_users, rank, world_size, user_begin=0):
    """-> (first_user_id, count) of `rank`: contiguous, sizes differ by at most one, covers [0,n)."""
    n_users, rank, world_size = int(n_users), int(rank), int(world_size)
    assert 0 <= rank < world_size and n_users >= 0
    base, rem = divmod(n_users, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return user_begin + first, count


def allreduce_counters(*tensors, group=None):
    """In-place sum over ranks of int64 counter tensors (no-op when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def dist_info(group=None):
    """-> (rank, world_size); (0, 1) when torch.distributed is not initialised"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def resolve_carry(carries, rank, initial):
    """BanditCount's `last_product_viewed` across the shards of ONE training log (statistics reduced at
    step boundaries).  Every rank derives its columns with "no previous view" (-1) and reports the view its last train
    call left behind (-1: none of its calls carried a session).  The reference trains the users in id order, so the
    view rank r starts from is the last one reported by a LOWER rank, else `initial` (the agent's state before the
    log); the state after the log is the last one reported by ANY rank.  -> (previous view of `rank`, final view)"""
    prev = final = int(initial)
    for r, c in enumerate(carries):
        if int(c) >= 0:
            if r < rank:
                prev = int(c)
            final = int(c)
    return prev, final


def allgather_ints(values, device, group=None):
    """values: list of python ints of this rank -> [world][len(values)] python ints (one small all-gather)"""
    import torch
    import torch.distributed as dist
    rank, world = dist_info(group)
    mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    if world == 1:
        return [mine.tolist()]
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [t.tolist() for t in out]


def ctr_quantiles(successes, failures):
    """Beta(s+1, f+1) median and 95% credible interval (bench_agents.py:247-251)."""
    return (beta.ppf(0.500, successes + 1, failures + 1), beta.ppf(0.025, successes + 1, failures + 1),
            beta.ppf(0.975, successes + 1, failures + 1))


def sharded_rollout(device_env, n_users, policy="uniform", table=None, user_begin=0, math_mode=1, hist=False,
                    group=None):
    """Rollout of users [user_begin, user_begin+n_users) split over the ranks of `group`; returns the
    all-reduced totals (and histograms) as device tensors, identical on every rank."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_range(n_users, rank, world, user_begin)
    r = device_env.rollout_device(count, policy, table, first, 0, math_mode, per_user=False, hist=hist)
    red = [r["totals"]] + ([r["view_hist"], r["action_hist"], r["click_hist"]] if hist else [])
    allreduce_counters(*red, group=group)
    return r
