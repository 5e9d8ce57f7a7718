"""Device-resident logs and what consumes them.

`generate_logs_device` is AbstractEnv.generate_logs (reference: recogym/envs/abstract.py:241-327) kept on the GPU: the
fused rollout, then an exclusive scan + one scatter put the rows into the reference's row order and a column writer
produces the DataFrame columns plus the per-row context (acting last view, session position) the consumers need:

  * `CountStats`  -- OrganicCount.train / BanditCount.train over a whole log as scatter-adds, and the agents' acting
    tables by a row arg-max (reference: agents/organic_count.py:54-87, agents/bandit_count.py:28-62);
  * `ips_table_policy` -- evaluate_IPS / evaluate_SNIPS of a table policy (reference: evaluate_agent.py:753-835).

torch supplies the device buffers; every computation is a kernel of librecogym_b200.so behind the C ABI.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .envs.device_env import _policy_struct

_COLS = (("t", "float32"), ("u", "int32"), ("z", "int8"), ("v", "int32"), ("a", "int32"), ("c", "float32"),
         ("ps", "float64"), ("lv", "int32"), ("lvp", "int32"), ("sess", "int32"), ("flags", "int8"))


def _torch():
    import torch
    return torch


class DeviceLog:
    """Columns of one generate_logs() call on the device (torch tensors, first `rows` entries valid)."""

    def __init__(self, env, capacity):
        torch = _torch()
        self.env = env
        self.capacity = int(capacity)
        self.rows = 0
        for name, dt in _COLS:
            setattr(self, name, torch.empty(self.capacity, dtype=getattr(torch, dt), device=env.torch_device))
        self.row_offset = None
        self.last_view = None
        self.totals = None

    def struct(self):
        s = L.LogColumns()
        s.capacity = self.capacity
        for name, _ in _COLS:
            setattr(s, name, getattr(self, name).data_ptr())
        return s

    def to_host(self):
        """dict of numpy columns (u, t, z, v, a, c as rollout_log returns them, plus ps / lv / lvp / sess)."""
        n = self.rows
        out = {name: getattr(self, name)[:n].cpu().numpy() for name, _ in _COLS}
        out["t"] = out["t"].astype(np.int32)
        c = out["c"]
        out["c"] = np.where(np.isnan(c), -1, c).astype(np.int8)
        return out

    def to_dataframe(self, zero_based_users=False):
        """The generate_logs DataFrame: t f32, u/v/a Int32 (nullable), z str, c f32, ps, ps-a; `u` = the global user id, or
        (zero_based_users, what the reference's per-call labels are) the id minus the call's user_begin.  Built from whole columns:
        the nullable integer columns straight from (values, mask), the `z` strings by decoding a two-entry Arrow
        dictionary (a numpy unicode array costs pandas 0.5 us per row to convert: 5.6 of the 6.5 s of a 10M-row frame)."""
        import pandas as pd
        n = self.rows
        z = self.z[:n].cpu().numpy()
        organic = z == 0
        bandit = ~organic
        try:
            import pyarrow as pa
            zcol = pd.array(pa.DictionaryArray.from_arrays(pa.array(z), pa.array(["organic", "bandit"])).dictionary_decode(),
                            dtype="str")
        except Exception:       # no pyarrow (or an object-string pandas): the plain numpy route
            zcol = np.where(organic, "organic", "bandit")
        none = np.zeros(n, dtype=bool)
        df = pd.DataFrame({
            "t": self.t[:n].cpu().numpy(),
            "u": pd.arrays.IntegerArray(self.u[:n].cpu().numpy() - np.int32(self.user_begin if zero_based_users else 0), none),
            "z": zcol,
            "v": pd.arrays.IntegerArray(self.v[:n].cpu().numpy(), bandit),
            "a": pd.arrays.IntegerArray(self.a[:n].cpu().numpy(), organic),
            "c": self.c[:n].cpu().numpy(),
            "ps": self.ps[:n].cpu().numpy(),
        })
        psa = np.empty(n, dtype=object)
        psa[:] = None
        empty = np.empty(1, dtype=object)
        empty[0] = ()
        psa[~organic] = empty
        df["ps-a"] = psa
        return df


FLAG_PHANTOM = 1     # csrc/rg_logs.cu: the bandit row after `done`


class MaskedTrainingLog:
    """A DeviceLog seen through one of evaluate_agent's selective TrainingApproaches (evaluate_agent.py:105-119): only
    the train calls of SOME bandit rows of an evolution step are made -- `not action['greedy']` (ALL_EXPLORATION_DATA),
    `samples % window == 0` (SLIDING_WINDOW_ALL_DATA), or both.  A train call owns the organic session right before its
    bandit row, so the rows to drop are marked in a COPY of the z column (value 2: neither organic nor bandit for the
    scatter-add kernels), and BanditCount's `last_product_viewed` -- which only moves when train() is called with a
    non-empty session -- is re-derived for the kept rows (lvp column) and for the end of the log (carry).  Index plumbing
    on the device (cumulative sums / maxima over the log columns); the statistics kernels run unchanged.

    greedy action of a row = acting_table[lv]; `samples` = evaluate_agent's running count of bandit steps before this
    log.  After construction: .n_samples (count after this log), .keep_rows (kept train calls)."""

    def __init__(self, log, acting_table, explored_only, window, samples, prev_last_view):
        torch = _torch()
        n = int(log.rows)
        dev = log.env.torch_device
        self.env, self.rows, self.train_mode, self.capacity = log.env, n, log.train_mode, log.capacity
        self._log = log
        z, flags, a, lv, u = log.z[:n], log.flags[:n], log.a[:n], log.lv[:n], log.u[:n]
        if not (hasattr(acting_table, "is_cuda") and acting_table.is_cuda):
            acting_table = torch.as_tensor(np.ascontiguousarray(acting_table, np.int32), device=dev)
        ar = torch.arange(n, device=dev)
        is_b = (z == 1) & ((flags & FLAG_PHANTOM) == 0)
        greedy = a == acting_table[lv.clamp(min=0).long()]
        ordinal = int(samples) + torch.cumsum(is_b.long(), 0)
        keep_b = is_b.clone()
        if explored_only:
            keep_b &= ~greedy
        if window:
            keep_b &= (ordinal % int(window)) == 0
        # organic rows follow the bandit row that closes their session (always inside the same user: a regular user's log
        # ends with its phantom bandit row)
        pos = torch.where(z == 1, ar, torch.full_like(ar, n))
        nxt = torch.flip(torch.cummin(torch.flip(pos, [0]), 0).values, [0]) if n else pos
        keep_o = (z == 0) & (nxt < n) & keep_b[nxt.clamp(max=max(n - 1, 0))]
        self.z = log.z.clone()
        self.z[:n] = torch.where(keep_b | keep_o | ((z == 1) & ~is_b), z, torch.full_like(z, 2))
        # last_product_viewed before each kept train call: the last view of the closest earlier KEPT call that came with a
        # non-empty session (BanditCount.train -> update_lpv, bandit_count.py:54-57); before the first one: the carry
        prev_org = torch.zeros(n, dtype=torch.bool, device=dev)
        if n > 1:
            prev_org[1:] = (z[:-1] == 0) & (u[1:] == u[:-1])
        cand = keep_b & prev_org
        run = torch.cummax(torch.where(cand, ar, torch.full_like(ar, -1)), 0).values if n else ar
        prev = torch.cat([torch.full((1,), -1, dtype=run.dtype, device=dev), run[:-1]]) if n else run
        self.lvp = log.lvp.clone()
        self.lvp[:n] = torch.where(prev >= 0, lv[prev.clamp(min=0)], torch.full_like(lv, int(prev_last_view))).to(log.lvp.dtype)
        last = int(run[-1].item()) if n else -1
        self.carry = torch.tensor([int(lv[last].item()) if last >= 0 else int(prev_last_view)], dtype=torch.int32, device=dev)
        self.n_samples = int(samples) + int(is_b.sum().item())
        self.keep_rows = int(keep_b.sum().item())
        self.sess, self.flags = log.sess, log.flags

    def struct(self):
        s = self._log.struct()
        s.z = self.z.data_ptr()
        s.lvp = self.lvp.data_ptr()
        return s


def user_views_ps(u, z, v, a, P):
    """'ps' of OrganicUserEventCounterAgent's logged actions (agents/organic_user_count.py:66-72 with epsilon = 0):
    ps = action_proba[a] = views of product a by this user so far / all its views so far, NaN on organic rows.  Running
    counts per (user, product) over ordered log columns (torch tensors, any device): a stable sort by the pair key,
    segmented cumulative sums (cumsum minus its value at the segment start, propagated by a running maximum), scattered
    back.  The columns kernel leaves this policy's ps undefined (it would need the per-user view lists again)."""
    torch = _torch()
    n = int(u.numel())
    organic = z == 0
    org = organic.to(torch.int64)
    x = torch.where(organic, v, a).to(torch.int64)
    key = u.to(torch.int64) * int(P) + x

    def segmented_cumsum(keys, vals):
        c = torch.cumsum(vals, 0)
        start = torch.ones(n, dtype=torch.bool, device=keys.device)
        if n > 1:
            start[1:] = keys[1:] != keys[:-1]
        base = torch.cummax(torch.where(start, c - vals, torch.zeros_like(c)), 0).values   # c - vals never decreases
        return c - base

    if n == 0:
        return torch.empty(0, dtype=torch.float64, device=u.device)
    order = torch.argsort(key, stable=True)
    seen = torch.empty(n, dtype=torch.int64, device=u.device)
    seen[order] = segmented_cumsum(key[order], org[order])
    total = segmented_cumsum(u.to(torch.int64), org)            # the log is ordered by user already
    ps = seen.to(torch.float64) / total.clamp(min=1).to(torch.float64)
    return torch.where(organic, torch.full_like(ps, float("nan")), ps)


def generate_logs_device(env, n_users, policy="uniform", table=None, ps_table=None, user_begin=0, n_organic_users=0,
                         math_mode=L.MATH_TENSOR, prev_last_view=-1, rows_per_user=108.0, log=None,
                         train_mode=L.TRAIN_TEST_AGENT, epsilon=0.0):
    """generate_logs(n_users, agent, n_organic_users) -> DeviceLog, everything on the GPU.  `env` is a DeviceEnv.
    One synchronisation (the row count) per call; a log that does not fit is regenerated into larger columns.
    train_mode (L.TRAIN_*) names the reference training loop whose train calls the derived columns describe."""
    torch = _torch()
    dev = env.torch_device
    n = int(n_users) + int(n_organic_users)
    if policy in ("table", "eps_table") and not (hasattr(table, "is_cuda") and table.is_cuda):
        table = torch.as_tensor(np.ascontiguousarray(table, np.int32), device=dev)
    if ps_table is not None and not (hasattr(ps_table, "is_cuda") and ps_table.is_cuda):
        ps_table = torch.as_tensor(np.ascontiguousarray(ps_table, np.float64), device=dev)
    cap = int(n * rows_per_user) + 4096
    keep = []
    pol = _policy_struct(keep, policy, table, env.P, host=False, epsilon=epsilon)
    while True:
        if log is None or log.capacity < cap:
            log = DeviceLog(env, cap)
        log.row_offset = torch.empty(n + 1, dtype=torch.int64, device=dev)
        log.last_view = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        log.totals = torch.zeros(16, dtype=torch.int64, device=dev)
        rows = torch.zeros(1, dtype=torch.int64, device=dev)
        log.carry = torch.zeros(1, dtype=torch.int32, device=dev)
        log.train_mode = int(train_mode)
        cols = log.struct()
        L.check(L.lib().rg_generate_logs_device(
            env._h, C.byref(pol), C.c_void_p(ps_table.data_ptr()) if ps_table is not None else None, int(user_begin),
            int(n_organic_users), int(n_users), int(prev_last_view), int(train_mode), C.c_void_p(log.carry.data_ptr()),
            C.byref(cols),
            C.c_void_p(log.row_offset.data_ptr()), C.c_void_p(log.last_view.data_ptr()), C.c_void_p(rows.data_ptr()),
            C.c_void_p(log.totals.data_ptr()), int(math_mode), env._stream()))
        need = int(rows.item())
        if need <= log.capacity:
            if policy == "user_views" and need:
                log.ps[:need] = user_views_ps(log.u[:need], log.z[:need], log.v[:need], log.a[:need], env.P)
            log.rows = need
            log.n_users_total = n
            log.n_organic_users = int(n_organic_users)
            log.user_begin = int(user_begin)
            return log
        cap = need + 1024
        log = None


SPARSE_MIN_PRODUCTS = 16384   # above this the [P][P] statistics (8 P^2 bytes each) give way to device hash tables


def CountStats(P, device=0, organic=True, bandit=True, sparse=None):
    """Sufficient statistics of OrganicCount (co_counts) and BanditCount (pulls, clicks) on the GPU: dense int64 [P][P]
    (`DenseCountStats`) or, for large catalogues, device hash tables (`SparseCountStats`; P = 100 000 would need
    3 x 80 GB dense).  Same interface and the same acting tables either way."""
    if sparse is None:
        sparse = int(P) >= SPARSE_MIN_PRODUCTS
    cls = SparseCountStats if sparse else DenseCountStats
    return cls(P, device, organic, bandit)


class DenseCountStats:
    """int64 [P][P] matrices (the reference holds float64 ones: agents/organic_count.py:49, agents/bandit_count.py:22-24).
    Holds no environment handle, so agents that own one stay deep-copyable / picklable."""

    sparse = False

    def __init__(self, P, device=0, organic=True, bandit=True):
        torch = _torch()
        self.P = int(P)
        self.device = int(device)
        dev = torch.device("cuda", self.device)
        self.co_counts = torch.zeros((self.P, self.P), dtype=torch.int64, device=dev) if organic else None
        self.pulls = torch.zeros((self.P, self.P), dtype=torch.int64, device=dev) if bandit else None
        self.clicks = torch.zeros((self.P, self.P), dtype=torch.int64, device=dev) if bandit else None
        self.last_view = -1     # BanditCount.last_product_viewed carried from one log to the next (-1 = None)

    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream)

    def train(self, log):
        """agent.train(...) for every train call the loop named by log.train_mode makes on `log`'s users, in the
        reference's order (bench_agents.py:169-190 / evaluate_agent.py:56-78,100-139).  The log must have been
        generated with prev_last_view = self.last_view."""
        cols = log.struct()
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        L.check(L.lib().rg_count_stats_device(C.byref(cols), int(log.rows), self.P, int(log.train_mode),
                                              p(self.co_counts), p(self.pulls), p(self.clicks), self._stream()))
        self.last_view = int(log.carry.item())

    def _policy(self, kind, A, B):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        table = torch.empty(self.P, dtype=torch.int32, device=dev)
        ps = torch.empty(self.P, dtype=torch.float64, device=dev)
        L.check(L.lib().rg_count_policy_device(kind, self.P, C.c_void_p(A.data_ptr()),
                                               C.c_void_p(B.data_ptr()) if B is not None else None,
                                               C.c_void_p(table.data_ptr()), C.c_void_p(ps.data_ptr()), self._stream()))
        return {"policy": "table", "table": table, "ps_table": ps}

    def zeros_like(self):
        return DenseCountStats(self.P, self.device, self.co_counts is not None, self.pulls is not None)

    def add_reduced(self, delta, group=None):
        """self += sum over ranks of `delta` (the statistics each rank collected on ITS shard of a training log): the
        count matrices are additive, so one all-reduce(sum) per matrix at the step boundary keeps every rank's copy equal to
        the single-GPU statistics (SURVEY 8e "not shardable ... deferred to (f)" / f1)."""
        import torch.distributed as dist
        for name in ("co_counts", "pulls", "clicks"):
            mine, d = getattr(self, name), getattr(delta, name)
            if mine is None:
                continue
            dist.all_reduce(d, op=dist.ReduceOp.SUM, group=group)
            mine.add_(d)

    def organic_count_policy(self):
        """OrganicCount.act as a device table (agents/organic_count.py:54-72)."""
        return self._policy(0, self.co_counts, None)

    def bandit_count_policy(self):
        """BanditCount.act as a device table, 'ps' = the smoothed click rate (agents/bandit_count.py:28-47)."""
        return self._policy(1, self.pulls, self.clicks)


class _DeviceHash:
    """One rg_sparse_counts table: keys / a / b [capacity], whole-row increments [P], stats[2] (device tensors)."""

    def __init__(self, P, capacity, device, bandit):
        torch = _torch()
        dev = torch.device("cuda", device)
        self.capacity = int(capacity)
        self.keys = torch.full((self.capacity,), -1, dtype=torch.int64, device=dev)       # 0xFF.. = empty
        self.a = torch.zeros(self.capacity, dtype=torch.int64, device=dev)
        self.b = torch.zeros(self.capacity, dtype=torch.int64, device=dev) if bandit else None
        self.row_a = torch.zeros(P, dtype=torch.int64, device=dev) if bandit else None
        self.row_b = torch.zeros(P, dtype=torch.int64, device=dev) if bandit else None
        self.stats = torch.zeros(2, dtype=torch.int64, device=dev)

    def struct(self):
        s = L.SparseCounts()
        s.capacity = self.capacity
        for n in ("keys", "a", "b", "row_a", "row_b", "stats"):
            t = getattr(self, n)
            setattr(s, n, t.data_ptr() if t is not None else None)
        return s

    def cells(self):
        """host dict {(row, col): (a, b)} (tests, host synchronisation of small tables)."""
        keys = self.keys.cpu().numpy()
        used = keys != -1
        k = keys[used].astype(np.uint64)
        a = self.a.cpu().numpy()[used]
        b = self.b.cpu().numpy()[used] if self.b is not None else np.zeros_like(a)
        return {(int(x >> np.uint64(32)), int(x & np.uint64(0xFFFFFFFF))): (int(y), int(z)) for x, y, z in zip(k, a, b)}


class SparseCountStats:
    """The same statistics in device hash tables (open addressing, key = row * 2^32 + col; csrc/rg_logs.cu): what the
    count agents need at P = 100 000 (BASELINE configs[3]).  Capacity is provisioned before every train(): room for the
    worst case of that log (one new cell per trained bandit row / per ordered pair of a session), growing the table by
    a device re-insertion when needed, so an insertion can never find the table full."""

    sparse = True

    def __init__(self, P, device=0, organic=True, bandit=True):
        self.P = int(P)
        self.device = int(device)
        self._want = (bool(organic), bool(bandit))
        self.co = None
        self.bandit = None
        self.last_view = -1

    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream)

    def _room(self, table, bandit, new_cells):
        """a table with at least 2 x (occupied + new_cells) slots: allocate, or grow by re-inserting the old cells"""
        used = int(table.stats[0].item()) if table is not None else 0
        need = 2 * (used + int(new_cells)) + 1024
        if table is not None and table.capacity >= need:
            return table
        cap = 1 << max(10, int(need - 1).bit_length())
        new = _DeviceHash(self.P, cap, self.device, bandit)
        if table is not None:
            _reinsert(table, new, self._stream())
        return new

    def train(self, log):
        torch = _torch()
        n = int(log.rows)
        organic, bandit = self._want
        if organic:
            z = log.z[:n]
            # ordered pairs a session of length s adds: s^2 (b b^T); the sum over views of (2 sess + 1)
            pairs = int((2 * log.sess[:n].long() + 1)[z == 0].sum().item()) if n else 0
            self.co = self._room(self.co, False, pairs)
        if bandit:
            self.bandit = self._room(self.bandit, True, n)
        cols = log.struct()
        co_s = self.co.struct() if organic else None
        ba_s = self.bandit.struct() if bandit else None
        L.check(L.lib().rg_sparse_count_stats_device(C.byref(cols), n, int(log.train_mode),
                                                     C.byref(co_s) if co_s is not None else None,
                                                     C.byref(ba_s) if ba_s is not None else None, self._stream()))
        for t in (self.co, self.bandit):
            if t is not None and int(t.stats[1].item()) != 0:
                raise RuntimeError("sparse count statistics: hash table overflow (capacity %d)" % t.capacity)
        self.last_view = int(log.carry.item())

    def _policy(self, kind, table_obj):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        table = torch.zeros(self.P, dtype=torch.int32, device=dev)
        ps = torch.full((self.P,), 1.0 if kind == 0 else 0.5, dtype=torch.float64, device=dev)
        if table_obj is not None:
            st = table_obj.struct()
            L.check(L.lib().rg_sparse_count_policy_device(kind, self.P, C.byref(st), C.c_void_p(table.data_ptr()),
                                                          C.c_void_p(ps.data_ptr()), self._stream()))
            if int(table_obj.stats[1].item()) != 0:
                raise RuntimeError("sparse count policy: a row whose first 64 cells are all below its base value")
        return {"policy": "table", "table": table, "ps_table": ps}

    def organic_count_policy(self):
        return self._policy(0, self.co)

    def bandit_count_policy(self):
        return self._policy(1, self.bandit)

    def zeros_like(self):
        return SparseCountStats(self.P, self.device, *self._want)

    def add_reduced(self, delta, group=None):
        """self += sum over ranks of `delta`: every rank all-gathers the occupied cells of all the deltas (padded to the
        largest) into an un-hashed staging table and adds them through the device insertion path (duplicates of a key
        accumulate); the whole-row increments are all-reduced.  Every rank ends with the same cells."""
        torch = _torch()
        import torch.distributed as dist
        world = dist.get_world_size(group)
        dev = torch.device("cuda", self.device)
        for name, is_bandit in (("co", False), ("bandit", True)):
            if not self._want[1 if is_bandit else 0]:
                continue
            t = getattr(delta, name)
            if t is None:
                t = _DeviceHash(self.P, 1024, self.device, is_bandit)
            used = t.keys != -1
            n_mine = int(used.sum().item())
            sizes = torch.tensor([n_mine], dtype=torch.int64, device=dev)
            all_sizes = [torch.empty_like(sizes) for _ in range(world)]
            dist.all_gather(all_sizes, sizes, group=group)
            n_max = max(int(s.item()) for s in all_sizes)
            total = sum(int(s.item()) for s in all_sizes)
            cols = [t.keys[used], t.a[used]] + ([t.b[used]] if is_bandit else [])
            gathered = []
            for ci, c in enumerate(cols):
                pad = torch.full((n_max,), -1 if ci == 0 else 0, dtype=torch.int64, device=dev)
                pad[:n_mine] = c
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad, group=group)
                gathered.append(torch.cat(parts))
            if is_bandit:
                dist.all_reduce(t.row_a, op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(t.row_b, op=dist.ReduceOp.SUM, group=group)
            target = self._room(getattr(self, name), is_bandit, total)
            if total:
                m = world * n_max
                stage = _DeviceHash(self.P, 1 << max(10, int(m - 1).bit_length()), self.device, is_bandit)
                stage.keys[:m] = gathered[0]
                stage.a[:m] = gathered[1]
                if is_bandit:
                    stage.b[:m] = gathered[2]
                so, sn = stage.struct(), target.struct()
                L.check(L.lib().rg_sparse_rehash_device(C.byref(so), C.byref(sn), self._stream()))
                if int(target.stats[1].item()) != 0:
                    raise RuntimeError("sparse count statistics: hash table overflow while merging the ranks' cells")
            if is_bandit:
                target.row_a.add_(t.row_a)
                target.row_b.add_(t.row_b)
            setattr(self, name, target)

    # ---- host <-> device (agents/_sparse.py HostSparseCounts) --------------------------------------------------------
    def load_host(self, co_host, bandit_host):
        """statistics an agent accumulated on the host (facade training) -> device tables"""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        for host, is_bandit in ((co_host, False), (bandit_host, True)):
            if host is None:
                continue
            keys, a, b = host.arrays()
            # an un-hashed staging table (re-insertion only walks its slots), then the real one
            cap = 1 << max(10, int(max(len(keys), 1) - 1).bit_length())
            stage = _DeviceHash(self.P, cap, self.device, is_bandit)
            n = len(keys)
            stage.keys[:n] = torch.as_tensor(keys.view(np.int64), device=dev)
            stage.a[:n] = torch.as_tensor(a, device=dev)
            if is_bandit:
                stage.b[:n] = torch.as_tensor(b, device=dev)
                for i, v in host.row_a.items():
                    stage.row_a[i] = v
                for i, v in host.row_b.items():
                    stage.row_b[i] = v
            stage.stats[0] = n
            table = _DeviceHash(self.P, 1 << max(10, int(2 * n + 1023).bit_length()), self.device, is_bandit)
            _reinsert(stage, table, self._stream())
            if is_bandit:
                self.bandit = table
            else:
                self.co = table

    def to_host(self):
        """-> (HostSparseCounts of co_counts or None, HostSparseCounts of pulls/clicks or None)"""
        from .agents._sparse import HostSparseCounts
        co = HostSparseCounts.from_cells(self.P, self.co.cells()) if self.co is not None else (
            HostSparseCounts(self.P) if self._want[0] else None)
        ba = None
        if self.bandit is not None:
            ba = HostSparseCounts.from_cells(self.P, self.bandit.cells(), self.bandit.row_a.cpu().numpy(),
                                             self.bandit.row_b.cpu().numpy())
        elif self._want[1]:
            ba = HostSparseCounts(self.P)
        return co, ba


def _reinsert(old, new, stream):
    """growth of a hash table: one kernel adds every occupied cell of `old` to `new` through the device insertion path"""
    so, sn = old.struct(), new.struct()
    L.check(L.lib().rg_sparse_rehash_device(C.byref(so), C.byref(sn), stream))
    if old.row_a is not None:
        new.row_a.copy_(old.row_a)
        new.row_b.copy_(old.row_b)
    if int(new.stats[0].item()) != int(old.stats[0].item()) or int(new.stats[1].item()) != 0:
        raise RuntimeError("sparse count statistics: growth lost cells")


def one_hot_top_k_table(P, k, actions=None):
    """topk[g] = the k indices the reference's recall@k takes as the "top k" of a one-hot 'ps-a' at g:
    numpy.argpartition(one_hot(g), -k)[-k:] (evaluate_agent.py:884) -- g itself plus k - 1 zero-probability products that
    only numpy's partition algorithm determines, so the table is built WITH numpy, one call per distinct action.
    Rows of actions that never occur stay -1.  int32 [P][k]."""
    k = int(min(k, P))
    acts = range(P) if actions is None else sorted(set(int(a) for a in np.asarray(actions).ravel()))
    out = np.full((P, k), -1, np.int32)
    prob = np.zeros(P)
    for g in acts:
        prob[g] = 1.0
        out[g] = np.argpartition(prob, -k)[-k:]
        prob[g] = 0.0
    return out


def recall_table_policy(log, target_table, k=5, u_end=None):
    """evaluate_recall_at_k / verify_agents_recall_at_k (evaluate_agent.py:852-908) of the deterministic policy
    a = target_table[last view] on a DeviceLog: one segmented pass over the log columns.  Returns dict(hits, n, mean, se)
    with se = std(hits) / sqrt(n) as in the reference; the last user of the log is left out, as there."""
    torch = _torch()
    dev = log.env.torch_device
    host_table = target_table.detach().cpu().numpy() if hasattr(target_table, "is_cuda") else np.asarray(target_table)
    if not (hasattr(target_table, "is_cuda") and target_table.is_cuda):
        target_table = torch.as_tensor(np.ascontiguousarray(target_table, np.int32), device=dev)
    P = int(host_table.shape[0])
    topk = one_hot_top_k_table(P, k, host_table)
    topk_d = torch.as_tensor(topk, device=dev)
    if u_end is None:
        u_end = log.user_begin + log.n_users_total - 1
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    cols = log.struct()
    L.check(L.lib().rg_recall_device(C.byref(cols), int(log.rows), int(u_end), C.c_void_p(target_table.data_ptr()),
                                     C.c_void_p(topk_d.data_ptr()), int(topk.shape[1]), C.c_void_p(out.data_ptr()),
                                     log.env._stream()))
    hits, n = (int(x) for x in out.cpu().numpy())
    mean = hits / n if n else float("nan")
    se = (mean * (1.0 - mean)) ** 0.5 / n ** 0.5 if n else float("nan")
    return {"hits": hits, "n": n, "mean": mean, "se": se}


def ips_table_policy(log, target_table, u_end=None):
    """evaluate_IPS / evaluate_SNIPS of the deterministic policy a = target_table[last view] on a DeviceLog
    (evaluate_agent.py:753-835).  Returns dict(ips_mean, ips_se, snips_mean, snips_se, n) with the reference's
    definitions: se = std(ee) / sqrt(n), SNIPS mean = sum(ee) / sum(ratio).  As in the reference the last user of the
    log is left out (`for u in range(max(reco_log.u))`)."""
    torch = _torch()
    dev = log.env.torch_device
    if not (hasattr(target_table, "is_cuda") and target_table.is_cuda):
        target_table = torch.as_tensor(np.ascontiguousarray(target_table, np.int32), device=dev)
    if u_end is None:
        u_end = log.user_begin + log.n_users_total - 1
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    cols = log.struct()
    L.check(L.lib().rg_ips_device(C.byref(cols), int(log.rows), int(u_end), C.c_void_p(target_table.data_ptr()),
                                  C.c_void_p(out.data_ptr()), log.env._stream()))
    s_ee, s_ee2, s_ratio, n = (float(x) for x in out.cpu().numpy())
    mean = s_ee / n
    var = max(s_ee2 / n - mean * mean, 0.0)
    se = var ** 0.5 / n ** 0.5
    return {"ips_mean": mean, "ips_se": se, "snips_mean": s_ee / s_ratio if s_ratio else float("nan"), "snips_se": se,
            "n": int(n), "sum_ee": s_ee, "sum_ratio": s_ratio}
