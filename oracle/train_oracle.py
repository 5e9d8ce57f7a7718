"""CPU ORACLE of the log consumers -- TEST INFRASTRUCTURE ONLY.

Plain-Python/numpy restatements, call by call, of what the reference does with a generate_logs() log:
  * test_agent's training loop (recogym/bench_agents.py:169-190) feeding OrganicCount.train
    (recogym/agents/organic_count.py:74-82) and BanditCount.train (recogym/agents/bandit_count.py:49-62);
  * the acting tables OrganicCount.act / BanditCount.act (organic_count.py:54-72, bandit_count.py:28-47);
  * evaluate_IPS / evaluate_SNIPS / verify_agents_IPS / verify_agents_SNIPS (recogym/evaluate_agent.py:753-835).
Pinned against the reference's own agents and estimators by tests/golden/t2_train_*.npz (tests/golden/make_golden.py).
Only tests/ may import this module.
"""
import numpy as np


def iter_train_calls(log, n_organic_users, user_begin=0, mode=0):
    """Yields (user, session_views, action, reward) for every agent.train call of a reference training loop, in order.
    `log`: dict of ordered columns u, z (0 organic / 1 bandit), v, a, c.
    mode 0 = test_agent (bench_agents.py:169-190): organic-only users yield ONE call with action None (:169-173); every
      bandit row -- the phantom row after done included -- is one call whose observation is the organic session right
      before it (:176-190).
    mode 1 = evaluate_agent's step loop (evaluate_agent.py:100-139): `while not done` -- no call for the phantom row,
      the user's final session is never trained on.
    mode 2 = evaluate_agent's warm-up (evaluate_agent.py:56-78): as mode 1, except that a user whose first observation
      is already done gets the phantom call (:65-73)."""
    u, z, v, a, c = (np.asarray(log[k]) for k in ("u", "z", "v", "a", "c"))
    n = len(u)
    i = 0
    while i < n:
        user = int(u[i])
        j = i
        while j < n and u[j] == user:
            j += 1
        if user - user_begin < n_organic_users:
            yield user, [int(x) for x in v[i:j]], None, None
        else:
            session = []
            n_real = int((z[i:j] == 1).sum()) - 1
            for r in range(i, j):
                if z[r] == 0:
                    session.append(int(v[r]))
                else:
                    phantom = r == j - 1
                    if mode == 0 or not phantom or (mode == 2 and n_real == 0):
                        yield user, session, int(a[r]), int(max(c[r], 0))
                    session = []
            assert not session, "a regular user's log ends with the phantom bandit row"
        i = j


def count_stats(P, log, n_organic_users, user_begin=0, last_product_viewed=None, mode=0):
    """-> co_counts, pulls_a, clicks_a (float64 [P][P], as the reference holds them) and BanditCount's final
    last_product_viewed after training on `log`."""
    co = np.zeros((P, P))
    pulls = np.zeros((P, P))
    clicks = np.zeros((P, P))
    lpv = last_product_viewed
    for _, session, action, reward in iter_train_calls(log, n_organic_users, user_begin, mode):
        if session:                                   # OrganicCount.train, organic_count.py:76-82
            b = np.bincount(session, minlength=P).astype(np.float64).reshape(P, 1)
            co = co + np.matmul(b, b.T)
        if action is not None and reward is not None:  # BanditCount.train, bandit_count.py:52-58
            ix, jx = lpv, action
            if session:
                lpv = session[-1]
            pulls[ix, jx] += 1                        # ix None: numpy's [None, jx] is [np.newaxis, jx] -> the whole ROW jx
            clicks[ix, jx] += reward
    return co, pulls, clicks, lpv


def organic_count_table(co):
    return co.argmax(axis=1).astype(np.int32)                    # organic_count.py:56


def bandit_count_table(pulls, clicks):
    ctr = (clicks + 1) / (pulls + 2)                             # bandit_count.py:26,59-61
    table = ctr.argmax(axis=1).astype(np.int32)
    return table, ctr[np.arange(len(table)), table]


def acting_last_view(log):
    """lv per row: the view of the latest organic row at or before it (Agent.update_lpv)."""
    u, z, v = (np.asarray(log[k]) for k in ("u", "z", "v"))
    lv = np.full(len(u), -1, np.int64)
    cur, cur_user = -1, None
    for r in range(len(u)):
        if u[r] != cur_user:
            cur_user, cur = u[r], -1
        if z[r] == 0:
            cur = int(v[r])
        lv[r] = cur
    return lv


def evolution(rollout, P, kind, n_init, n_step, n_steps, approach="all", epsilon=0.0, window=0):
    """evaluate_agent (evaluate_agent.py:51-146) for a count agent (kind "organic" | "bandit"), bare or under
    EpsilonGreedy(epsilon), restated call by call on the oracle's logs.  rollout(user_begin, n, policy, table, epsilon)
    -> ordered log dict (u, z, v, a, c; the phantom row after `done` included).  approach: "all" | "last" | "swall"
    (samples % window == 0) | "expl" (not greedy) | "swexpl".  -> dict(success, failure, success_greedy, failure_greedy,
    actions [P][n_steps + 1])."""
    new = lambda: {"co": np.zeros((P, P)), "pulls": np.zeros((P, P)), "clicks": np.zeros((P, P)), "lpv": None}
    copy = lambda s: {"co": s["co"].copy(), "pulls": s["pulls"].copy(), "clicks": s["clicks"].copy(), "lpv": s["lpv"]}

    def train(st, session, action, reward):       # OrganicCount.train / BanditCount.train on one call
        if session:
            b = np.bincount(session, minlength=P).astype(np.float64).reshape(P, 1)
            st["co"] = st["co"] + np.matmul(b, b.T)
        ix = st["lpv"]
        if session:
            st["lpv"] = session[-1]
        st["pulls"][ix, action] += 1               # ix None -> numpy's [np.newaxis, action]: the whole row `action`
        st["clicks"][ix, action] += reward

    def table_of(st):
        return organic_count_table(st["co"]) if kind == "organic" else bandit_count_table(st["pulls"], st["clicks"])[0]

    agent = new()
    initial = copy(agent)
    if n_init > 0:
        log = rollout(0, n_init, "uniform", None, 0.0)
        for _, session, action, reward in iter_train_calls(log, 0, 0, mode=2):
            train(agent, session, action, reward)
    uid = n_init
    training = copy(agent)
    samples = 0
    out = {k: [] for k in ("success", "failure", "success_greedy", "failure_greedy")}
    actions = [[0] for _ in range(P)]
    for _ in range(n_steps):
        table = np.asarray(table_of(agent), np.int32)
        log = rollout(uid, n_step, "eps_table" if epsilon > 0.0 else "table", table, epsilon)
        u, z, v, a, c = (np.asarray(log[k]) for k in ("u", "z", "v", "a", "c"))
        s = f = sg = fg = 0
        i, n = 0, len(u)
        while i < n:
            j = i
            while j < n and u[j] == u[i]:
                j += 1
            session, last_view = [], None
            for r in range(i, j):
                if z[r] == 0:
                    session.append(int(v[r])); last_view = int(v[r])
                    continue
                if r == j - 1:                       # the phantom row after done: `while not done` never reaches it
                    break
                samples += 1
                greedy = int(a[r]) == int(table[last_view])
                upd = {"all": True, "last": True, "swall": window and samples % window == 0, "expl": not greedy,
                       "swexpl": (not greedy) and window and samples % window == 0}[approach]
                if upd:
                    train(training, session, int(a[r]), int(max(c[r], 0)))
                if c[r] == 1:
                    s += 1; sg += int(greedy and epsilon_wrapped(epsilon)); actions[int(a[r])][-1] += 1
                else:
                    f += 1; fg += int(greedy and epsilon_wrapped(epsilon))
                session = []
            i = j
        uid += n_step
        agent = training
        for p in range(P):
            actions[p].append(0)
        training = copy(initial if approach == "last" else agent)
        out["success"].append(s); out["failure"].append(f); out["success_greedy"].append(sg); out["failure_greedy"].append(fg)
    out["actions"] = np.array(actions, np.int64)
    return out


def epsilon_wrapped(epsilon):
    """'greedy' exists in the action dict only under EpsilonGreedy (evaluate_agent.py:122-129); the goldens with
    epsilon > 0 are the wrapped ones."""
    return epsilon > 0.0


def ips_table_policy(log, ps, target_table):
    """evaluate_IPS + evaluate_SNIPS (evaluate_agent.py:753-811) for pi(a) = [a == target_table[last view]]:
    -> ee list (IPS), rewards list, p_ratio list, over the bandit rows of users 0..max(u)-1."""
    u, z, a, c = (np.asarray(log[k]) for k in ("u", "z", "a", "c"))
    lv = acting_last_view(log)
    umax = int(u.max())
    ee, rewards, ratio = [], [], []
    for r in range(len(u)):
        if z[r] != 1 or u[r] >= umax:
            continue
        pi = 1.0 if int(a[r]) == int(target_table[lv[r]]) else 0.0
        cr = float(max(c[r], 0))
        ee.append(cr * pi / ps[r])
        rewards.append(cr)
        ratio.append(pi / ps[r])
    return np.array(ee), np.array(rewards), np.array(ratio)


def recall_at_k_table_policy(log, target_table, k=5):
    """evaluate_recall_at_k (evaluate_agent.py:852-893) for pi = one-hot at target_table[last view], row by row: an
    un-clicked bandit row of a user < max(u) that is directly followed by an organic row of the same user is a hit when
    the next view is in numpy.argpartition(one_hot, -k)[-k:].  -> list of 0/1."""
    u, z, v, c = (np.asarray(log[k_]) for k_ in ("u", "z", "v", "c"))
    lv = acting_last_view(log)
    umax = int(u.max())
    P = len(target_table)
    hits = []
    for r in range(len(u) - 1):
        if z[r] != 1 or u[r] >= umax:
            continue
        if u[r + 1] != u[r] or z[r + 1] != 0 or c[r] > 0:
            continue
        prob = np.zeros(P)
        prob[int(target_table[lv[r]])] = 1.0
        hits.append(1 if int(v[r + 1]) in set(np.argpartition(prob, -k)[-k:]) else 0)
    return hits


def ips_summary(ee, rewards, ratio):
    """verify_agents_IPS / verify_agents_SNIPS (evaluate_agent.py:813-835): mean and standard error."""
    se = np.std(ee) / np.sqrt(len(ee))
    return {"ips_mean": float(np.mean(ee)), "ips_se": float(se),
            "snips_mean": float(np.sum(rewards * ratio) / np.sum(ratio)), "snips_se": float(se)}


# ---- the same statistics held sparsely (P x P is 80 GB at P = 100 000: BASELINE configs[3]) -----------------------------
def sparse_count_stats(P, log, n_organic_users, user_begin=0, last_product_viewed=None, mode=0):
    """count_stats with dict-of-keys matrices: co[(i, j)], pulls[(i, j)], clicks[(i, j)], plus the whole-row increments
    numpy makes while BanditCount.last_product_viewed is still None (`pulls_a[None, jx] += 1` adds 1 to every cell of
    ROW jx): row_pulls[jx], row_clicks[jx].  -> (co, pulls, clicks, row_pulls, row_clicks, last_product_viewed)."""
    co, pulls, clicks, row_pulls, row_clicks = {}, {}, {}, {}, {}
    lpv = last_product_viewed
    for _, session, action, reward in iter_train_calls(log, n_organic_users, user_begin, mode):
        if session:
            views, cnt = np.unique(np.asarray(session, np.int64), return_counts=True)
            for vi, ci in zip(views.tolist(), cnt.tolist()):
                for vj, cj in zip(views.tolist(), cnt.tolist()):
                    co[(vi, vj)] = co.get((vi, vj), 0) + ci * cj
        if action is not None and reward is not None:
            ix, jx = lpv, action
            if session:
                lpv = session[-1]
            if ix is None:
                row_pulls[jx] = row_pulls.get(jx, 0) + 1
                row_clicks[jx] = row_clicks.get(jx, 0) + reward
            else:
                pulls[(ix, jx)] = pulls.get((ix, jx), 0) + 1
                clicks[(ix, jx)] = clicks.get((ix, jx), 0) + reward
    return co, pulls, clicks, row_pulls, row_clicks, lpv


def sparse_organic_count_table(P, co):
    """co_counts.argmax(axis=1) (organic_count.py:56) from the sparse cells: the largest count of the row, the first
    (lowest) column among ties; a row without cells is all zero -> 0."""
    best = {}
    for (i, j), n in co.items():
        cur = best.get(i)
        if n > 0 and (cur is None or n > cur[0] or (n == cur[0] and j < cur[1])):
            best[i] = (n, j)
    table = np.zeros(P, np.int32)
    for i, (_, j) in best.items():
        table[i] = j
    return table


def sparse_bandit_count_table(P, pulls, clicks, row_pulls=None, row_clicks=None):
    """((clicks + 1) / (pulls + 2)).argmax(axis=1) and that maximum (bandit_count.py:26-47,59-61) from the sparse cells.
    Every cell of row i that was never touched holds the row's base value b_i = (rc_i + 1) / (rp_i + 2) (0.5 without
    whole-row increments); numpy.argmax returns the FIRST column attaining the maximum, so the answer is the best
    touched cell when it beats b_i or when the whole row has been touched (lowest column among ties), else the lowest
    column whose cell is not below b_i."""
    row_pulls, row_clicks = row_pulls or {}, row_clicks or {}
    rows = {}
    for (i, j), n in pulls.items():
        rows.setdefault(i, []).append(j)
    table = np.zeros(P, np.int32)
    ps = np.empty(P, np.float64)
    for i in range(P):
        rp, rc = row_pulls.get(i, 0), row_clicks.get(i, 0)
        ps[i] = (rc + 1) / (rp + 2)
    for i, cols in rows.items():
        rp, rc = row_pulls.get(i, 0), row_clicks.get(i, 0)
        base = (rc + 1) / (rp + 2)
        val = {j: (clicks.get((i, j), 0) + rc + 1) / (pulls[(i, j)] + rp + 2) for j in cols}
        vmax = max(val.values())
        if vmax > base or len(val) == P:       # (a row with every cell touched has no cell left at the base value)
            table[i] = min(j for j, x in val.items() if x == vmax)
            ps[i] = vmax
        else:
            j = 0
            while j in val and val[j] < base:
                j += 1
            table[i] = j
            ps[i] = base
    return table, ps


def sparse_count_table(kind, P, log, n_organic_users, user_begin=0, mode=0):
    """The acting table of OrganicCount / BanditCount after test_agent-style training on `log` (sparse statistics)."""
    co, pulls, clicks, rp, rc, _ = sparse_count_stats(P, log, n_organic_users, user_begin, None, mode)
    if kind == "organic_count":
        return sparse_organic_count_table(P, co)
    if kind == "bandit_count":
        return sparse_bandit_count_table(P, pulls, clicks, rp, rc)[0]
    raise ValueError(kind)
