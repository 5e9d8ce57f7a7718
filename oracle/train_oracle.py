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


def ips_summary(ee, rewards, ratio):
    """verify_agents_IPS / verify_agents_SNIPS (evaluate_agent.py:813-835): mean and standard error."""
    se = np.std(ee) / np.sqrt(len(ee))
    return {"ips_mean": float(np.mean(ee)), "ips_se": float(se),
            "snips_mean": float(np.sum(rewards * ratio) / np.sum(ratio)), "snips_se": float(se)}
