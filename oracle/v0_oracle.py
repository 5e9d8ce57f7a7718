"""CPU oracle of reco-gym-v0 -- TEST INFRASTRUCTURE ONLY (tests/, never the product path).

A plain Python / NumPy restatement of the reference's toy environment (recogym/envs/reco_env_v0.py:17-69) on the loops
of recogym/envs/abstract.py:90-327, driven by the build's counter RNG (include/recogym_b200.h "reco-gym-v0").  It works
on the REFERENCE-SHAPED parameters (state_transition, product_transition, click_probs, initial_product_probs), not on the
search tables the CUDA library receives, and calls NumPy's own cumsum / searchsorted per draw as RandomState.choice does.
Pinned: tests/golden/t2_v0_*.npz are logs of the unmodified reference with this RNG injected (oracle/ref_harness.py
make_env0); `binomial1` is checked against numpy.random.RandomState.binomial itself (tests/test_env0_cpu.py).
Pure-Python loops: small cases only (a few thousand users).
"""
import math

import numpy as np

from . import oracle as O

ORGANIC, BANDIT, STOP = 0, 1, 2


def binomial1(p, U):
    """RandomState.binomial(1, p) given the uniform U it consumes (numpy legacy distributions.c: legacy_random_binomial ->
    random_binomial_inversion since n p <= 30; for p > 0.5 the draw is made with 1 - p and mirrored)."""
    if p <= 0.5:
        qn = math.exp(math.log(1.0 - p)) if p < 1.0 else 0.0
        return 1 if U > qn else 0
    q = 1.0 - p
    qn = math.exp(math.log(1.0 - q)) if q < 1.0 else 0.0
    return 1 - (1 if U > qn else 0)


def choice(p, x):
    """RandomState.choice(len(p), p=p) given its uniform x (reco_env_v0.py:54,60,68)."""
    cdf = np.asarray(p, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(x, side="right"))


def uniform53(w0, w1):
    return ((int(w0) >> 5) * 67108864.0 + (int(w1) >> 6)) / 9007199254740992.0


def rollout(params, seed, user_begin, n_users, policy="uniform", table=None, n_organic_users=0, actions=None):
    """generate_logs(n_users, agent, n_organic_users) for users user_begin.. (abstract.py:241-327).  policy 'uniform' |
    'table' (a = table[current view]) | 'replay' (ONE user, the given actions; stops where the next one is missing).
    -> dict: log columns u, e, z, x (view | action), c, phantom in the reference's row order; per-user counters; totals."""
    trans, T = params["state_transition"], params["product_transition"]
    cp, init = params["click_probs"], params["initial_product_probs"]
    P = T.shape[0]
    rows = []
    n = n_organic_users + n_users
    per = {k: np.zeros(n, np.int32) for k in ("n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "state")}
    totals = np.zeros(8, np.int64)
    for i in range(n):
        user = user_begin + i
        organic_only = i < n_organic_users
        w = O.philox(seed, user, 0, 3)
        view = choice(init, uniform53(w[0], w[1]))                                    # reset (reco_env_v0.py:50-56)
        state, e, na, waiting = ORGANIC, 0, 0, False
        n_org = n_ban = n_clk = n_ph = 0
        while True:
            while state == ORGANIC:                                                    # abstract.py:105-121
                view = choice(T[view, :], O.event_uniform(seed, user, e, 0))           # update_product_view
                rows.append((i, e, 0, view, 0, 0))
                n_org += 1
                state = choice(trans[ORGANIC, :], O.event_uniform(seed, user, e, 1))   # update_state
                e += 1
            if state == STOP or organic_only:
                break
            if policy == "replay":
                if na >= len(actions):
                    waiting = True
                    break
                a = int(actions[na]); na += 1
            elif policy == "table":
                a = int(table[view])
            else:
                a = O.policy_uniform_action(seed, user, e, P)
            c = binomial1(float(cp[a, view]), O.event_uniform(seed, user, e, 0))       # draw_click
            state = choice(trans[BANDIT, :], O.event_uniform(seed, user, e, 1))
            if c == 1:
                state = ORGANIC                                                        # abstract.py:180-181
            rows.append((i, e, 1, a, c, 0))
            e += 1
            n_ban += 1
            n_clk += c
        if not organic_only and not waiting and policy != "replay":
            a = int(table[view]) if policy == "table" else O.policy_uniform_action(seed, user, e, P)
            rows.append((i, e, 1, a, 0, 1))                                            # the phantom row after done
            n_ph = 1
        per["n_events"][i], per["n_organic"][i], per["n_bandit"][i], per["n_clicks"][i] = n_org + n_ban, n_org, n_ban, n_clk
        per["last_view"][i] = view
        per["state"][i] = BANDIT if waiting else (state if organic_only and state != STOP else STOP)
        totals[:6] += (1, n_org + n_ban, n_org, n_ban, n_clk, n_ph)
    r = np.array(rows, np.int64).reshape(-1, 6)
    out = {"u": r[:, 0].astype(np.int32), "e": r[:, 1].astype(np.int32), "z": r[:, 2].astype(np.int8),
           "x": r[:, 3].astype(np.int32), "c": r[:, 4].astype(np.int8), "phantom": r[:, 5].astype(np.int8), "totals": totals}
    out.update(per)
    return out
