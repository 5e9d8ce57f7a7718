"""CPU: pins oracle/rg_oracle.c against the reference's own outputs (tests/golden, made by
tests/golden/make_golden.py from /root/reference) and against published known-answer vectors."""
import math

import numpy as np
import pytest

from conftest import T2_CASES, T2_CASES_ORACLE_ONLY, load_golden, oracle_env_from_golden


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10: (ctr, key) -> out
    def run(c, k):
        seed = k[0] | (k[1] << 32)
        user = c[0] | (c[1] << 32)
        return [int(x) for x in oracle.philox(seed, user, c[2], c[3])]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert run((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_elementary_functions(oracle):
    from scipy.special import ndtri
    rs = np.random.RandomState(0)
    for x in rs.uniform(-60, 60, 3000):
        assert abs(oracle.exp(x) - math.exp(x)) <= 4e-16 * math.exp(x)
    for x in np.exp(rs.uniform(-700, 40, 3000)):
        assert abs(oracle.log(x) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    for p in np.concatenate([rs.uniform(0, 1, 3000), np.exp(rs.uniform(-700, -1, 2000))]):
        assert abs(oracle.norm_ppf(p) - ndtri(p)) <= 5e-15 * max(1.0, abs(ndtri(p)))
    # ff = sig(5 sig(2 sig(0.3 x) - 2) - 6), reco_env_v1.py:32-41
    sig = lambda v: 1.0 / (1.0 + math.exp(-v))
    for x in rs.uniform(-30, 30, 500):
        assert abs(oracle.ff(x) - sig(5 * sig(2 * sig(0.3 * x) - 2) - 6)) < 1e-16 + 1e-15


def test_uniform_and_normal_streams(oracle):
    w = oracle.philox(42, 7, 3, 0)
    u0 = ((int(w[0]) >> 5) * 67108864.0 + (int(w[1]) >> 6)) / 9007199254740992.0
    u1 = ((int(w[2]) >> 5) * 67108864.0 + (int(w[3]) >> 6)) / 9007199254740992.0
    assert oracle.event_uniform(42, 7, 3, 0) == u0 and oracle.event_uniform(42, 7, 3, 1) == u1
    z = np.concatenate([oracle.normals(1, u, 0, 20) for u in range(5000)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    from scipy import stats
    assert stats.kstest(z, "norm").pvalue > 1e-3
    a = np.array([oracle.policy_uniform_action(1, u, 5, 10) for u in range(20000)])
    assert a.min() == 0 and a.max() == 9
    assert stats.chisquare(np.bincount(a, minlength=10)).pvalue > 1e-3


@pytest.mark.parametrize("case", T2_CASES + T2_CASES_ORACLE_ONLY)
def test_rollout_matches_reference_logic(oracle, case):
    """T2: reference code + injected Philox RNG.  Every discrete log column must be identical."""
    g = load_golden(case)
    env = oracle_env_from_golden(g)
    table = g["table"] if "table" in g.files else None
    eps = float(g["epsilon"]) if "epsilon" in g.files else 0.0
    r = oracle.rollout(env, 0, int(g["n_users"]), policy=str(g["policy"]), table=table,
                       n_organic_users=int(g["n_organic_users"]), log=True, hist=True, epsilon=eps)
    L = r["log"]
    if "normal_time" in g.files:     # float clock: the DataFrame's float32 't' column, rows keyed by the event index
        assert np.array_equal(L["tf"].astype(np.float32), g["log_tf"]), "t (float clock)"
    for k in ("uzvac" if "normal_time" in g.files else "utzvac"):
        assert np.array_equal(L[k], g["log_" + k]), k
    # counters are consistent with the log
    z, c = g["log_z"], g["log_c"]
    n_users_total = int(g["n_users"]) + int(g["n_organic_users"])
    assert r["totals"][0] == n_users_total
    assert r["totals"][4] == int((c == 1).sum())
    assert r["totals"][2] == int((z == 0).sum())
    assert r["totals"][3] + r["totals"][5] == int((z == 1).sum())
    assert np.array_equal(r["view_hist"], np.bincount(g["log_v"][z == 0], minlength=env.P))
    # multi-threaded run gives the same per-user results
    r4 = oracle.rollout(env, 0, int(g["n_users"]), policy=str(g["policy"]), table=table,
                        n_organic_users=int(g["n_organic_users"]), threads=4, epsilon=eps)
    for k in ("n_events", "n_organic", "n_bandit", "n_clicks", "hash", "last_view"):
        assert np.array_equal(r[k], r4[k]), k


def test_batch_step_reproduces_rollout(oracle):
    """gym-style reset/step on a batch == generate_logs episode by episode."""
    g = load_golden("t2_seed42_P10_K5_u100")
    env = oracle_env_from_golden(g)
    n = int(g["n_users"])
    ref = oracle.rollout(env, 0, n, omega=True)
    st, off, sv = oracle.batch_reset(env, np.arange(n))
    n_org = np.diff(off).astype(np.int64)
    n_clicks = np.zeros(n, np.int64)
    for _ in range(2000):
        if (st["state"] == 2).all():
            break
        act = np.array([oracle.policy_uniform_action(int(g["seed"]), u, int(t), env.P)
                        for u, t in zip(st["user_id"], st["t"])], np.int32)
        rew, off, sv = oracle.batch_step(env, st, act)
        n_clicks += rew
        n_org += np.diff(off)
    assert (st["state"] == 2).all()
    assert np.array_equal(st["t"], ref["n_events"])
    assert np.array_equal(n_clicks, ref["n_clicks"])
    assert np.array_equal(n_org, ref["n_organic"])
    assert np.array_equal(st["omega"], ref["omega"])


def test_t1_statistics_are_plausible_under_oracle(oracle):
    """T1 (literal reference, MT19937) can only be compared statistically: events/user law,
    CTR and organic-view histogram of 100 literal users vs 20k oracle users on identical
    static parameters."""
    g = load_golden("t1_seed42_P10_K5")
    from oracle import oracle as O
    env = O.OracleEnv(g["Gamma"], g["mu_organic"], g["beta"], g["mu_bandit"], g["state_transition"],
                      1.0, 0.1, 42)
    r = oracle.rollout(env, 0, 20000, hist=True, threads=4)
    mean_rows = (r["n_events"] + 1).mean()
    ref_rows = g["rows_per_user"]
    assert abs(mean_rows - ref_rows.mean()) < 4 * ref_rows.std() / math.sqrt(len(ref_rows))
    ctr = r["totals"][4] / (r["totals"][3] + r["totals"][5])
    ref_ctr = int(g["n_clicks"]) / int(g["n_bandit"])
    assert abs(ctr - ref_ctr) < 4 * math.sqrt(ctr * (1 - ctr) / int(g["n_bandit"]))
    # organic view histogram: total-variation distance small given 2310 literal views
    p = r["view_hist"] / r["view_hist"].sum()
    q = g["view_hist"] / g["view_hist"].sum()
    assert 0.5 * np.abs(p - q).sum() < 0.08


def test_hole_cdf_table_equals_numpy_choice(oracle):
    """The table form of EpsilonGreedy's exploring draw (groundwork for the device policy) == numpy's legacy
    choice(P, p) algorithm (cumsum, /= cdf[-1], searchsorted right) for every greedy action and many uniforms."""
    rs = np.random.RandomState(3)
    for P in (2, 3, 10, 257, 1000):
        S = oracle.hole_cdf_table(P)
        for g in sorted(set([0, 1, P // 2, P - 2, P - 1]) & set(range(P))):
            p = np.ones(P); p[g] = 0.0; p = p / np.sum(p)
            cdf = p.cumsum(); cdf /= cdf[-1]
            us = np.concatenate([rs.random_sample(300), cdf[:-1], np.nextafter(cdf[:-1], 0), [0.0, 1 - 2.0 ** -53]])
            ref = cdf.searchsorted(us, side="right")
            got = np.array([oracle.hole_choice(S, g, float(u)) for u in us])
            assert np.array_equal(got, ref), (P, g)
            assert (got != g).all()
