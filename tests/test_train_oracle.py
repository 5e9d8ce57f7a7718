"""CPU: pins oracle/train_oracle.py (the restatement of test_agent's training loop, the count agents' train/act and
the IPS/SNIPS estimators) against the reference's own agents and estimators (tests/golden/t2_train_*.npz)."""
import numpy as np

from conftest import load_golden, oracle_env_from_golden

CASE = "t2_train_seed5_P60_K8_u40_o4"


def golden_log(g, prefix="log_"):
    return {k: g[prefix + k] for k in "utzvac"}


def test_count_stats_match_reference_agents():
    from oracle import train_oracle as T
    g = load_golden(CASE)
    P = g["Gamma"].shape[0]
    co, pulls, clicks, lpv = T.count_stats(P, golden_log(g), int(g["n_organic_users"]))
    assert np.array_equal(co, g["co_counts"])
    assert np.array_equal(pulls, g["pulls_a"]) and np.array_equal(clicks, g["clicks_a"])
    assert lpv == int(g["bandit_lpv"])
    # the whole-row update of the first train call (last_product_viewed is None -> pulls_a[np.newaxis, a]) is in the
    # reference's statistics
    first = np.flatnonzero((g["log_z"] == 1) & (g["log_u"] == int(g["n_organic_users"])))[0]
    assert g["pulls_a"][int(g["log_a"][first])].min() >= 1
    table, ps = T.bandit_count_table(pulls, clicks)
    assert np.array_equal(table, g["ctr"].argmax(axis=1)) and np.array_equal(ps, g["ctr"].max(axis=1))


def test_ips_snips_match_reference_estimators():
    from oracle import train_oracle as T
    g = load_golden(CASE)
    P = g["Gamma"].shape[0]
    log = golden_log(g, "ipslog_")
    ps = np.full(len(log["u"]), 1.0 / P)
    table = T.organic_count_table(g["co_counts"])
    ee, rewards, ratio = T.ips_table_policy(log, ps, table)
    assert np.array_equal(rewards, g["ips_rewards"]) and np.array_equal(ratio, g["ips_p_ratio"])
    s = T.ips_summary(ee, rewards, ratio)
    lo, mid, hi = g["snips"]
    assert s["snips_mean"] == mid and abs((mid - 2 * s["snips_se"]) - lo) < 1e-15 and abs((mid + 2 * s["snips_se"]) - hi) < 1e-15


def test_oracle_rollout_reproduces_training_log(oracle):
    """the C oracle's log for the same users == the reference's generate_logs rows the agents were trained on"""
    g = load_golden(CASE)
    env = oracle_env_from_golden(g)
    r = oracle.rollout(env, 0, int(g["n_users"]), n_organic_users=int(g["n_organic_users"]), log=True)
    for k in "utzvac":
        assert np.array_equal(r["log"][k], g["log_" + k]), k


def test_evolution_warmup_stats_match_reference_evaluate_agent(oracle):
    """mode 2 (evaluate_agent's warm-up loop) of the training oracle on the C oracle's log of the warm-up users == the
    statistics the reference's evaluate_agent leaves in the agent it was given."""
    from oracle import train_oracle as T
    g = load_golden("t2_evolution_seed8_P30_K5")
    env = oracle_env_from_golden(g)
    r = oracle.rollout(env, 0, int(g["n_init"]), log=True)
    co, pulls, clicks, _ = T.count_stats(env.P, r["log"], 0, mode=2)
    assert np.array_equal(pulls, g["bc_all_warm_pulls"]) and np.array_equal(clicks, g["bc_all_warm_clicks"])
    assert np.array_equal(co, g["oc_all_warm_co"])
    co0, pulls0, _, _ = T.count_stats(env.P, r["log"], 0, mode=0)
    assert not np.array_equal(pulls0, pulls) and not np.array_equal(co0, co)      # the modes really differ


def test_host_count_agents_match_reference_agents():
    """recogym_b200's host OrganicCount / BanditCount (train / act, CPU only) fed with the reference's training log,
    call by call as test_agent's loop does, end in the reference agents' statistics and acting tables."""
    import recogym_b200 as rg
    from oracle import train_oracle as T
    from recogym_b200.envs.context import DefaultContext
    from recogym_b200.envs.observation import Observation
    from recogym_b200.envs.session import OrganicSessions
    g = load_golden(CASE)
    P = g["Gamma"].shape[0]
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    bc = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    for user, session, action, reward in T.iter_train_calls(golden_log(g), int(g["n_organic_users"])):
        s = OrganicSessions()
        for v in session:
            s.next(DefaultContext(0, user), v)
        obs = Observation(DefaultContext(0, user), s)
        act = None if action is None else {"t": 0, "u": user, "a": action, "ps": 1.0, "ps-a": ()}
        oc.train(obs, act, reward, action is None)
        bc.train(obs, act, reward, action is None)
    assert np.array_equal(oc.co_counts, g["co_counts"])
    assert np.array_equal(bc.pulls_a, g["pulls_a"]) and np.array_equal(bc.clicks_a, g["clicks_a"])
    assert np.array_equal(bc.ctr, g["ctr"]) and bc.last_product_viewed == int(g["bandit_lpv"])
    assert np.array_equal(oc.device_policy()["table"], g["co_counts"].argmax(axis=1))
    pol = bc.device_policy()
    assert np.array_equal(pol["table"], g["ctr"].argmax(axis=1)) and np.array_equal(pol["ps_table"], g["ctr"].max(axis=1))
    # act() on an observation whose last view is v == the table
    s = OrganicSessions(); s.next(DefaultContext(0, 0), 7)
    a = oc.act(Observation(DefaultContext(1, 0), s), 0, False)
    assert a["a"] == g["co_counts"][7].argmax() and a["ps"] == 1.0


def test_bandit_mf_table_equals_probing_act():
    """BanditMFSquare.device_policy (block-wise arg-max) == probing its unchanged act() once per last view."""
    import torch
    import recogym_b200 as rg
    from recogym_b200.policy import probe_last_view_table
    torch.manual_seed(3)
    ag = rg.agents.BanditMFSquare(rg.Configuration({**rg.agents.bandit_mf_square_args, "num_products": 300}))
    pol = ag.device_policy(rows_per_block=64)
    assert np.array_equal(pol["table"], probe_last_view_table(ag, 300))
    with torch.no_grad():
        ag.last_product_viewed = 5
        assert float(ag.get_logits()[pol["table"][5]]) == float(pol["ps_table"][5])


def test_host_ips_snips_estimators_match_reference():
    """recogym_b200.verify_agents_SNIPS / evaluate_SNIPS / evaluate_IPS on the reference's logged DataFrame (rebuilt from
    the golden columns) == the reference's own numbers, through the vectorised table path AND the row-by-row replay."""
    import pandas as pd
    import recogym_b200 as rg
    g = load_golden(CASE)
    P = g["Gamma"].shape[0]
    z = g["ipslog_z"]
    df = pd.DataFrame({
        "t": g["ipslog_t"].astype(np.float32), "u": pd.array(g["ipslog_u"], dtype=pd.Int32Dtype()),
        "z": np.where(z == 0, "organic", "bandit"),
        "v": pd.array(np.where(z == 0, g["ipslog_v"], 0), dtype=pd.Int32Dtype()),
        "a": pd.array(np.where(z == 1, g["ipslog_a"], 0), dtype=pd.Int32Dtype()),
        "c": np.where(z == 1, g["ipslog_c"], np.nan).astype(np.float32), "ps": np.where(z == 1, 1.0 / P, np.nan)})
    df.loc[z == 1, "v"] = pd.NA
    df.loc[z == 0, "a"] = pd.NA
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    oc.co_counts = g["co_counts"].copy()
    rewards, ratio = rg.evaluate_SNIPS(oc, df)                   # vectorised table path
    assert np.array_equal(rewards, g["ips_rewards"]) and np.array_equal(ratio, g["ips_p_ratio"])
    assert np.array_equal(rg.evaluate_IPS(oc, df), g["ips_rewards"] * g["ips_p_ratio"])
    out = rg.verify_agents_SNIPS(df, {"oc": oc})
    assert list(out.columns) == ["Agent", "0.025", "0.500", "0.975"]
    assert np.allclose([out["0.025"][0], out["0.500"][0], out["0.975"][0]], g["snips"], rtol=0, atol=1e-15)

    class Replay(rg.agents.OrganicCount):                        # no table -> the reference's row-by-row replay
        def device_policy(self):
            return None
    rp = Replay(rg.Configuration({**rg.agents.organic_count_args, "num_products": P, "with_ps_all": True}))
    rp.co_counts = g["co_counts"].copy()
    r2, q2 = rg.evaluate_SNIPS(rp, df[df["u"] < 6])
    m = g["ipslog_u"][g["ipslog_z"] == 1]
    nrow = int(((m < 5)).sum())                                  # users 0..4: the last user of the sub-log is left out
    assert np.array_equal(r2, g["ips_rewards"][:nrow]) and np.array_equal(q2, g["ips_p_ratio"][:nrow])


def _golden_frame(g, prefix="ipslog_"):
    import pandas as pd
    P = g["Gamma"].shape[0]
    z = g[prefix + "z"]
    df = pd.DataFrame({
        "t": g[prefix + "t"].astype(np.float32), "u": pd.array(g[prefix + "u"], dtype=pd.Int32Dtype()),
        "z": np.where(z == 0, "organic", "bandit"),
        "v": pd.array(np.where(z == 0, g[prefix + "v"], 0), dtype=pd.Int32Dtype()),
        "a": pd.array(np.where(z == 1, g[prefix + "a"], 0), dtype=pd.Int32Dtype()),
        "c": np.where(z == 1, g[prefix + "c"], np.nan).astype(np.float32), "ps": np.where(z == 1, 1.0 / P, np.nan)})
    df.loc[z == 1, "v"] = pd.NA
    df.loc[z == 0, "a"] = pd.NA
    return df


def test_recall_at_k_matches_reference():
    """evaluate_recall_at_k / verify_agents_recall_at_k (evaluate_agent.py:852-908): the oracle's row-by-row restatement,
    the package's vectorised table path and its row-by-row replay == the reference's own hit lists for its trained
    OrganicCount and BanditCount (k = 1 and 5), and the summary frame."""
    import recogym_b200 as rg
    from oracle import train_oracle as T
    g = load_golden(CASE)
    P = g["Gamma"].shape[0]
    log = golden_log(g, "ipslog_")
    df = _golden_frame(g)
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    oc.co_counts = g["co_counts"].copy()
    bc = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    bc.pulls_a, bc.clicks_a, bc.ctr = g["pulls_a"].copy(), g["clicks_a"].copy(), g["ctr"].copy()
    bc.last_product_viewed = int(g["bandit_lpv"])
    t_oc = T.organic_count_table(g["co_counts"])
    t_bc, _ = T.bandit_count_table(g["pulls_a"], g["clicks_a"])
    for k in (1, 5):
        assert np.array_equal(T.recall_at_k_table_policy(log, t_oc, k), g["recall_hits_oc_k%d" % k]), k
        assert np.array_equal(T.recall_at_k_table_policy(log, t_bc, k), g["recall_hits_bc_k%d" % k]), k
        assert np.array_equal(rg.evaluate_recall_at_k(oc, df, k=k), g["recall_hits_oc_k%d" % k]), k
        assert np.array_equal(rg.evaluate_recall_at_k(bc, df, k=k), g["recall_hits_bc_k%d" % k]), k
    out = rg.verify_agents_recall_at_k(df, {"oc": oc, "bc": bc}, k=5)
    assert list(out.columns) == ["Agent", "0.025", "0.500", "0.975"] and list(out["Agent"]) == ["oc", "bc"]
    got = np.array([[out["0.025"][i], out["0.500"][i], out["0.975"][i]] for i in range(2)])
    assert np.allclose(got, g["recall_frame_k5"], rtol=0, atol=1e-15)

    class Replay(rg.agents.OrganicCount):                        # no table -> the reference's row-by-row replay
        def device_policy(self):
            return None
    rp = Replay(rg.Configuration({**rg.agents.organic_count_args, "num_products": P, "with_ps_all": True}))
    rp.co_counts = g["co_counts"].copy()
    sub = rg.evaluate_recall_at_k(rp, df[df["u"] < 8], k=5)      # users 0..6 (the last user of the sub-log is left out)
    assert len(sub) > 10 and np.array_equal(sub, g["recall_hits_oc_k5"][:len(sub)])


def test_sparse_count_statistics_equal_dense(oracle):
    from oracle import train_oracle as T
    """oracle/train_oracle.py's dict-of-keys statistics (the checker of the device hash-table statistics used at
    P = 100 000, where P x P does not exist) == the dense restatement that is pinned to the reference's agents:
    acting tables, 'ps' values and the carried last view, for all three training loops, including rows in which every
    cell was touched and the whole-row increments made while last_product_viewed is None."""
    import recogym_b200 as rg
    from recogym_b200.envs.static_params import draw_static_params
    for P, K, seed, n, no in ((12, 3, 5, 300, 4), (60, 8, 7, 300, 0), (7, 2, 1, 500, 3), (200, 4, 9, 80, 2)):
        cfg = rg.Configuration({**rg.env_1_args, "num_products": P, "K": K, "random_seed": seed})
        sp = draw_static_params(cfg)
        oe = oracle.OracleEnv(sp["Gamma"], sp["mu_organic"], sp["beta"], sp["mu_bandit"], sp["state_transition"],
                              1, 0.1, seed)
        log = oracle.rollout(oe, 0, n, n_organic_users=no, log=True)["log"]
        for mode in (0, 1, 2):
            co, pu, cl, lpv = T.count_stats(P, log, no, mode=mode)
            sco, spu, scl, rp, rc, slpv = T.sparse_count_stats(P, log, no, mode=mode)
            assert slpv == lpv
            assert np.array_equal(T.sparse_organic_count_table(P, sco), T.organic_count_table(co))
            tb, psb = T.sparse_bandit_count_table(P, spu, scl, rp, rc)
            tb_d, ps_d = T.bandit_count_table(pu, cl)
            assert np.array_equal(tb, tb_d) and np.array_equal(psb, ps_d)


EVO_EPS = "t2_evolution_eps_seed9_P20_K4"


def test_evolution_under_epsilon_greedy_matches_reference_for_every_training_approach(oracle):
    """oracle/train_oracle.py's call-by-call evolution loop on the C oracle's logs (acting policy = EpsilonGreedy(0.3)
    around the count agent's table, on the policy stream) == the reference's own evaluate_agent for both count agents and
    all five TrainingApproaches: per-step successes / failures, their greedy parts, per-action click histories."""
    from oracle import train_oracle as T
    g = load_golden(EVO_EPS)
    env = oracle_env_from_golden(g)
    P = g["Gamma"].shape[0]
    roll = lambda ub, n, pol, tab, eps: oracle.rollout(env, ub, n, policy=pol, table=tab, log=True, epsilon=eps)["log"]
    for itag, kind in (("bc", "bandit"), ("oc", "organic")):
        for atag in ("all", "last", "swall", "expl", "swexpl"):
            r = T.evolution(roll, P, kind, int(g["n_init"]), int(g["n_step"]), int(g["n_steps"]), atag,
                            float(g["epsilon"]), int(g["window"]))
            tag = itag + "_" + atag
            for k in ("success", "failure", "success_greedy", "failure_greedy"):
                assert list(r[k]) == list(g[tag + "_" + k]), (tag, k, r[k], g[tag + "_" + k])
            assert np.array_equal(r["actions"], g[tag + "_actions"]), tag


def _derived_columns(log, prev_lpv=-1):
    """numpy restatement of the device log's derived columns (csrc/rg_logs.cu columns_kernel) for a log without
    organic-only users: lv, lvp (BanditCount's last_product_viewed before each bandit row when EVERY row trains), sess,
    flags (1 phantom, 2 final session, 4 user without a real bandit row)."""
    u, z, v = (np.asarray(log[k]) for k in ("u", "z", "v"))
    n = len(u)
    lv = np.full(n, -1, np.int32); lvp = np.full(n, -1, np.int32); sess = np.zeros(n, np.int32); flags = np.zeros(n, np.int8)
    lpv = prev_lpv
    i = 0
    while i < n:
        j = i
        while j < n and u[j] == u[i]:
            j += 1
        cur, run, pending = -1, 0, None
        last_real = max([r for r in range(i, j - 1) if z[r] == 1], default=-1)
        for r in range(i, j):
            if z[r] == 0:
                cur = int(v[r]); sess[r] = run; run += 1; pending = cur
                if r > last_real:
                    flags[r] |= 2
            else:
                lvp[r] = lpv
                if pending is not None:
                    lpv = pending
                run, pending = 0, None
                if r == j - 1:
                    flags[r] |= 1
            if last_real < 0:
                flags[r] |= 4
            lv[r] = cur
        i = j
    return lv, lvp, sess, flags


def test_masked_training_log_selects_the_reference_train_calls(oracle):
    """logs.MaskedTrainingLog (the device evolution loop's view of a step log under the selective TrainingApproaches) on
    CPU tensors: the rows it keeps and the last_product_viewed it re-derives give, through the scatter-add kernels'
    semantics (restated in numpy), the statistics of the call-by-call walk that is pinned to the reference."""
    import torch
    from oracle import train_oracle as T
    from recogym_b200.logs import MaskedTrainingLog
    g = load_golden(EVO_EPS)
    env = oracle_env_from_golden(g)
    P = g["Gamma"].shape[0]
    table = np.random.RandomState(1).randint(0, P, P).astype(np.int32)
    log = oracle.rollout(env, 40, 60, policy="eps_table", table=table, log=True, epsilon=0.4)["log"]
    n = len(log["u"])
    lv, lvp, sess, flags = _derived_columns(log, prev_lpv=3)

    class FakeEnv:
        torch_device = torch.device("cpu")

    class FakeLog:
        pass
    fl = FakeLog()
    fl.env, fl.rows, fl.train_mode, fl.capacity = FakeEnv(), n, 1, n
    for name, arr in (("z", log["z"].astype(np.int8)), ("flags", flags), ("a", log["a"].astype(np.int32)), ("lv", lv), ("u", log["u"].astype(np.int32)),
                      ("lvp", lvp), ("sess", sess)):
        setattr(fl, name, torch.as_tensor(arr))
    fl.struct = lambda: None
    for explored_only, window, samples0 in ((True, 0, 0), (False, 5, 13), (True, 3, 7)):
        view = MaskedTrainingLog(fl, table, explored_only, window, samples0, 3)
        z2, lvp2 = view.z.numpy(), view.lvp.numpy()
        # the kernels' semantics (organic_pairs_kernel / bandit_stats_kernel, train_mode 1) on the masked columns
        co = np.zeros((P, P)); pulls = np.zeros((P, P)); clicks = np.zeros((P, P))
        for r in range(n):
            if z2[r] == 0 and not (flags[r] & 2):
                vr = int(log["v"][r]); co[vr, vr] += 1
                for k in range(1, int(sess[r]) + 1):
                    vj = int(log["v"][r - k]); co[vr, vj] += 1; co[vj, vr] += 1
            elif z2[r] == 1 and not (flags[r] & 1):
                ix, ar, rew = int(lvp2[r]), int(log["a"][r]), int(log["c"][r] == 1)
                if ix >= 0:
                    pulls[ix, ar] += 1; clicks[ix, ar] += rew
                else:
                    pulls[ar, :] += 1; clicks[ar, :] += rew
        # the walk: evaluate_agent's step loop with the approach's update rule (oracle/train_oracle.py evolution())
        wco = np.zeros((P, P)); wp = np.zeros((P, P)); wc = np.zeros((P, P)); lpv = 3; samples = samples0; kept = 0
        u, z, v, a, c = (np.asarray(log[k]) for k in ("u", "z", "v", "a", "c"))
        i = 0
        while i < n:
            j = i
            while j < n and u[j] == u[i]:
                j += 1
            session, last_view = [], None
            for r in range(i, j):
                if z[r] == 0:
                    session.append(int(v[r])); last_view = int(v[r]); continue
                if r == j - 1:
                    break
                samples += 1
                greedy = int(a[r]) == int(table[last_view])
                upd = (not explored_only or not greedy) and (not window or samples % window == 0)
                if upd:
                    kept += 1
                    if session:
                        b = np.bincount(session, minlength=P).astype(np.float64).reshape(P, 1); wco += b @ b.T
                    ix = lpv
                    if session:
                        lpv = session[-1]
                    wp[ix, int(a[r])] += 1; wc[ix, int(a[r])] += int(c[r] == 1)
                session = []
            i = j
        assert kept == view.keep_rows and samples == view.n_samples and kept > 3
        assert np.array_equal(co, wco) and np.array_equal(pulls, wp) and np.array_equal(clicks, wc), (explored_only, window)
        assert int(view.carry.item()) == lpv


def test_device_log_dataframe_has_the_reference_schema():
    """DeviceLog.to_dataframe (whole-column construction: masked Int32 arrays, Arrow-decoded `z`) == the frame the
    reference's generate_logs builds row by row (abstract.py:256-327): dtypes, NA pattern, `()` in 'ps-a' of bandit rows."""
    import pandas as pd
    import torch
    from recogym_b200.logs import DeviceLog
    rs = np.random.RandomState(3)
    n = 500
    z = (rs.rand(n) < 0.7).astype(np.int8)
    organic = z == 0
    log = DeviceLog.__new__(DeviceLog)
    log.rows, log.user_begin = n, 0
    cols = {"t": rs.randint(0, 50, n + 7).astype(np.float32), "u": np.sort(rs.randint(0, 40, n + 7)).astype(np.int32),
            "z": np.concatenate([z, np.zeros(7, np.int8)]), "v": rs.randint(0, 9, n + 7).astype(np.int32),
            "a": rs.randint(0, 9, n + 7).astype(np.int32),
            "c": np.concatenate([np.where(organic, np.nan, rs.rand(n) < 0.1), np.zeros(7)]).astype(np.float32),
            "ps": np.concatenate([np.where(organic, np.nan, 1.0 / 9), np.zeros(7)])}
    for k, v in cols.items():
        setattr(log, k, torch.from_numpy(v.copy()))
    df = log.to_dataframe()
    # the reference's construction: python lists with None, then typed (abstract.py:318-326; ids widened to Int32)
    want = pd.DataFrame({
        "t": cols["t"][:n],
        "u": pd.array(cols["u"][:n], dtype=pd.Int32Dtype()),
        "z": ["organic" if o else "bandit" for o in organic],
        "v": pd.array([int(x) if o else None for x, o in zip(cols["v"][:n], organic)], dtype=pd.Int32Dtype()),
        "a": pd.array([None if o else int(x) for x, o in zip(cols["a"][:n], organic)], dtype=pd.Int32Dtype()),
        "c": cols["c"][:n], "ps": cols["ps"][:n],
        "ps-a": [None if o else () for o in organic]})
    assert list(df.columns) == ["t", "u", "z", "v", "a", "c", "ps", "ps-a"] and len(df) == n
    assert (df.dtypes == want.dtypes).all(), (df.dtypes, want.dtypes)
    pd.testing.assert_frame_equal(df, want)
    assert df["v"].isna().sum() == (~organic).sum() and df["a"].isna().sum() == organic.sum()
    assert (df["z"] == "organic").sum() == organic.sum() and df.loc[~organic, "ps-a"].map(lambda x: x == ()).all()


def test_user_views_ps_equals_the_reference_agents_propensity():
    """logs.user_views_ps (segmented running counts over the ordered log columns) == ps of
    OrganicUserEventCounterAgent.act with the default arguments (agents/organic_user_count.py:49-72: action ~
    Categorical(view counts), ps = counts[a] / sum(counts)), recomputed row by row."""
    import torch
    from recogym_b200.logs import user_views_ps
    rs = np.random.RandomState(8)
    P, rows = 7, []
    for user in range(40):
        seen = np.zeros(P)
        for _ in range(rs.randint(1, 6)):
            for _ in range(rs.randint(1, 5)):
                v = rs.randint(P); seen[v] += 1
                rows.append((user + 100, 0, v, -1, np.nan))
            a = rs.choice(P, p=seen / seen.sum())
            rows.append((user + 100, 1, -1, a, seen[a] / seen.sum()))
    u, z, v, a, want = (np.array(c) for c in zip(*rows))
    got = user_views_ps(torch.from_numpy(u.astype(np.int32)), torch.from_numpy(z.astype(np.int8)),
                        torch.from_numpy(v.astype(np.int32)), torch.from_numpy(a.astype(np.int32)), P).numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(want))
    assert user_views_ps(*(torch.zeros(0, dtype=torch.int32),) * 4, P).numel() == 0
