"""GPU: device-resident ordered logs, count-agent training statistics, acting tables and IPS/SNIPS sums
(recogym_b200/logs.py, csrc/rg_logs.cu) against the reference goldens and the CPU oracles."""
import numpy as np
import pytest

from conftest import load_golden, oracle_env_from_golden
from test_gpu_parity import MODES, both_envs, device_env_from_golden

pytestmark = pytest.mark.gpu

CASE = "t2_train_seed5_P60_K8_u40_o4"


def host_cols(log):
    h = log.to_host()
    return {k: h[k] for k in "utzvac"}, h


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_device_log_is_the_reference_log_in_reference_order(oracle, mode):
    from recogym_b200.logs import generate_logs_device
    for case in ("t2_seed7_P100_K5_u60_o5", "t2_seed5_P60_K8_u40_bandit_count_flips",
                 "t2_seed13_P40_K6_u60_o3_organic_user_count", CASE):
        g = load_golden(case)
        de = device_env_from_golden(g)
        table = g["table"] if "table" in g.files else None
        # rows_per_user=3 forces the regenerate-into-larger-columns path
        log = generate_logs_device(de, int(g["n_users"]), str(g["policy"]), table, None, 0, int(g["n_organic_users"]),
                                   MODES[mode], rows_per_user=3.0)
        cols, h = host_cols(log)
        for k in "utzvac":
            assert np.array_equal(cols[k].astype(np.int64), g["log_" + k].astype(np.int64)), (case, k)
        ro = log.row_offset.cpu().numpy()
        assert ro[0] == 0 and ro[-1] == log.rows and np.array_equal(np.diff(ro), np.bincount(cols["u"], minlength=len(ro) - 1))
        if str(g["policy"]) == "uniform":
            assert np.array_equal(h["ps"][cols["z"] == 1], np.full(int((cols["z"] == 1).sum()), 1.0 / de.P))
        assert np.isnan(h["ps"][cols["z"] == 0]).all()
        de.close()


def test_derived_columns_equal_oracle_walk(oracle):
    from oracle import train_oracle as T
    from recogym_b200.logs import generate_logs_device
    oe, de = both_envs(300, 8, 5, prob_organic_to_bandit=0.1)       # long organic sessions: > 32 rows per session
    log = generate_logs_device(de, 500, "uniform", None, None, 10, 7, 1, prev_last_view=123)
    cols, h = host_cols(log)
    ro = oracle.rollout(oe, 10, 500, n_organic_users=7, log=True)
    for k in "utzvac":
        assert np.array_equal(cols[k].astype(np.int64), ro["log"][k].astype(np.int64)), k
    assert np.array_equal(h["lv"], T.acting_last_view(cols))
    # lvp / sess against a row-by-row walk of the train calls
    lvp = np.full(len(cols["u"]), -1, np.int64)
    sess = np.zeros(len(cols["u"]), np.int64)
    lpv, run, cur_user = 123, 0, None
    for r in range(len(cols["u"])):
        if cols["u"][r] != cur_user:
            cur_user, run, pending = cols["u"][r], 0, None
        if cols["z"][r] == 0:
            sess[r] = run
            run += 1
            pending = int(cols["v"][r])
        else:
            lvp[r] = lpv
            if pending is not None:
                lpv = pending
            run, pending = 0, None
    assert np.array_equal(h["sess"], sess) and sess.max() > 40
    assert np.array_equal(h["lvp"], lvp)
    de.close()


def test_count_stats_tables_and_ips_equal_reference(oracle):
    from oracle import train_oracle as T
    from recogym_b200.logs import CountStats, generate_logs_device, ips_table_policy
    g = load_golden(CASE)
    de = device_env_from_golden(g)
    P = de.P
    log = generate_logs_device(de, int(g["n_users"]), "uniform", None, None, 0, int(g["n_organic_users"]), 2)
    st = CountStats(de.P)
    st.train(log)
    assert np.array_equal(st.co_counts.cpu().numpy().astype(np.float64), g["co_counts"])
    assert np.array_equal(st.pulls.cpu().numpy().astype(np.float64), g["pulls_a"])
    assert np.array_equal(st.clicks.cpu().numpy().astype(np.float64), g["clicks_a"])
    assert st.last_view == int(g["bandit_lpv"])
    oc, bc = st.organic_count_policy(), st.bandit_count_policy()
    assert np.array_equal(oc["table"].cpu().numpy(), g["co_counts"].argmax(axis=1))
    assert np.array_equal(bc["table"].cpu().numpy(), g["ctr"].argmax(axis=1))
    assert np.array_equal(bc["ps_table"].cpu().numpy(), g["ctr"].max(axis=1))
    # second log continues the statistics exactly as a second training phase would (lpv carried across logs)
    log2 = generate_logs_device(de, 50, "uniform", None, None, 44, 0, 1, prev_last_view=st.last_view)
    st.train(log2)
    both = {k: np.concatenate([log.to_host()[k], log2.to_host()[k]]) for k in "utzvac"}
    co, pulls, clicks, lpv = T.count_stats(P, both, int(g["n_organic_users"]))
    assert np.array_equal(st.co_counts.cpu().numpy(), co) and np.array_equal(st.pulls.cpu().numpy(), pulls)
    assert np.array_equal(st.clicks.cpu().numpy(), clicks) and st.last_view == lpv
    # IPS / SNIPS of the reference-trained OrganicCount on the uniform-policy log (evaluate_agent.py:753-835)
    ilog = generate_logs_device(de, int(g["n_users"]), "uniform", None, None, 0, 0, 2)
    for k in "utzvac":
        assert np.array_equal(ilog.to_host()[k].astype(np.int64), g["ipslog_" + k].astype(np.int64)), k
    r = ips_table_policy(ilog, g["co_counts"].argmax(axis=1).astype(np.int32))
    ee = g["ips_rewards"] * g["ips_p_ratio"]
    assert r["n"] == len(ee) and r["sum_ee"] == ee.sum() and r["sum_ratio"] == g["ips_p_ratio"].sum()
    assert abs(r["ips_mean"] - ee.mean()) < 1e-15 and abs(r["snips_mean"] - g["snips"][1]) < 1e-15
    assert abs(r["ips_se"] - ee.std() / np.sqrt(len(ee))) < 1e-12
    # recall@k of both reference-trained count agents on the same log (evaluate_agent.py:852-908): the device pass ==
    # the reference's own hit lists, and a larger log == the oracle's row-by-row restatement
    from recogym_b200.logs import recall_table_policy
    for tag, table in (("oc", g["co_counts"].argmax(axis=1)), ("bc", g["ctr"].argmax(axis=1))):
        for k in (1, 5):
            rk = recall_table_policy(ilog, table.astype(np.int32), k)
            hits = g["recall_hits_%s_k%d" % (tag, k)]
            assert rk["n"] == len(hits) and rk["hits"] == int(hits.sum()), (tag, k, rk)
            assert abs(rk["mean"] - hits.mean()) < 1e-15 and abs(rk["se"] - hits.std() / np.sqrt(len(hits))) < 1e-12
    big = generate_logs_device(de, 700, "uniform", None, None, 1000, 0, 2)
    tab = np.random.RandomState(3).randint(0, P, P).astype(np.int32)
    want = T.recall_at_k_table_policy(big.to_host(), tab, 3)
    rk = recall_table_policy(big, tab, 3)
    assert rk["n"] == len(want) and rk["hits"] == int(np.sum(want))
    de.close()


def test_trained_on_device_equals_trained_through_the_facade():
    """test_agent's training loop through the per-step gym facade (host agents) == CountStats over one fused log."""
    import recogym_b200 as rg
    from recogym_b200.logs import CountStats, generate_logs_device
    P = 80
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 17, "num_products": P, "K": 6})
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    bc = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    n_org, n = 5, 60
    uid = 0
    for _ in range(n_org):
        env.reset(uid); uid += 1
        obs, _, _, _ = env.step(None)
        oc.train(obs, None, None, True); bc.train(obs, None, None, True)
    for _ in range(n):
        env.reset(uid); uid += 1
        new_obs, _, done, reward = env.step(None)
        while not done:
            old = new_obs
            action, new_obs, reward, done, _ = env.step_offline(old, reward, done)
            oc.train(old, action, reward, False); bc.train(old, action, reward, False)
        action, _, reward, done, _ = env.step_offline(new_obs, reward, done)
        oc.train(new_obs, action, reward, True); bc.train(new_obs, action, reward, True)
    st = CountStats(P)
    st.train(generate_logs_device(env.device_env, n, "uniform", None, None, 0, n_org, 2))
    assert np.array_equal(st.co_counts.cpu().numpy(), oc.co_counts)
    assert np.array_equal(st.pulls.cpu().numpy(), bc.pulls_a) and np.array_equal(st.clicks.cpu().numpy(), bc.clicks_a)
    assert np.array_equal(st.organic_count_policy()["table"].cpu().numpy(), oc.device_policy()["table"])
    pol = st.bandit_count_policy()
    assert np.array_equal(pol["table"].cpu().numpy(), bc.device_policy()["table"])
    assert np.array_equal(pol["ps_table"].cpu().numpy(), bc.device_policy()["ps_table"])


def test_full_size_log_properties():
    """C2 size (P=100, K=5, 1M users): row counts, ordering and column invariants of the device log."""
    import torch
    from recogym_b200 import DeviceEnv
    from recogym_b200.logs import generate_logs_device
    from test_gpu_parity import synthetic_env
    cfg, sp = synthetic_env(100, 5, 42)
    de = DeviceEnv(sp, 1, 0.1, 42)
    N = 1_000_000
    log = generate_logs_device(de, N, "uniform", None, None, 0, 0, 2)
    tot = log.totals.cpu().numpy()
    assert log.rows == tot[1] + tot[5] == tot[2] + tot[3] + N
    n = log.rows
    u, t, z = log.u[:n].long(), log.t[:n], log.z[:n]
    key = u * 4096 + t.long()
    assert bool((key[1:] > key[:-1]).all())                       # sorted by (u, t), no duplicates
    assert int((z == 0).sum()) == tot[2] and int((z == 1).sum()) == tot[3] + N
    assert int(torch.nan_to_num(log.c[:n], nan=0.0).sum()) == tot[4]
    ro = log.row_offset
    assert bool((log.t[:n][ro[:-1]] == 0).all()) and bool((log.z[:n][ro[1:] - 1] == 1).all())   # first row t=0, last row phantom
    de.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_training_modes_equal_oracle_walk(oracle, mode):
    """RG_TRAIN_EVOLUTION_STEP / _INIT: which train calls evaluate_agent's loops make (no phantom call, final session
    skipped, warm-up exception), incl. the carried last_product_viewed across users without a real bandit row and
    across consecutive logs."""
    from oracle import train_oracle as T
    from recogym_b200.logs import CountStats, generate_logs_device
    # prob_leave_organic high: many users stop inside their first session (no real bandit row)
    oe, de = both_envs(40, 5, 31, prob_leave_organic=0.15, prob_organic_to_bandit=0.3)
    st = CountStats(de.P)
    logs = []
    for begin, n in ((0, 300), (300, 200)):
        log = generate_logs_device(de, n, "uniform", None, None, begin, 0, 1, prev_last_view=st.last_view, train_mode=mode)
        st.train(log)
        logs.append({k: log.to_host()[k] for k in "utzvac"})
        fl = log.to_host()["flags"]
        assert ((fl & 4) != 0).sum() > 0                     # users without a real bandit row exist
    both = {k: np.concatenate([l[k] for l in logs]) for k in "utzvac"}
    co, pulls, clicks, lpv = T.count_stats(de.P, both, 0, mode=mode)
    assert np.array_equal(st.co_counts.cpu().numpy(), co)
    assert np.array_equal(st.pulls.cpu().numpy(), pulls) and np.array_equal(st.clicks.cpu().numpy(), clicks)
    assert st.last_view == (-1 if lpv is None else lpv)
    de.close()


def test_device_evaluate_agent_equals_reference_evaluate_agent():
    """recogym_b200.evaluate_agent with every phase on the GPU == the reference's evaluate_agent run on the same (T2)
    streams: per-step successes, failures, per-action click histories, and the warm-up statistics."""
    import recogym_b200 as rg
    g = load_golden("t2_evolution_seed8_P30_K5")
    A = rg.agents
    for tag, ctor, cargs, approach in (("bc_all", A.BanditCount, A.bandit_count_args, rg.TrainingApproach.ALL_DATA),
                                       ("oc_all", A.OrganicCount, A.organic_count_args, rg.TrainingApproach.ALL_DATA),
                                       ("bc_last", A.BanditCount, A.bandit_count_args, rg.TrainingApproach.LAST_STEP)):
        env = rg.make("reco-gym-v1")
        env.init_gym({**rg.env_1_args, "random_seed": 8, "num_products": 30, "K": 5})
        assert np.array_equal(env.Gamma, g["Gamma"])
        agent = ctor(rg.Configuration({**cargs, "num_products": 30}))
        rw = rg.evaluate_agent(env, agent, int(g["n_init"]), int(g["n_step"]), int(g["n_steps"]), approach)
        assert rw[rg.EvolutionCase.SUCCESS] == g[tag + "_success"].tolist(), tag
        assert rw[rg.EvolutionCase.FAILURE] == g[tag + "_failure"].tolist(), tag
        acts = np.array([rw[rg.EvolutionCase.ACTIONS][a] for a in range(30)])
        assert np.array_equal(acts, g[tag + "_actions"]), tag
        agent.sync_host()
        if ctor is A.BanditCount:
            assert np.array_equal(agent.pulls_a, g[tag + "_warm_pulls"]) and np.array_equal(agent.clicks_a, g[tag + "_warm_clicks"])
        else:
            assert np.array_equal(agent.co_counts, g[tag + "_warm_co"])
    # the per-user facade loop (Python agents, one gym step at a time) gives the same numbers, also under a
    # never-exploring EpsilonGreedy whose actions are all 'greedy'
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 8, "num_products": 30, "K": 5})
    mk = lambda: A.EpsilonGreedy(rg.Configuration({**A.epsilon_greedy_args, "num_products": 30, "epsilon": 0.0,
                                                   "random_seed": 1}),
                                 A.BanditCount(rg.Configuration({**A.bandit_count_args, "num_products": 30})))
    dev = rg.evaluate_agent(env, mk(), 12, 15, 2)
    host = rg.evaluate_agent(env, mk(), 12, 15, 2, device=False)
    for k in (rg.EvolutionCase.SUCCESS, rg.EvolutionCase.FAILURE, rg.EvolutionCase.SUCCESS_GREEDY,
              rg.EvolutionCase.FAILURE_GREEDY, rg.EvolutionCase.ACTIONS):
        assert dev[k] == host[k], k


def test_test_agent_device_training_equals_facade_training():
    """test_agent with the training phase on the GPU (fused log + scatter-add kernels) == the same call with the
    per-step facade training loop: identical Beta quantiles (identical successes / failures)."""
    import copy
    import recogym_b200 as rg
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 23, "num_products": 25, "K": 4})
    for ctor, cargs in ((rg.agents.OrganicCount, rg.agents.organic_count_args),
                        (rg.agents.BanditCount, rg.agents.bandit_count_args)):
        agent = ctor(rg.Configuration({**cargs, "num_products": 25}))
        dev = rg.test_agent(copy.deepcopy(env), copy.deepcopy(agent), 40, 3000, 6)
        host = rg.test_agent(copy.deepcopy(env), copy.deepcopy(agent), 40, 3000, 6, device_training=False)
        assert dev == host, ctor.__name__


def _sparse_oracle_cells(T, P, cols, n_org, mode=0, lpv=None):
    co, pulls, clicks, rp, rc, out_lpv = T.sparse_count_stats(P, cols, n_org, last_product_viewed=lpv, mode=mode)
    return co, {k: (pulls[k], clicks.get(k, 0)) for k in pulls}, rp, rc, out_lpv


def test_sparse_count_stats_equal_dense_and_reference(oracle):
    """Device hash-table statistics (SparseCountStats) on the reference golden: cells == the dense matrices the
    reference's own agents hold, acting tables and 'ps' == numpy's argmax over them, growth by re-insertion keeps
    every cell, and a second log continues the statistics."""
    from oracle import train_oracle as T
    from recogym_b200.logs import CountStats, generate_logs_device
    g = load_golden(CASE)
    de = device_env_from_golden(g)
    P = de.P
    log = generate_logs_device(de, int(g["n_users"]), "uniform", None, None, 0, int(g["n_organic_users"]), 2)
    st = CountStats(P, sparse=True)
    assert st.sparse
    st.train(log)
    co = st.co.cells()
    dense_co = np.zeros((P, P)); dense_p = np.zeros((P, P)); dense_c = np.zeros((P, P))
    for (i, j), (a, _) in co.items():
        dense_co[i, j] = a
    ra, rb = st.bandit.row_a.cpu().numpy(), st.bandit.row_b.cpu().numpy()
    for (i, j), (a, b) in st.bandit.cells().items():
        dense_p[i, j] = a; dense_c[i, j] = b
    dense_p += ra[:, None]; dense_c += rb[:, None]         # whole-row increments (last_product_viewed None)
    assert np.array_equal(dense_co, g["co_counts"]) and np.array_equal(dense_p, g["pulls_a"])
    assert np.array_equal(dense_c, g["clicks_a"]) and st.last_view == int(g["bandit_lpv"])
    assert ra.sum() > 0                                    # the golden exercises the None-row quirk
    oc, bc = st.organic_count_policy(), st.bandit_count_policy()
    assert np.array_equal(oc["table"].cpu().numpy(), g["co_counts"].argmax(axis=1))
    assert np.array_equal(bc["table"].cpu().numpy(), g["ctr"].argmax(axis=1))
    assert np.array_equal(bc["ps_table"].cpu().numpy(), g["ctr"].max(axis=1))
    cap0 = st.bandit.capacity
    # many more users: the tables grow (device re-insertion) and rows fill up completely (every cell touched)
    log2 = generate_logs_device(de, 3000, "uniform", None, None, 44, 0, 1, prev_last_view=st.last_view)
    st.train(log2)
    assert st.bandit.capacity > cap0
    dn = CountStats(P, sparse=False)
    dn.train(log); dn.last_view = int(g["bandit_lpv"]); dn.train(log2)
    for a, b in ((st.organic_count_policy(), dn.organic_count_policy()), (st.bandit_count_policy(), dn.bandit_count_policy())):
        assert np.array_equal(a["table"].cpu().numpy(), b["table"].cpu().numpy())
        assert np.array_equal(a["ps_table"].cpu().numpy(), b["ps_table"].cpu().numpy())
    assert st.last_view == dn.last_view
    de.close()


def test_sparse_bandit_count_at_catalogue_scale(oracle):
    """BASELINE configs[3] catalogue size (P = 100 000; dense statistics would be 3 x 80 GB): BanditCount and OrganicCount
    trained on the device == the oracle's dict-of-keys restatement of the reference's train calls on the same log;
    test_agent runs end to end with the package-level agent."""
    import recogym_b200 as rg
    from oracle import train_oracle as T
    from recogym_b200.logs import generate_logs_device
    P, K = 100000, 4
    oe, de = both_envs(P, K, 31)
    n, n_org = 400, 5
    ro = oracle.rollout(oe, 0, n, n_organic_users=n_org, log=True, threads=16)["log"]
    bc = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    oc = rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, "num_products": P}))
    assert bc.pulls_a is None and oc.co_counts is None          # nothing dense was allocated
    log = generate_logs_device(de, n, "uniform", None, None, 0, n_org, 2)
    for k in "utzvac":
        assert np.array_equal(log.to_host()[k].astype(np.int64), ro[k].astype(np.int64)), k
    bc.device_train(log); oc.device_train(log)
    co, cells, rp, rc, lpv = _sparse_oracle_cells(T, P, ro, n_org)
    assert bc._stats.bandit.cells() == cells and {k: v[0] for k, v in oc._stats.co.cells().items()} == co
    assert bc.device_last_view() == lpv
    tb, psb = T.sparse_bandit_count_table(P, {k: v[0] for k, v in cells.items()}, {k: v[1] for k, v in cells.items()}, rp, rc)
    pol = bc.device_policy()
    assert np.array_equal(pol["table"].cpu().numpy(), tb) and np.array_equal(pol["ps_table"].cpu().numpy(), psb)
    assert np.array_equal(oc.device_policy()["table"].cpu().numpy(), T.sparse_organic_count_table(P, co))
    # host synchronisation: the facade-side agent acts like the table
    bc.sync_host(); oc.sync_host()
    t2, ps2 = bc._sp.bandit_table()
    assert np.array_equal(t2, tb) and np.array_equal(ps2, psb) and np.array_equal(oc._sp.organic_table(), T.sparse_organic_count_table(P, co))
    # and back to the device (facade-trained statistics continue on the GPU)
    log2 = generate_logs_device(de, 100, "uniform", None, None, n + n_org, 0, 2, prev_last_view=bc.device_last_view())
    bc.device_train(log2)
    both = {k: np.concatenate([ro[k], log2.to_host()[k]]) for k in "utzvac"}
    _, cells2, rp2, rc2, _ = _sparse_oracle_cells(T, P, both, n_org)
    assert bc._stats.bandit.cells() == cells2
    de.close()
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": 31, "num_products": P, "K": K})
    q = rg.test_agent(env, rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P})),
                      300, 2000, 10)
    assert 0.0 < q[1] < q[0] < q[2] < 0.1
