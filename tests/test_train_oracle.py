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
