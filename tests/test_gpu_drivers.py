"""GPU: the sweep / leaderboard drivers on BASELINE config[0] sizes (reference: recogym/evaluate_agent.py:158-447,
recogym/competition.py:15-96, sim_test.py:59-115), a foreign duck-typed agent through the facade and through the probe
compiler, and the train-then-evaluate hygiene of the drivers (fresh evaluation users, device-trained tables)."""
import copy

import numpy as np
import pytest
from scipy.stats import beta

pytestmark = pytest.mark.gpu

P, K, SEED = 10, 5, 42      # config[0]: reco-env-v1, num_products=10, K=5, ~100 users


def make_env(**over):
    import recogym_b200 as rg
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": SEED, "num_products": P, "K": K, **over})
    return env


def oracle_env_of(env, oracle):
    c = env.config
    return oracle.OracleEnv(env.Gamma, env.mu_organic, env.beta, env.mu_bandit, env.state_transition,
                            c.sigma_omega_initial, c.sigma_omega, int(c.random_seed), c.change_omega_for_bandits)


def quantiles(successes, failures):
    return tuple(beta.ppf(q, successes + 1, failures + 1) for q in (0.025, 0.500, 0.975))


def test_gather_agent_stats_equals_oracle_composition(oracle):
    """gather_agent_stats -> test_agent -> (fused training log + scatter-add statistics, fused evaluation rollout) ==
    the same chain built from the CPU oracle: training log of users [0, n_org + n), count statistics by the call-by-call
    restatement, arg-max table, evaluation rollout of the FRESH users that follow, Beta quantiles."""
    import recogym_b200 as rg
    from oracle import train_oracle as T
    env = make_env()
    oe = oracle_env_of(env, oracle)
    n_org, n_online, samples = 5, 100, (20, 60)
    inits = {**rg.build_agent_init("organic", rg.agents.OrganicCount, {**rg.agents.organic_count_args}),
             **rg.build_agent_init("bandit", rg.agents.BanditCount, {**rg.agents.bandit_count_args})}
    stats = rg.gather_agent_stats(rg.make("reco-gym-v1"), {**rg.env_1_args, "random_seed": SEED, "num_products": P, "K": K},
                                  {}, inits, samples, n_online, 1, False, n_org)
    assert stats[rg.AgentStats.SAMPLES] == samples and set(stats[rg.AgentStats.AGENTS]) == {"organic", "bandit"}
    for name in ("organic", "bandit"):
        got = stats[rg.AgentStats.AGENTS][name]
        for i, n in enumerate(samples):
            log = oracle.rollout(oe, 0, n, n_organic_users=n_org, log=True)["log"]
            co, pulls, clicks, _ = T.count_stats(P, log, n_org)
            table = T.organic_count_table(co) if name == "organic" else T.bandit_count_table(pulls, clicks)[0]
            ev = oracle.rollout(oe, n_org + n, n_online, policy="table", table=np.asarray(table, np.int32))
            tot = ev["totals"]
            want = quantiles(int(tot[4]), int(tot[3] + tot[5] - tot[4]))
            assert (got[rg.AgentStats.Q0_025][i], got[rg.AgentStats.Q0_500][i], got[rg.AgentStats.Q0_975][i]) == want, (name, n)


def test_test_agent_evaluates_fresh_users_and_epochs_continue_the_stream(oracle):
    """The evaluation users are the RNG keys AFTER the training users (the reference env keeps drawing from one stream);
    a second epoch continues after the first one's evaluation users; epoch_with_random_reset re-seeds and starts over."""
    import recogym_b200 as rg
    from oracle import train_oracle as T
    from recogym_b200.bench_agents import _collect_stats
    env = make_env()
    oe = oracle_env_of(env, oracle)
    agent = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": P}))
    n_org, n, n_online = 3, 40, 150
    s = f = 0
    begin = 0
    for epoch in range(2):
        log = oracle.rollout(oe, begin, n, n_organic_users=n_org, log=True)["log"]
        _, pulls, clicks, _ = T.count_stats(P, log, n_org, user_begin=begin)
        table = T.bandit_count_table(pulls, clicks)[0]
        tot = oracle.rollout(oe, begin + n_org + n, n_online, policy="table", table=np.asarray(table, np.int32))["totals"]
        s += int(tot[4]); f += int(tot[3] + tot[5] - tot[4])
        begin += n_org + n + n_online
    q = rg.test_agent(copy.deepcopy(env), copy.deepcopy(agent), n, n_online, n_org, num_epochs=2)
    assert q == (beta.ppf(0.5, s + 1, f + 1), beta.ppf(0.025, s + 1, f + 1), beta.ppf(0.975, s + 1, f + 1))
    r = _collect_stats({"env": copy.deepcopy(env), "agent": copy.deepcopy(agent), "num_offline_users": n,
                        "num_online_users": n_online, "num_organic_offline_users": n_org, "epoch_with_random_reset": True,
                        "epoch": 3, "user_begin": 0})
    assert r["next_user"] == n_online          # re-seeded training and evaluation envs both start at the stream's start
    # the facade training loop (host agent.train calls) sees the same users and gives the same numbers
    q_host = rg.test_agent(copy.deepcopy(env), copy.deepcopy(agent), n, n_online, n_org, num_epochs=2, device_training=False)
    assert q_host == q


def test_generate_logs_accepts_agents_trained_on_the_device():
    """After device_train the count agents hold CUDA tables; generate_logs (fused path) must take them
    (np.ascontiguousarray on a CUDA tensor) and give the rows of the host-trained twin."""
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.logs import generate_logs_device
    env = make_env(num_products=30)
    for ctor, args in ((rg.agents.OrganicCount, rg.agents.organic_count_args), (rg.agents.BanditCount, rg.agents.bandit_count_args)):
        dev_agent = ctor(rg.Configuration({**args, "num_products": 30}))
        log = generate_logs_device(env.device_env, 80, "uniform", None, None, 0, 4, L.MATH_TENSOR,
                                   train_mode=L.TRAIN_TEST_AGENT)
        dev_agent.device_train(log)
        host_agent = ctor(rg.Configuration({**args, "num_products": 30}))
        e = copy.deepcopy(env)
        uid = 0
        for _ in range(4):
            e.reset(uid); uid += 1
            obs, _, _, _ = e.step(None)
            host_agent.train(obs, None, None, True)
        for _ in range(80):
            e.reset(uid); uid += 1
            new_obs, _, done, reward = e.step(None)
            while not done:
                old = new_obs
                action, new_obs, reward, done, _ = e.step_offline(old, reward, done)
                host_agent.train(old, action, reward, False)
            action, _, reward, done, _ = e.step_offline(new_obs, reward, done)
            host_agent.train(new_obs, action, reward, True)
        a = env.generate_logs(25, agent=dev_agent, user_begin=84)
        b = env.generate_logs(25, agent=host_agent, user_begin=84)
        c = env.generate_logs(25, agent=host_agent, user_begin=84, fused=False)
        assert len(a) > 500
        for x, y in ((a, b), (a, c)):
            assert list(x.columns) == list(y.columns) and len(x) == len(y)
            for col in ("t", "u", "v", "a", "c", "ps"):
                xv = x[col].to_numpy(dtype="float64", na_value=np.nan)
                yv = y[col].to_numpy(dtype="float64", na_value=np.nan)
                assert np.array_equal(np.isnan(xv), np.isnan(yv)), col
                assert np.array_equal(np.nan_to_num(xv, nan=-1), np.nan_to_num(yv, nan=-1)), col
            assert (x["z"] == y["z"]).all()
        assert list(a["u"].unique()) == list(range(25))       # labels are 0-based whatever user_begin is


class ForeignAgent:
    """An agent written against the REFERENCE's protocol only (recogym/agents/abstract.py:9-34): nothing from
    recogym_b200.agents.  Recommends the product after the last one viewed; its state is the last view."""

    def __init__(self, config):
        self.config = config
        self.last = 0
        self.trained = 0

    def act(self, observation, reward, done):
        sessions = observation.sessions()
        if sessions:
            self.last = sessions[-1]["v"]
        a = (int(self.last) + 1) % self.config.num_products
        return {"t": observation.context().time(), "u": observation.context().user(), "a": a, "ps": 1.0, "ps-a": ()}

    def train(self, observation, action, reward, done=False):
        self.trained += 1

    def reset(self):
        pass


def test_foreign_duck_typed_agent_through_the_facade_and_the_probe_compiler(oracle):
    """north_star: 'shipped leaderboard agents drop in unchanged'.  A foreign agent (a) drives the per-step facade --
    its log is the oracle's log under the table that says what it does --, (b) compiles through allow_probe to a device
    table, so that verify_agents / test_agent run it as one fused rollout with the same counts as (a)."""
    import recogym_b200 as rg
    env = make_env()
    oe = oracle_env_of(env, oracle)
    cfg = rg.Configuration({"num_products": P})
    table = ((np.arange(P) + 1) % P).astype(np.int32)
    df = env.generate_logs(60, agent=ForeignAgent(cfg))
    ref = oracle.rollout(oe, 0, 60, policy="table", table=table, log=True)
    log, tot = ref["log"], ref["totals"]
    z = (df["z"] == "bandit").to_numpy()
    assert np.array_equal(z.astype(np.int64), log["z"]) and np.array_equal(df["u"].to_numpy(dtype="int64"), log["u"])
    assert np.array_equal(df["a"].to_numpy(dtype="float64", na_value=-1)[z].astype(np.int64), log["a"][z])
    assert np.array_equal(df["v"].to_numpy(dtype="float64", na_value=-1)[~z].astype(np.int64), log["v"][~z])
    assert np.array_equal(df["c"].to_numpy(dtype="float64")[z].astype(np.int64), log["c"][z])
    from recogym_b200.policy import device_policy
    assert device_policy(ForeignAgent(cfg), P) is None
    pol = device_policy(ForeignAgent(cfg), P, allow_probe=True)
    assert np.array_equal(pol["table"], table)
    out = rg.verify_agents(env, 60, {"foreign": ForeignAgent(cfg)}, allow_probe=True)
    s, f = int(tot[4]), int(tot[3] + tot[5] - tot[4])
    assert (out["0.025"][0], out["0.500"][0], out["0.975"][0]) == quantiles(s, f)
    slow = rg.verify_agents(env, 60, {"foreign": ForeignAgent(cfg)})          # not compilable without the probe: facade
    assert (slow["0.025"][0], slow["0.500"][0], slow["0.975"][0]) == quantiles(s, f)
    agent = ForeignAgent(cfg)
    q = rg.test_agent(copy.deepcopy(env), agent, 10, 60, 2, allow_probe=True)    # trains through the facade (train() calls)
    tot2 = oracle.rollout(oe, 12, 60, policy="table", table=table)["totals"]
    s2, f2 = int(tot2[4]), int(tot2[3] + tot2[5] - tot2[4])
    assert q == (beta.ppf(0.5, s2 + 1, f2 + 1), beta.ppf(0.025, s2 + 1, f2 + 1), beta.ppf(0.975, s2 + 1, f2 + 1))


def test_gather_exploration_stats_structure_and_device_equivalence():
    """gather_exploration_stats (evaluate_agent.py:344-447) on config[0] sizes: one evolution per (agent, epsilon); the
    never-exploring point runs on the device and equals a direct evaluate_agent through the facade; exploring points
    keep the books the reference keeps (greedy successes/failures are a subset, every step is counted)."""
    import recogym_b200 as rg
    args = {**rg.env_1_args, "random_seed": SEED, "num_products": P, "K": K}
    inits = rg.build_agent_init("organic", rg.agents.OrganicCount, {**rg.agents.organic_count_args})
    eps = (0.0, 0.2)
    out = rg.gather_exploration_stats(rg.make("reco-gym-v1"), args, {}, inits, rg.TrainingApproach.ALL_DATA,
                                      num_initial_train_users=15, num_step_users=20, epsilons=eps, num_evolution_steps=3)
    assert set(out) == {"organic"} and set(out["organic"]) == {"0.00", "0.20"}
    E = rg.EvolutionCase
    for key in ("0.00", "0.20"):
        r = out["organic"][key]
        assert all(len(r[c]) == 3 for c in (E.SUCCESS, E.SUCCESS_GREEDY, E.FAILURE, E.FAILURE_GREEDY))
        assert all(g <= s for g, s in zip(r[E.SUCCESS_GREEDY], r[E.SUCCESS])) and all(g <= f for g, f in zip(r[E.FAILURE_GREEDY], r[E.FAILURE]))
        assert set(r[E.ACTIONS]) == set(range(P)) and all(len(v) == 4 for v in r[E.ACTIONS].values())
        assert [sum(r[E.ACTIONS][a][i] for a in range(P)) for i in range(3)] == list(r[E.SUCCESS])
    r0 = out["organic"]["0.00"]
    assert r0[E.SUCCESS_GREEDY] == r0[E.SUCCESS] and r0[E.FAILURE_GREEDY] == r0[E.FAILURE]
    env = make_env()
    eg = rg.agents.EpsilonGreedy(rg.Configuration({**rg.agents.epsilon_greedy_args, **args, "epsilon": 0.0}),
                                 rg.agents.OrganicCount(rg.Configuration({**rg.agents.organic_count_args, **args})))
    host = rg.evaluate_agent(copy.deepcopy(env), eg, 15, 20, 3, rg.TrainingApproach.ALL_DATA, device=False)
    assert host[E.SUCCESS] == r0[E.SUCCESS] and host[E.FAILURE] == r0[E.FAILURE] and host[E.ACTIONS] == r0[E.ACTIONS]
    r2 = out["organic"]["0.20"]
    assert sum(r2[E.SUCCESS_GREEDY]) + sum(r2[E.FAILURE_GREEDY]) < sum(r2[E.SUCCESS]) + sum(r2[E.FAILURE])   # it did explore


def test_competition_score_runs_the_leaderboard_protocol():
    """competition_score (competition.py:15-96): logs of the OrganicUserEventCounterAgent logging policy, training of the
    entrant, online scoring with a re-seeded env; one-row frame.  Equal to the test_agent call it stands for."""
    import recogym_b200 as rg
    df = rg.competition_score(P, 5, 40, 120, SEED, K, 0, 0.0, 0.1, rg.agents.BanditCount, {**rg.agents.bandit_count_args},
                              "bandit-count")
    assert list(df.columns) == ["q0.025", "q0.500", "q0.975", "time"] and len(df) == 1
    assert 0.0 < df["q0.025"][0] < df["q0.500"][0] < df["q0.975"][0] < 0.2
    args = {**rg.env_1_args, "random_seed": SEED, "num_products": P, "K": K, "sigma_omega": 0.1, "number_of_flips": 0}
    logger = rg.agents.OrganicUserEventCounterAgent(rg.Configuration({
        **rg.agents.organic_user_count_args, **args, "select_randomly": True, "epsilon": 0.0, "num_products": P}))
    env = rg.make("reco-gym-v1")
    env.init_gym({**args, "agent": logger})
    agent = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, **args, "agent": logger, "num_products": P}))
    q = rg.test_agent(env, agent, 40, 120, 5, 1, True)
    assert (df["q0.500"][0], df["q0.025"][0], df["q0.975"][0]) == q


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_device_epsilon_greedy_policy_equals_reference_and_oracle(oracle, mode):
    """RG_POLICY_EPS_TABLE (EpsilonGreedy(epsilon) around a table agent, agents/epsilon_greedy.py:30-68): the fused rollout
    == the reference's own EpsilonGreedy(OrganicCount) on the T2 policy stream (golden t2_seed41: event log in reference
    order) and == the C oracle per user at other shapes and epsilons."""
    from conftest import load_golden, oracle_env_from_golden
    from test_gpu_parity import MODES, assert_rollout_equal, both_envs, device_env_from_golden
    from recogym_b200.envs.device_env import rollout_log
    g = load_golden("t2_seed41_P25_K4_u60_eps_organic_count")
    oe, de = oracle_env_from_golden(g), device_env_from_golden(g)
    n, eps, table = int(g["n_users"]), float(g["epsilon"]), g["table"].astype(np.int32)
    rc = oracle.rollout(oe, 0, n, policy="eps_table", table=table, hist=True, omega=True, epsilon=eps)
    rg_ = de.rollout_host(n, "eps_table", table, 0, 0, MODES[mode], per_user=True, omega=True, hist=True, epsilon=eps)
    assert_rollout_equal(rg_, rc)
    log = rollout_log(de, n, "eps_table", table, 0, 0, MODES[mode], epsilon=eps)
    for k in "utzvac":
        assert np.array_equal(log[k].astype(np.int64), g["log_" + k].astype(np.int64)), k
    de.close()
    for P, K, e in ((2, 3, 0.5), (300, 8, 0.05), (1025, 20, 0.9)):
        oe, de = both_envs(P, K, 900 + P)
        table = np.random.RandomState(P).randint(0, P, P).astype(np.int32)
        rc = oracle.rollout(oe, 7, 300, policy="eps_table", table=table, n_organic_users=2, hist=True, omega=True, epsilon=e, threads=4)
        rg_ = de.rollout_host(300, "eps_table", table, 7, 2, MODES[mode], per_user=True, omega=True, hist=True, epsilon=e)
        assert_rollout_equal(rg_, rc)
        de.close()


def test_epsilon_greedy_agent_compiles_to_the_device_policy_and_logs_its_propensities():
    """EpsilonGreedy(epsilon > 0) around a count agent: device_policy() is RG_POLICY_EPS_TABLE; generate_logs (fused) gives
    'ps' = (1 - eps) * greedy ps for greedy actions and eps / (P - 1) for exploring ones (epsilon_greedy.py:44-68), the
    exploring fraction is epsilon, and the device log's ps column agrees."""
    import recogym_b200 as rg
    from recogym_b200 import _lib as L
    from recogym_b200.logs import generate_logs_device
    Pn, eps = 40, 0.2
    env = make_env(num_products=Pn)
    inner = rg.agents.BanditCount(rg.Configuration({**rg.agents.bandit_count_args, "num_products": Pn}))
    log0 = generate_logs_device(env.device_env, 300, "uniform", None, None, 0, 0, L.MATH_TENSOR, train_mode=L.TRAIN_TEST_AGENT)
    inner.device_train(log0)
    eg = rg.agents.EpsilonGreedy(rg.Configuration({**rg.agents.epsilon_greedy_args, "num_products": Pn, "epsilon": eps}), inner)
    pol = eg.device_policy()
    assert pol["policy"] == "eps_table" and pol["epsilon"] == eps
    table = pol["table"].cpu().numpy() if hasattr(pol["table"], "cpu") else np.asarray(pol["table"])
    ps_table = pol["ps_table"].cpu().numpy() if hasattr(pol["ps_table"], "cpu") else np.asarray(pol["ps_table"])
    df = env.generate_logs(400, agent=eg, user_begin=300)
    z = (df["z"] == "bandit").to_numpy()
    v = df["v"].to_numpy(dtype="float64", na_value=np.nan)
    import pandas as pd
    lv = pd.Series(np.where(z, np.nan, v)).groupby(df["u"].to_numpy(dtype="int64")).ffill().to_numpy()[z].astype(np.int64)
    a = df["a"].to_numpy(dtype="float64", na_value=-1)[z].astype(np.int64)
    ps = df["ps"].to_numpy(dtype="float64")[z]
    greedy = a == table[lv]
    assert np.array_equal(ps[greedy], (1.0 - eps) * ps_table[lv[greedy]])
    assert np.array_equal(ps[~greedy], np.full(int((~greedy).sum()), eps * (1.0 / (Pn - 1))))
    frac = 1.0 - greedy.mean()
    assert abs(frac - eps) < 4 * np.sqrt(eps * (1 - eps) / len(a)), frac
    dlog = generate_logs_device(env.device_env, 400, "eps_table", pol["table"], pol["ps_table"], 300, 0, L.MATH_TENSOR, epsilon=eps)
    h = dlog.to_host()
    assert np.array_equal(h["ps"][h["z"] == 1], ps) and np.array_equal(h["a"][h["z"] == 1], a)
    # the A/B driver runs it as one fused rollout
    out = rg.verify_agents(env, 300, {"eg": eg, "inner": inner}, user_begin=700)
    assert len(out) == 2 and (out["0.025"] < out["0.975"]).all()


def test_device_evolution_under_epsilon_greedy_equals_reference_for_every_training_approach():
    """evaluate_agent on the GPU (fused rollouts under RG_POLICY_EPS_TABLE, ordered device logs, scatter-add training of
    the calls each TrainingApproach selects, arg-max tables) == the reference's own evaluate_agent for EpsilonGreedy(0.3)
    around both count agents under all five approaches (golden t2_evolution_eps: per-step successes / failures, their
    greedy parts, per-action click histories)."""
    import recogym_b200 as rg
    from conftest import load_golden
    g = load_golden("t2_evolution_eps_seed9_P20_K4")
    Pn = g["Gamma"].shape[0]
    env = rg.make("reco-gym-v1")
    env.init_gym({**rg.env_1_args, "random_seed": int(g["seed"]), "num_products": Pn, "K": g["Gamma"].shape[1]})
    assert np.array_equal(env.Gamma, g["Gamma"]) and np.array_equal(np.ravel(env.mu_organic), np.ravel(g["mu_organic"]))
    A, T, E = rg.agents, rg.TrainingApproach, rg.EvolutionCase
    for itag, ctor, cargs in (("bc", A.BanditCount, A.bandit_count_args), ("oc", A.OrganicCount, A.organic_count_args)):
        for atag, approach in (("all", T.ALL_DATA), ("last", T.LAST_STEP), ("swall", T.SLIDING_WINDOW_ALL_DATA),
                               ("expl", T.ALL_EXPLORATION_DATA), ("swexpl", T.SLIDING_WINDOW_EXPLORATION_DATA)):
            eg = A.EpsilonGreedy(rg.Configuration({**A.epsilon_greedy_args, "num_products": Pn, "epsilon": float(g["epsilon"])}),
                                 ctor(rg.Configuration({**cargs, "num_products": Pn})))
            rw = rg.evaluate_agent(copy.deepcopy(env), eg, int(g["n_init"]), int(g["n_step"]), int(g["n_steps"]), approach,
                                   sliding_window_samples=int(g["window"]))
            tag = itag + "_" + atag
            assert rw[E.SUCCESS] == list(g[tag + "_success"]) and rw[E.FAILURE] == list(g[tag + "_failure"]), tag
            assert rw[E.SUCCESS_GREEDY] == list(g[tag + "_success_greedy"]), tag
            assert rw[E.FAILURE_GREEDY] == list(g[tag + "_failure_greedy"]), tag
            assert np.array_equal(np.array([rw[E.ACTIONS][a] for a in range(Pn)]), g[tag + "_actions"]), tag
    # a bare count agent has no 'greedy' key: the selective approaches fail as in the reference, SLIDING_WINDOW_ALL_DATA runs
    bare = A.BanditCount(rg.Configuration({**A.bandit_count_args, "num_products": Pn}))
    rw = rg.evaluate_agent(copy.deepcopy(env), bare, 10, 20, 2, T.SLIDING_WINDOW_ALL_DATA, sliding_window_samples=5)
    assert len(rw[E.SUCCESS]) == 2 and rw[E.SUCCESS_GREEDY] == [0, 0]
    with pytest.raises(KeyError):
        rg.evaluate_agent(copy.deepcopy(env), bare, 10, 20, 1, T.ALL_EXPLORATION_DATA, device=True)


@pytest.mark.parametrize("mode", ["exact", "fast", "tensor"])
def test_normal_time_generator_on_the_device_equals_reference_and_oracle(oracle, mode):
    """NormalTimeGenerator dynamics (normal_time_generator.py:7-31; reco_env_v1.py:89-99: the omega drift is scaled by the
    clock step): the fused rollout == the oracle per user (final omega bit-exact) and == the reference's own generator on
    T2 streams (golden t2_seed51: every log column incl. the float32 clock), in all three math modes."""
    from conftest import load_golden, oracle_env_from_golden
    from test_gpu_parity import MODES, assert_rollout_equal, synthetic_env
    from recogym_b200 import DeviceEnv
    from recogym_b200.logs import generate_logs_device
    g = load_golden("t2_seed51_P30_K4_u50_o4_normal_time")
    nt = tuple(float(x) for x in g["normal_time"])
    oe = oracle_env_from_golden(g)
    sp = {k: g[k] for k in ("Gamma", "mu_organic", "beta", "mu_bandit", "state_transition")}
    de = DeviceEnv(sp, float(g["sigma_omega_initial"]), float(g["sigma_omega"]), int(g["seed"]), False, 0, nt)
    n, no = int(g["n_users"]), int(g["n_organic_users"])
    rc = oracle.rollout(oe, 0, n, n_organic_users=no, hist=True, omega=True)
    rg_ = de.rollout_host(n, "uniform", None, 0, no, MODES[mode], per_user=True, omega=True, hist=True)
    assert_rollout_equal(rg_, rc)
    log = generate_logs_device(de, n, "uniform", None, None, 0, no, MODES[mode])
    h = log.to_host()
    for k in "uzvac":
        assert np.array_equal(h[k].astype(np.int64), g["log_" + k].astype(np.int64)), k
    assert np.array_equal(log.t[:log.rows].cpu().numpy(), g["log_tf"]), "float32 clock column"
    de.close()
    cfg, sp2 = synthetic_env(500, 12, 77, sigma_omega=0.3)
    from oracle import oracle as O
    oe2 = O.OracleEnv(sp2["Gamma"], sp2["mu_organic"], sp2["beta"], sp2["mu_bandit"], sp2["state_transition"],
                      cfg.sigma_omega_initial, cfg.sigma_omega, 77, normal_time=(1.0, 0.7))
    de2 = DeviceEnv(sp2, cfg.sigma_omega_initial, cfg.sigma_omega, 77, False, 0, (1.0, 0.7))
    rc = oracle.rollout(oe2, 11, 400, hist=True, omega=True, threads=4)
    rg_ = de2.rollout_host(400, "uniform", None, 11, 0, MODES[mode], per_user=True, omega=True, hist=True)
    assert_rollout_equal(rg_, rc)
    de2.close()


def test_facade_with_normal_time_generator_reproduces_the_reference_log():
    """env.init_gym(time_generator=NormalTimeGenerator(...)): generate_logs through the per-step facade and through the
    fused path give the reference's own rows (golden t2_seed51), float32 clock column included."""
    import recogym_b200 as rg
    from conftest import load_golden
    g = load_golden("t2_seed51_P30_K4_u50_o4_normal_time")
    mu, sigma = (float(x) for x in g["normal_time"])
    args = {**rg.env_1_args, "random_seed": int(g["seed"]), "num_products": g["Gamma"].shape[0], "K": g["Gamma"].shape[1],
            "sigma_omega": float(g["sigma_omega"]), "normal_time_mu": mu, "normal_time_sigma": sigma}
    args["time_generator"] = rg.NormalTimeGenerator(rg.Configuration(args))
    env = rg.make("reco-gym-v1")
    env.init_gym(args)
    assert np.array_equal(env.Gamma, g["Gamma"])
    n, no = int(g["n_users"]), int(g["n_organic_users"])
    for fused in (True, False):
        df = env.generate_logs(n, num_organic_offline_users=no, fused=fused)
        z = (df["z"] == "bandit").to_numpy()
        assert np.array_equal(z.astype(np.int64), g["log_z"].astype(np.int64)), fused
        assert np.array_equal(df["t"].to_numpy(dtype="float32"), g["log_tf"]), fused
        assert np.array_equal(df["v"].to_numpy(dtype="float64", na_value=-1)[~z].astype(np.int64), g["log_v"][~z]), fused
        assert np.array_equal(df["a"].to_numpy(dtype="float64", na_value=-1)[z].astype(np.int64), g["log_a"][z]), fused
        assert np.array_equal(df["c"].to_numpy(dtype="float64")[z].astype(np.int64), g["log_c"][z]), fused
