"""Regenerates tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py

T1 = literal reference (MT19937); T2 = reference logic + injected Philox streams
(oracle/ref_harness.py).  The fixtures hold inputs (static parameters, seeds, policy tables) and
the reference's outputs (generate_logs columns); tests compare oracle/rg_oracle.c and the CUDA
path against them.
"""
import os
import sys
from copy import deepcopy

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

OUT = os.environ.get("RG_GOLDEN_OUT", os.path.dirname(os.path.abspath(__file__)))


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path), "bytes")


def t1_static_and_stats():
    """Literal reference, seed 42, defaults (P=10, K=5): SURVEY.md section 4 items 1-2."""
    env = H.make_env({"random_seed": 42}, "T1")
    sp = H.static_params(env)
    # 'Getting Started.ipynb':80-83 trace: step_offline loop for user 0
    env.reset()
    trace = []
    obs, reward, done, info = env.step(None)
    trace.append(("organic", [int(s["v"]) for s in obs.sessions()], -1, -1))
    while not done:
        action, obs, reward, done, info = env.step_offline(obs, reward, done)
        trace.append(("bandit", [int(s["v"]) for s in obs.sessions()], int(action["a"]), int(reward)))
    gs_views0 = np.array(trace[0][1], np.int32)
    gs_actions = np.array([t[2] for t in trace[1:]], np.int32)
    gs_rewards = np.array([t[3] for t in trace[1:]], np.int32)
    # derived statistics of generate_logs(100) with the default random policy
    env = H.make_env({"random_seed": 42}, "T1")
    df = env.generate_logs(100)
    A = H.logs_to_arrays(df)
    rows_per_user = np.bincount(A["u"], minlength=100)
    view_hist = np.bincount(A["v"][A["z"] == 0], minlength=10)
    action_hist = np.bincount(A["a"][A["z"] == 1], minlength=10)
    save("t1_seed42_P10_K5", seed=42, **sp, gs_views0=gs_views0, gs_actions=gs_actions,
         gs_rewards=gs_rewards, n_rows=len(df), n_organic=int((A["z"] == 0).sum()),
         n_bandit=int((A["z"] == 1).sum()), n_clicks=int((A["c"] == 1).sum()),
         rows_per_user=rows_per_user, view_hist=view_hist, action_hist=action_hist)
    # static parameters for two more shapes (bit-exact parameter parity)
    for seed, P, K, flips in ((7, 100, 5, 0), (11, 50, 8, 10)):
        env = H.make_env({"random_seed": seed, "num_products": P, "K": K, "number_of_flips": flips}, "T1")
        save("t1_static_seed%d_P%d_K%d_F%d" % (seed, P, K, flips), seed=seed, flips=flips,
             **H.static_params(env))


def t2_case(name, args, n_users, n_organic=0, agent_kind=None, n_train=0, normal_time=None):
    env = H.make_env(args, "T2", normal_time=normal_time)
    sp = H.static_params(env)
    cfg = env.config
    extra = {}
    agent = None
    if agent_kind is not None:
        ref = H.load_reference()
        from recogym.agents import OrganicCount, organic_count_args, BanditCount, bandit_count_args
        if agent_kind == "eps_organic_count":
            agent = OrganicCount(ref.Configuration({**organic_count_args, "num_products": cfg.num_products}))  # wrapped below
        elif agent_kind == "organic_user_count":
            from recogym.agents import OrganicUserEventCounterAgent, organic_user_count_args
            agent = OrganicUserEventCounterAgent(ref.Configuration(
                {**organic_user_count_args, "num_products": cfg.num_products, "random_seed": cfg.random_seed}))
            # build the lazily-built model now and key its RNG by the env's (user, time): policy stream
            agent.feature_provider, agent.model = agent.model_builder.build()
            agent.model.rng = H.PhiloxAgentRNG(env, cfg.random_seed)
        elif agent_kind == "organic_count":
            agent = OrganicCount(ref.Configuration({**organic_count_args, "num_products": cfg.num_products}))
        else:
            agent = BanditCount(ref.Configuration({**bandit_count_args, "num_products": cfg.num_products}))
        # train with the reference's own loop (bench_agents.py:176-190) on users 0..n_train-1
        train_env = H.make_env(args, "T2")
        for uid in range(n_train):
            train_env.reset(uid)
            new_obs, _, done, reward = train_env.step(None)
            while not done:
                old_obs = new_obs
                action, new_obs, reward, done, _ = train_env.step_offline(old_obs, reward, done)
                agent.train(old_obs, action, reward, False)
            old_obs = new_obs
            action, _, reward, done, _ = train_env.step_offline(old_obs, reward, done)
            agent.train(old_obs, action, reward, True)
        if agent_kind == "eps_organic_count":
            from recogym.agents import EpsilonGreedy, epsilon_greedy_args
            extra["table"] = agent.co_counts.argmax(axis=1).astype(np.int32)
            extra["epsilon"] = 0.25
            agent = EpsilonGreedy(ref.Configuration({**epsilon_greedy_args, "num_products": cfg.num_products,
                                                     "epsilon": 0.25, "random_seed": cfg.random_seed}), agent)
            agent.rng = H.PhiloxAgentRNG(env, cfg.random_seed)
        elif agent_kind == "organic_count":
            extra["table"] = agent.co_counts.argmax(axis=1).astype(np.int32)
        elif agent_kind == "bandit_count":
            extra["table"] = agent.ctr.argmax(axis=1).astype(np.int32)
    df = env.generate_logs(n_users, agent=agent, num_organic_offline_users=n_organic)
    A = H.logs_to_arrays(df)
    if normal_time is not None:
        extra["normal_time"] = np.array(normal_time, np.float64)
        extra["log_tf"] = df["t"].to_numpy().astype(np.float32)       # the DataFrame's float32 clock column
    save(name, seed=int(cfg.random_seed), n_users=n_users, n_organic_users=n_organic,
         sigma_omega_initial=float(cfg.sigma_omega_initial), sigma_omega=float(cfg.sigma_omega),
         change_omega_for_bandits=int(bool(cfg.change_omega_for_bandits)),
         policy=("user_views" if agent_kind == "organic_user_count" else "eps_table" if agent_kind == "eps_organic_count"
                 else "table" if agent is not None else "uniform"),
         **sp, **{"log_" + k: v for k, v in A.items()}, **extra)
    print(name, "rows", len(df), "clicks", int((A["c"] == 1).sum()))


def v0_cases():
    """reco-gym-v0 (recogym/envs/reco_env_v0.py): T1 static parameters + the literal generate_logs statistics, T2 logs
    (uniform policy with organic-only users; the reference's own BanditCount trained with its own loop and then logged
    as the acting agent; a step() trace under given actions)."""
    ref = H.load_reference()
    # T1: literal reference, seed 42, defaults
    env = H.make_env0({"random_seed": 42}, "T1")
    sp = H.static_params_v0(env)
    df = env.generate_logs(200)
    A = H.logs_to_arrays(df)
    save("t1_v0_seed42_P10", seed=42, num_clusters=2, **sp, n_rows=len(df), n_organic=int((A["z"] == 0).sum()),
         n_bandit=int((A["z"] == 1).sum()), n_clicks=int((A["c"] == 1).sum()),
         view_hist=np.bincount(A["v"][A["z"] == 0], minlength=10), t_values=np.unique(df["t"].to_numpy()))
    for name, args, n_users, n_org, agent_kind in (
            ("t2_v0_seed42_P10_u120_o5", {"random_seed": 42}, 120, 5, None),
            ("t2_v0_seed7_P30_c3_u150", {"random_seed": 7, "num_products": 30, "num_clusters": 3, "phi_var": 4.0,
                                         "prob_leave_organic": 0.03}, 150, 0, None),
            ("t2_v0_seed9_P12_c4_bandit_count", {"random_seed": 9, "num_products": 12, "num_clusters": 4, "phi_var": 40.0},
             150, 2, "bandit_count")):
        env = H.make_env0(args, "T2")
        cfg = env.config
        extra, agent = {}, None
        if agent_kind:
            from recogym.agents import BanditCount, bandit_count_args
            agent = BanditCount(ref.Configuration({**bandit_count_args, "num_products": cfg.num_products}))
            train_env = H.make_env0(args, "T2")
            for uid in range(60):              # the reference's training loop (bench_agents.py:176-190), users 0..59
                train_env.reset(uid)
                new_obs, _, done, reward = train_env.step(None)
                while not done:
                    old_obs = new_obs
                    action, new_obs, reward, done, _ = train_env.step_offline(old_obs, reward, done)
                    agent.train(old_obs, action, reward, False)
                old_obs = new_obs
                action, _, reward, done, _ = train_env.step_offline(old_obs, reward, done)
                agent.train(old_obs, action, reward, True)
            extra["table"] = agent.ctr.argmax(axis=1).astype(np.int32)
            extra["ps_table"] = agent.ctr[np.arange(cfg.num_products), extra["table"]]
            extra["pulls"], extra["clicks"] = agent.pulls_a.copy(), agent.clicks_a.copy()
        df = env.generate_logs(n_users, agent=agent, num_organic_offline_users=n_org)
        A = H.logs_to_arrays(df)
        extra["log_ps"] = df["ps"].to_numpy(dtype="float64", na_value=np.nan)
        # a step() trace of one user under GIVEN actions: views of every observation, rewards, done flags
        tenv = H.make_env0(args, "T2")
        tenv.reset(77)
        obs, reward, done, _ = tenv.step(None)
        views, rewards, dones, acts = [[int(x["v"]) for x in obs.sessions()]], [], [bool(done)], []
        rs = np.random.RandomState(5)
        while not done and len(acts) < 400:
            a = int(rs.randint(cfg.num_products))
            obs, reward, done, _ = tenv.step(a)
            acts.append(a); rewards.append(int(reward)); dones.append(bool(done))
            views.append([int(x["v"]) for x in obs.sessions()])
        save(name, seed=int(cfg.random_seed), n_users=n_users, n_organic_users=n_org, num_clusters=int(cfg.num_clusters),
             policy="table" if agent is not None else "uniform", **H.static_params_v0(env),
             **{"log_" + k: v for k, v in A.items()}, **extra,
             trace_user=77, trace_actions=np.array(acts, np.int32), trace_rewards=np.array(rewards, np.int32),
             trace_done=np.array(dones, np.int8), trace_views=np.concatenate([np.array(v, np.int32) for v in views] or [np.zeros(0, np.int32)]),
             trace_view_counts=np.array([len(v) for v in views], np.int32))
        print(name, "rows", len(df), "clicks", int((A["c"] == 1).sum()), "trace steps", len(acts))


def v0_evolution_case(name="t2_v0_evolution_seed4_P12_c3"):
    """The reference's evaluate_agent (evaluate_agent.py:51-146) and verify_agents (:732-750) on RecoEnv0 with the v0 T2
    streams, for its two count agents."""
    ref = H.load_reference()
    from recogym.agents import OrganicCount, organic_count_args, BanditCount, bandit_count_args
    from recogym import evaluate_agent, verify_agents, TrainingApproach, EvolutionCase
    args = {"random_seed": 4, "num_products": 12, "num_clusters": 3, "phi_var": 9.0}
    n_init, n_step, n_steps = 25, 35, 3
    out = {}
    for tag, ctor, cargs, approach in (("bc_all", BanditCount, bandit_count_args, TrainingApproach.ALL_DATA),
                                       ("oc_last", OrganicCount, organic_count_args, TrainingApproach.LAST_STEP)):
        env = H.make_env0(args, "T2")
        agent = ctor(ref.Configuration({**cargs, "num_products": 12}))
        rw = evaluate_agent(env, agent, n_init, n_step, n_steps, approach)
        out[tag + "_success"] = np.array(rw[EvolutionCase.SUCCESS], np.int64)
        out[tag + "_failure"] = np.array(rw[EvolutionCase.FAILURE], np.int64)
        out[tag + "_actions"] = np.array([rw[EvolutionCase.ACTIONS][a] for a in range(12)], np.int64)
        print(name, tag, "success", rw[EvolutionCase.SUCCESS], "failure", rw[EvolutionCase.FAILURE])
        # A/B of the agent evaluate_agent left trained (its warm-up statistics), 50 users from the env's seed
        env = H.make_env0(args, "T2")
        df = verify_agents(env, 50, {tag: agent})
        out[tag + "_ab"] = df[["0.025", "0.500", "0.975"]].to_numpy(dtype=np.float64)[0]
    env = H.make_env0(args, "T2")
    save(name, seed=4, n_init=n_init, n_step=n_step, n_steps=n_steps, num_clusters=3, phi_var=9.0, **H.static_params_v0(env), **out)


def main():
    if not H.reference_available():
        raise SystemExit("reference not available at " + H.REFERENCE_ROOT)
    if "--v0-only" in sys.argv:
        v0_cases()
        v0_evolution_case()
        return
    if "--v0-evolution-only" in sys.argv:
        v0_evolution_case()
        return
    if "--user-views-only" in sys.argv:
        user_views_cases()
        return
    if "--train-only" in sys.argv:
        train_case()
        return
    if "--evolution-only" in sys.argv:
        evolution_case()
        return
    if "--edge-only" in sys.argv:
        edge_cases()
        return
    if "--evolution-eps-only" in sys.argv:
        evolution_eps_case()
        return
    if "--eps-only" in sys.argv:
        eps_greedy_case()
        return
    if "--time-only" in sys.argv:
        normal_time_case()
        return
    t1_static_and_stats()
    t2_case("t2_seed42_P10_K5_u100", {"random_seed": 42}, 100)
    t2_case("t2_seed7_P100_K5_u60_o5", {"random_seed": 7, "num_products": 100, "K": 5}, 60, n_organic=5)
    t2_case("t2_seed3_P1000_K20_u24_organic_count",
            {"random_seed": 3, "num_products": 1000, "K": 20}, 24, agent_kind="organic_count", n_train=30)
    t2_case("t2_seed5_P60_K8_u40_bandit_count_flips",
            {"random_seed": 5, "num_products": 60, "K": 8, "number_of_flips": 12, "sigma_omega": 0.3},
            40, agent_kind="bandit_count", n_train=60)
    t2_case("t2_seed9_P33_K3_u40_omega_bandits",
            {"random_seed": 9, "num_products": 33, "K": 3, "change_omega_for_bandits": True,
             "prob_organic_to_bandit": 0.4, "prob_bandit_to_organic": 0.2, "prob_leave_organic": 0.03},
            40)
    user_views_cases()
    train_case()
    evolution_case()
    edge_cases()
    eps_greedy_case()
    evolution_eps_case()
    normal_time_case()
    v0_cases()
    v0_evolution_case()


def train_case(name="t2_train_seed5_P60_K8_u40_o4", n_users=40, n_organic=4):
    """The reference's count agents trained by test_agent's own loop (bench_agents.py:169-190) on T2 streams, the log
    the same users produce through generate_logs, and the reference's IPS / SNIPS estimates of the trained
    OrganicCount on that log (evaluate_agent.py:753-835)."""
    ref = H.load_reference()
    from recogym.agents import OrganicCount, organic_count_args, BanditCount, bandit_count_args
    from recogym.evaluate_agent import evaluate_SNIPS, verify_agents_SNIPS
    args = {"random_seed": 5, "num_products": 60, "K": 8, "sigma_omega": 0.3}
    env = H.make_env(args, "T2")
    sp = H.static_params(env)
    cfg = env.config
    P = cfg.num_products
    oc = OrganicCount(ref.Configuration({**organic_count_args, "num_products": P, "with_ps_all": True}))
    bc = BanditCount(ref.Configuration({**bandit_count_args, "num_products": P}))
    uid = 0
    for _ in range(n_organic):
        env.reset(uid); uid += 1
        obs, _, _, _ = env.step(None)
        for ag in (oc, bc):
            ag.train(obs, None, None, True)
    for _ in range(n_users):
        env.reset(uid); uid += 1
        new_obs, _, done, reward = env.step(None)
        while not done:
            old_obs = new_obs
            action, new_obs, reward, done, _ = env.step_offline(old_obs, reward, done)
            for ag in (oc, bc):
                ag.train(old_obs, action, reward, False)
        old_obs = new_obs
        action, _, reward, done, _ = env.step_offline(old_obs, reward, done)
        for ag in (oc, bc):
            ag.train(old_obs, action, reward, True)
    env2 = H.make_env(args, "T2")
    df = env2.generate_logs(n_users, num_organic_offline_users=n_organic)
    A = H.logs_to_arrays(df)
    # off-policy estimates of the trained OrganicCount on a fresh uniform-policy log of the same env
    env3 = H.make_env(args, "T2")
    df3 = env3.generate_logs(n_users)
    # (evaluate_IPS itself no longer runs under numpy 2: `prob_policy != ()` at evaluate_agent.py:776 raises; its
    # per-row terms c * pi(a) / ps are exactly evaluate_SNIPS's rewards * p_ratio, evaluate_agent.py:777,808-809)
    rewards, p_ratio = evaluate_SNIPS(oc, df3)
    snips = verify_agents_SNIPS(df3, {"oc": oc})
    # recall@k of both trained agents on the same log (evaluate_agent.py:852-908); BanditCount needs 'ps-a'
    from recogym.evaluate_agent import evaluate_recall_at_k, verify_agents_recall_at_k
    bc_all = BanditCount(ref.Configuration({**bandit_count_args, "num_products": P, "with_ps_all": True}))
    bc_all.pulls_a, bc_all.clicks_a, bc_all.ctr = bc.pulls_a.copy(), bc.clicks_a.copy(), bc.ctr.copy()
    bc_all.last_product_viewed = bc.last_product_viewed
    recall = {}
    for kk in (1, 5):
        recall["recall_hits_oc_k%d" % kk] = np.asarray(evaluate_recall_at_k(oc, df3, k=kk), np.int64)
        recall["recall_hits_bc_k%d" % kk] = np.asarray(evaluate_recall_at_k(deepcopy(bc_all), df3, k=kk), np.int64)
    rk = verify_agents_recall_at_k(df3, {"oc": oc, "bc": deepcopy(bc_all)}, k=5)
    recall["recall_frame_k5"] = np.array([[rk["0.025"][i], rk["0.500"][i], rk["0.975"][i]] for i in range(2)], np.float64)
    B = H.logs_to_arrays(df3)
    save(name, seed=int(cfg.random_seed), n_users=n_users, n_organic_users=n_organic,
         sigma_omega_initial=float(cfg.sigma_omega_initial), sigma_omega=float(cfg.sigma_omega),
         change_omega_for_bandits=0, policy="uniform", **sp, **{"log_" + k: v for k, v in A.items()},
         **{"ipslog_" + k: v for k, v in B.items()},
         co_counts=oc.co_counts, pulls_a=bc.pulls_a, clicks_a=bc.clicks_a, ctr=bc.ctr,
         bandit_lpv=int(bc.last_product_viewed),
         ips_rewards=np.asarray(rewards, np.float64), ips_p_ratio=np.asarray(p_ratio, np.float64),
         snips=np.array([snips["0.025"][0], snips["0.500"][0], snips["0.975"][0]], np.float64), **recall)
    print(name, "rows", len(df), "co sum", oc.co_counts.sum(), "pulls", bc.pulls_a.sum(), "snips", snips.values.tolist())


def evolution_case(name="t2_evolution_seed8_P30_K5"):
    """The reference's evaluate_agent (evaluate_agent.py:51-146) on T2 streams for its two count agents: per-step
    successes / failures / per-action clicks, and the warm-up statistics left in the agent that was passed in."""
    ref = H.load_reference()
    from recogym.agents import OrganicCount, organic_count_args, BanditCount, bandit_count_args
    from recogym import evaluate_agent, TrainingApproach, EvolutionCase
    args = {"random_seed": 8, "num_products": 30, "K": 5}
    n_init, n_step, n_steps = 25, 40, 3
    out = {}
    sp = cfg = None
    for tag, ctor, cargs, approach in (("bc_all", BanditCount, bandit_count_args, TrainingApproach.ALL_DATA),
                                       ("oc_all", OrganicCount, organic_count_args, TrainingApproach.ALL_DATA),
                                       ("bc_last", BanditCount, bandit_count_args, TrainingApproach.LAST_STEP)):
        env = H.make_env(args, "T2")
        sp, cfg = H.static_params(env), env.config
        agent = ctor(ref.Configuration({**cargs, "num_products": 30}))
        rw = evaluate_agent(env, agent, n_init, n_step, n_steps, approach)
        out[tag + "_success"] = np.array(rw[EvolutionCase.SUCCESS], np.int64)
        out[tag + "_failure"] = np.array(rw[EvolutionCase.FAILURE], np.int64)
        out[tag + "_actions"] = np.array([rw[EvolutionCase.ACTIONS][a] for a in range(30)], np.int64)
        if ctor is BanditCount:
            out[tag + "_warm_pulls"], out[tag + "_warm_clicks"] = agent.pulls_a, agent.clicks_a
        else:
            out[tag + "_warm_co"] = agent.co_counts
        print(name, tag, "success", rw[EvolutionCase.SUCCESS], "failure", rw[EvolutionCase.FAILURE])
    save(name, seed=int(cfg.random_seed), n_init=n_init, n_step=n_step, n_steps=n_steps,
         sigma_omega_initial=float(cfg.sigma_omega_initial), sigma_omega=float(cfg.sigma_omega),
         change_omega_for_bandits=0, **sp, **out)


def evolution_eps_case(name="t2_evolution_eps_seed9_P20_K4"):
    """The reference's evaluate_agent (evaluate_agent.py:51-146) for EpsilonGreedy(epsilon = 0.3) around its two count
    agents under every TrainingApproach, on T2 streams (the wrapper's coin and exploring choice drawn from the policy
    stream at the env's (user, time)): per-step successes / failures, their greedy parts, per-action click histories."""
    ref = H.load_reference()
    from recogym.agents import (OrganicCount, organic_count_args, BanditCount, bandit_count_args, EpsilonGreedy,
                                epsilon_greedy_args)
    from recogym import evaluate_agent, TrainingApproach, EvolutionCase
    P = 20
    args = {"random_seed": 9, "num_products": P, "K": 4}
    n_init, n_step, n_steps, eps, W = 20, 30, 3, 0.3, 7
    out = {}
    sp = cfg = None
    for itag, ctor, cargs in (("bc", BanditCount, bandit_count_args), ("oc", OrganicCount, organic_count_args)):
        for atag, approach in (("all", TrainingApproach.ALL_DATA), ("last", TrainingApproach.LAST_STEP),
                               ("swall", TrainingApproach.SLIDING_WINDOW_ALL_DATA),
                               ("expl", TrainingApproach.ALL_EXPLORATION_DATA),
                               ("swexpl", TrainingApproach.SLIDING_WINDOW_EXPLORATION_DATA)):
            env = H.make_env(args, "T2")
            sp, cfg = H.static_params(env), env.config
            inner = ctor(ref.Configuration({**cargs, "num_products": P}))
            eg = EpsilonGreedy(ref.Configuration({**epsilon_greedy_args, "num_products": P, "epsilon": eps,
                                                  "random_seed": cfg.random_seed}), inner)
            eg.rng = H.PhiloxAgentRNG(env, cfg.random_seed)
            rw = evaluate_agent(env, eg, n_init, n_step, n_steps, approach, sliding_window_samples=W)
            tag = itag + "_" + atag
            out[tag + "_success"] = np.array(rw[EvolutionCase.SUCCESS], np.int64)
            out[tag + "_failure"] = np.array(rw[EvolutionCase.FAILURE], np.int64)
            out[tag + "_success_greedy"] = np.array(rw[EvolutionCase.SUCCESS_GREEDY], np.int64)
            out[tag + "_failure_greedy"] = np.array(rw[EvolutionCase.FAILURE_GREEDY], np.int64)
            out[tag + "_actions"] = np.array([rw[EvolutionCase.ACTIONS][a] for a in range(P)], np.int64)
            print(name, tag, "success", rw[EvolutionCase.SUCCESS], "greedy", rw[EvolutionCase.SUCCESS_GREEDY],
                  "failure", rw[EvolutionCase.FAILURE], "greedy", rw[EvolutionCase.FAILURE_GREEDY])
    save(name, seed=int(cfg.random_seed), n_init=n_init, n_step=n_step, n_steps=n_steps, epsilon=eps, window=W,
         sigma_omega_initial=float(cfg.sigma_omega_initial), sigma_omega=float(cfg.sigma_omega),
         change_omega_for_bandits=0, **sp, **out)


def normal_time_case():
    t2_case("t2_seed51_P30_K4_u50_o4_normal_time",
            {"random_seed": 51, "num_products": 30, "K": 4, "sigma_omega": 0.3}, 50, n_organic=4, normal_time=(0.5, 2.0))


def eps_greedy_case():
    t2_case("t2_seed41_P25_K4_u60_eps_organic_count", {"random_seed": 41, "num_products": 25, "K": 4}, 60,
            agent_kind="eps_organic_count", n_train=40)


def edge_cases():
    """Degenerate shapes and unusual transition matrices through the UNMODIFIED reference (T2 streams)."""
    t2_case("t2_seed31_P1_K2_u80", {"random_seed": 31, "num_products": 1, "K": 2}, 80)
    t2_case("t2_seed32_P7_K1_u80_o6", {"random_seed": 32, "num_products": 7, "K": 1, "sigma_omega": 1.5,
                                        "sigma_omega_initial": 2.0}, 80, n_organic=6)
    t2_case("t2_seed33_P20_K3_u120_fastleave",
            {"random_seed": 33, "num_products": 20, "K": 3, "prob_leave_organic": 0.3, "prob_leave_bandit": 0.2,
             "prob_organic_to_bandit": 0.6, "prob_bandit_to_organic": 0.7, "sigma_mu_organic": 0.5}, 120)


def user_views_cases():
    t2_case("t2_seed13_P40_K6_u60_o3_organic_user_count",
            {"random_seed": 13, "num_products": 40, "K": 6, "sigma_mu_organic": 1.0}, 60, n_organic=3,
            agent_kind="organic_user_count")
    t2_case("t2_seed21_P500_K10_u30_organic_user_count",
            {"random_seed": 21, "num_products": 500, "K": 10}, 30, agent_kind="organic_user_count")


if __name__ == "__main__":
    main()
