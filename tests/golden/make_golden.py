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

OUT = os.path.dirname(os.path.abspath(__file__))


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


def t2_case(name, args, n_users, n_organic=0, agent_kind=None, n_train=0):
    env = H.make_env(args, "T2")
    sp = H.static_params(env)
    cfg = env.config
    extra = {}
    agent = None
    if agent_kind is not None:
        ref = H.load_reference()
        from recogym.agents import OrganicCount, organic_count_args, BanditCount, bandit_count_args
        if agent_kind == "organic_user_count":
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
        if agent_kind == "organic_count":
            extra["table"] = agent.co_counts.argmax(axis=1).astype(np.int32)
        elif agent_kind == "bandit_count":
            extra["table"] = agent.ctr.argmax(axis=1).astype(np.int32)
    df = env.generate_logs(n_users, agent=agent, num_organic_offline_users=n_organic)
    A = H.logs_to_arrays(df)
    save(name, seed=int(cfg.random_seed), n_users=n_users, n_organic_users=n_organic,
         sigma_omega_initial=float(cfg.sigma_omega_initial), sigma_omega=float(cfg.sigma_omega),
         change_omega_for_bandits=int(bool(cfg.change_omega_for_bandits)),
         policy=("user_views" if agent_kind == "organic_user_count" else "table" if agent is not None else "uniform"),
         **sp, **{"log_" + k: v for k, v in A.items()}, **extra)
    print(name, "rows", len(df), "clicks", int((A["c"] == 1).sum()))


def main():
    if not H.reference_available():
        raise SystemExit("reference not available at " + H.REFERENCE_ROOT)
    if "--user-views-only" in sys.argv:
        user_views_cases()
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


def user_views_cases():
    t2_case("t2_seed13_P40_K6_u60_o3_organic_user_count",
            {"random_seed": 13, "num_products": 40, "K": 6, "sigma_mu_organic": 1.0}, 60, n_organic=3,
            agent_kind="organic_user_count")
    t2_case("t2_seed21_P500_K10_u30_organic_user_count",
            {"random_seed": 21, "num_products": 500, "K": 10}, 30, agent_kind="organic_user_count")


if __name__ == "__main__":
    main()
