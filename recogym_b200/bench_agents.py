"""test_agent: train an agent on logged users, evaluate it online, return Beta-posterior CTR
quantiles (reference: recogym/bench_agents.py:66-251; the pickle cache :17-63 is out of scope).

Training runs through the RecoEnv1 facade because `agent.train` is arbitrary Python (count agents: one fused
rollout + scatter-add kernels); the online evaluation -- the expensive part -- runs as ONE fused CUDA rollout
whenever the trained agent compiles to a device policy (recogym_b200.policy), else through the facade's
generate_logs.  The evaluation users are FRESH (RNG keys after the training users'), as in the reference, whose env
keeps drawing from one sequential stream.
"""
from copy import deepcopy

import numpy as np
from scipy.stats.distributions import beta

from .constants import AgentStats
from .policy import device_policy


def _collect_stats(args):
    env, agent = args["env"], args["agent"]
    epoch = args["epoch"]
    new_agent = deepcopy(agent)
    if args["epoch_with_random_reset"]:
        train_env = deepcopy(env)
        train_env.reset_random_seed(epoch)
    else:
        train_env = env
    # First RNG key of this epoch's training users.  The reference env draws from one sequential stream: without
    # epoch_with_random_reset every epoch -- and the evaluation after the training -- sees FRESH users; with it, training
    # and evaluation envs are both re-seeded with seed + epoch and replay the stream from its start (bench_agents.py:
    # 79-86,198-199).  Here a user is a counter-RNG key, so "fresh" means "keys not used before".
    uid = int(args.get("user_begin", 0))
    train_begin = uid
    log_pol = device_policy(train_env.agent, train_env.config.num_products) if train_env.agent else {"policy": "uniform"}
    if args.get("device_training", True) and hasattr(new_agent, "device_train") and log_pol is not None \
            and not getattr(new_agent.config, "with_ps_all", False) and getattr(train_env, "supports_device_training", True):
        # count agents: the whole training phase is ONE fused rollout of the logging policy + scatter-add kernels over
        # its ordered device log (recogym_b200/logs.py) -- the same train calls as the loops below
        from . import _lib as L
        from .logs import generate_logs_device
        prev = new_agent.device_last_view() if hasattr(new_agent, "device_last_view") else -1
        log = generate_logs_device(train_env.device_env, args["num_offline_users"], log_pol["policy"],
                                   log_pol.get("table"), log_pol.get("ps_table"), train_begin,
                                   args["num_organic_offline_users"], L.MATH_TENSOR, prev_last_view=prev,
                                   train_mode=L.TRAIN_TEST_AGENT, epsilon=log_pol.get("epsilon", 0.0))
        new_agent.device_train(log)
        uid += args["num_organic_offline_users"] + args["num_offline_users"]
        args = {**args, "num_organic_offline_users": 0, "num_offline_users": 0}
    # offline organic training (bench_agents.py:169-173)
    for _ in range(args["num_organic_offline_users"]):
        train_env.reset(uid)
        uid += 1
        obs, _, _, _ = train_env.step(None)
        new_agent.train(obs, None, None, True)
    # offline organic + bandit training under the env's logging policy (bench_agents.py:176-190)
    for _ in range(args["num_offline_users"]):
        train_env.reset(uid)
        uid += 1
        new_obs, _, done, reward = train_env.step(None)
        while not done:
            old_obs = new_obs
            action, new_obs, reward, done, _ = train_env.step_offline(old_obs, reward, done)
            new_agent.train(old_obs, action, reward, False)
        old_obs = new_obs
        action, _, reward, done, _ = train_env.step_offline(old_obs, reward, done)
        new_agent.train(old_obs, action, reward, True)

    if args["epoch_with_random_reset"]:
        eval_env = deepcopy(env)
        eval_env.reset_random_seed(epoch)
    else:
        eval_env = env
    eval_begin = train_begin if args["epoch_with_random_reset"] else uid
    pol = device_policy(new_agent, eval_env.config.num_products, allow_probe=args.get("allow_probe", False))
    if pol is not None and pol["policy"] not in getattr(eval_env, "device_policies", (pol["policy"],)):
        pol = None          # (reco-gym-v0 runs uniform and table policies on the device; anything else through its gym protocol)
    if pol is not None:
        r = eval_env.generate_logs_counters(args["num_online_users"], pol["policy"], pol.get("table"),
                                            user_begin=eval_begin, epsilon=pol.get("epsilon", 0.0))
        successes, failures = r["successes"], r["failures"]
    else:
        data = eval_env.generate_logs(args["num_online_users"], agent=new_agent, user_begin=eval_begin)
        rewards = data[~np.isnan(data["a"].astype("float64"))]["c"]
        successes = np.sum(rewards)
        failures = rewards.shape[0] - successes
    return {AgentStats.SUCCESSES: successes, AgentStats.FAILURES: failures,
            "next_user": eval_begin + args["num_online_users"]}


def test_agent(env, agent, num_offline_users=1000, num_online_users=100, num_organic_offline_users=0,
               num_epochs=1, epoch_with_random_reset=False, with_cache=False, allow_probe=False, device_training=True):
    """-> (median, q0.025, q0.975) of Beta(successes+1, failures+1) (bench_agents.py:215-251)."""
    if with_cache:
        raise NotImplementedError("the on-disk training-log cache of the reference is out of scope")
    successes = failures = 0
    next_user = 0
    for epoch in range(num_epochs):
        res = _collect_stats({
            "user_begin": 0 if epoch_with_random_reset else next_user,
            "env": env, "agent": agent, "num_offline_users": num_offline_users,
            "num_online_users": num_online_users, "num_organic_offline_users": num_organic_offline_users,
            "epoch_with_random_reset": epoch_with_random_reset, "epoch": epoch, "allow_probe": allow_probe,
            "device_training": device_training,
        })
        successes += res[AgentStats.SUCCESSES]
        failures += res[AgentStats.FAILURES]
        next_user = res["next_user"]
    return (beta.ppf(0.500, successes + 1, failures + 1),
            beta.ppf(0.025, successes + 1, failures + 1),
            beta.ppf(0.975, successes + 1, failures + 1))


test_agent.__test__ = False  # not a pytest test
