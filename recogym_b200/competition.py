"""competition_score: the leaderboard entry point (reference: recogym/competition.py:15-96)."""
import datetime

from .constants import AgentStats
from .envs.configuration import Configuration
from .envs.reco_env_v1 import env_1_args
from .evaluate_agent import build_agent_init, gather_agent_stats


def competition_score(num_products, num_organic_users_to_train, num_users_to_train, num_users_to_score, random_seed,
                      latent_factor, num_flips, log_epsilon, sigma_omega, agent_class, agent_configs, agent_name,
                      with_cache=False):
    """Trains `agent_class` on logs of the OrganicUserEventCounterAgent logging policy (epsilon = log_epsilon) and
    scores it online: one-row DataFrame [q0.025, q0.500, q0.975, time]."""
    import pandas as pd
    from . import make
    from .agents import OrganicUserEventCounterAgent, organic_user_count_args
    std_env_args = {**env_1_args, "random_seed": random_seed, "num_products": num_products, "K": latent_factor,
                    "sigma_omega": sigma_omega, "number_of_flips": num_flips}
    env = make("reco-gym-v1")
    time_start = datetime.datetime.now()
    agent_stats = gather_agent_stats(
        env, std_env_args,
        {"agent": OrganicUserEventCounterAgent(Configuration({
            **organic_user_count_args, **std_env_args, "select_randomly": True, "epsilon": log_epsilon,
            "num_products": num_products}))},
        {**build_agent_init(agent_name, agent_class, {**agent_configs, "num_products": num_products})},
        (num_users_to_train,), num_users_to_score, 1, True, num_organic_users_to_train, with_cache)
    q0_025, q0_500, q0_975 = [], [], []
    for name in agent_stats[AgentStats.AGENTS]:
        v = agent_stats[AgentStats.AGENTS][name]
        q0_025.append(v[AgentStats.Q0_025][0]); q0_500.append(v[AgentStats.Q0_500][0]); q0_975.append(v[AgentStats.Q0_975][0])
    seconds = (datetime.datetime.now() - time_start).total_seconds()
    return pd.DataFrame({"q0.025": q0_025, "q0.500": q0_500, "q0.975": q0_975, "time": [seconds]})
