"""A/B and evolution drivers (reference: recogym/evaluate_agent.py:51-146,158-182,732-750)."""
from copy import deepcopy

from scipy.stats import beta

from .constants import AgentInit, EvolutionCase, TrainingApproach
from .policy import device_policy


def build_agent_init(agent_key, ctor, def_args):
    """evaluate_agent.py:149-155"""
    return {agent_key: {AgentInit.CTOR: ctor, AgentInit.DEF_ARGS: def_args}}


def verify_agents(env, number_of_users, agents, allow_probe=False):
    """A/B test: every agent is rolled out on a fresh copy of the env (same seed, same users) and
    scored by the Beta posterior of its click rate over bandit rows, phantom rows included
    (evaluate_agent.py:732-750).  Compilable agents use one fused CUDA rollout each."""
    import pandas as pd
    stat = {"Agent": [], "0.025": [], "0.500": [], "0.975": []}
    for agent_id, agent in agents.items():
        e = deepcopy(env)
        pol = device_policy(agent, e.config.num_products, allow_probe=allow_probe)
        if pol is not None:
            r = e.generate_logs_counters(number_of_users, pol["policy"], pol.get("table"))
            successes, failures = r["successes"], r["failures"]
        else:
            data = e.generate_logs(number_of_users, agent)
            bandits = data[data["z"] == "bandit"]
            successes = bandits[bandits["c"] == 1].shape[0]
            failures = bandits[bandits["c"] == 0].shape[0]
        stat["Agent"].append(agent_id)
        stat["0.025"].append(beta.ppf(0.025, successes + 1, failures + 1))
        stat["0.500"].append(beta.ppf(0.500, successes + 1, failures + 1))
        stat["0.975"].append(beta.ppf(0.975, successes + 1, failures + 1))
    return pd.DataFrame().from_dict(stat)


def evaluate_agent(env, agent, num_initial_train_users=100, num_step_users=1000, num_steps=10,
                   training_approach=TrainingApproach.ALL_DATA, sliding_window_samples=10000):
    """Explore/exploit evolution (evaluate_agent.py:51-146) through the facade: the acting policy is
    frozen inside a step and swapped at the step boundary.  Host-serial (Python agents train per
    sample); the on-device variant is SURVEY.md 8f rank 1."""
    initial_agent = deepcopy(agent)
    unique_user_id = 0
    for u in range(num_initial_train_users):
        env.reset(unique_user_id + u)
        agent.reset()
        new_observation, reward, done, _ = env.step(None)
        while True:
            old_observation = new_observation
            if done:
                # the reference steps an already-stopped env once here (evaluate_agent.py:65-73);
                # a stopped user has no further dynamics, so only the act/train bookkeeping remains
                action, new_observation, reward, done, _ = env.step_offline(new_observation, reward, True)
                agent.train(old_observation, action, reward, done)
                break
            action, new_observation, reward, done, _ = env.step_offline(new_observation, reward, False)
            agent.train(old_observation, action, reward, done)
            if done:
                break
    unique_user_id += num_initial_train_users

    rewards = {EvolutionCase.SUCCESS: [], EvolutionCase.SUCCESS_GREEDY: [], EvolutionCase.FAILURE: [],
               EvolutionCase.FAILURE_GREEDY: [], EvolutionCase.ACTIONS: dict()}
    training_agent = deepcopy(agent)
    samples = 0
    for action_id in range(env.config.num_products):
        rewards[EvolutionCase.ACTIONS][action_id] = [0]

    for step in range(num_steps):
        successes = successes_greedy = failures = failures_greedy = 0
        for u in range(num_step_users):
            env.reset(unique_user_id + u)
            agent.reset()
            new_observation, reward, done, _ = env.step(None)
            while not done:
                old_observation = new_observation
                action = agent.act(old_observation, reward, done)
                new_observation, reward, done, info = env.step(action["a"])
                samples += 1
                if training_approach in (TrainingApproach.ALL_DATA, TrainingApproach.LAST_STEP):
                    update = True
                elif training_approach == TrainingApproach.SLIDING_WINDOW_ALL_DATA:
                    update = samples % sliding_window_samples == 0
                elif training_approach == TrainingApproach.ALL_EXPLORATION_DATA:
                    update = not action["greedy"]
                elif training_approach == TrainingApproach.SLIDING_WINDOW_EXPLORATION_DATA:
                    update = (not action["greedy"]) and samples % sliding_window_samples == 0
                else:
                    assert False, f"Unknown Training Approach: {training_approach}"
                if update:
                    training_agent.train(old_observation, action, reward, done)
                greedy = "greedy" in action and action["greedy"]
                if reward:
                    successes += 1
                    successes_greedy += 1 if greedy else 0
                    rewards[EvolutionCase.ACTIONS][action["a"]][-1] += 1
                else:
                    failures_greedy += 1 if greedy else 0
                    failures += 1
        unique_user_id += num_step_users
        agent = training_agent
        for action_id in range(env.config.num_products):
            rewards[EvolutionCase.ACTIONS][action_id].append(0)
        training_agent = deepcopy(initial_agent if training_approach == TrainingApproach.LAST_STEP else agent)
        rewards[EvolutionCase.SUCCESS].append(successes)
        rewards[EvolutionCase.SUCCESS_GREEDY].append(successes_greedy)
        rewards[EvolutionCase.FAILURE].append(failures)
        rewards[EvolutionCase.FAILURE_GREEDY].append(failures_greedy)
    return rewards
