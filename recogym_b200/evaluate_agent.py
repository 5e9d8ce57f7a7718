"""A/B, sweep and evolution drivers (reference: recogym/evaluate_agent.py:51-447,732-750).

The reference spreads sweep points over a `multiprocessing.Pool` (evaluate_agent.py:255-277,416-445) because
one rollout occupies one core; here one rollout is one fused CUDA launch over all users, so the sweeps run
in-process, point after point (a forked child could not use the CUDA context anyway)."""
import time
from copy import deepcopy

from scipy.stats import beta

from .constants import AgentInit, AgentStats, EvolutionCase, TrainingApproach
from .envs.configuration import Configuration
from .policy import device_policy

EpsilonPrecision = 2
EvolutionEpsilons = (0.00, 0.01, 0.02, 0.03, 0.05, 0.08)


def build_agent_init(agent_key, ctor, def_args):
    """evaluate_agent.py:149-155"""
    return {agent_key: {AgentInit.CTOR: ctor, AgentInit.DEF_ARGS: def_args}}


def verify_agents(env, number_of_users, agents, allow_probe=False, user_begin=0):
    """A/B test: every agent is rolled out on a fresh copy of the env (same seed, same users) and
    scored by the Beta posterior of its click rate over bandit rows, phantom rows included
    (evaluate_agent.py:732-750).  Compilable agents use one fused CUDA rollout each.

    user_begin: first counter-RNG key of the A/B users.  The reference copies the env in its CURRENT state, i.e. after
    whatever users the agents were trained on; pass the number of users already consumed to keep the A/B users disjoint
    from the training users (a rollout here is a pure function of (seed, user key))."""
    import pandas as pd
    stat = {"Agent": [], "0.025": [], "0.500": [], "0.975": []}
    for agent_id, agent in agents.items():
        e = deepcopy(env)
        pol = device_policy(agent, e.config.num_products, allow_probe=allow_probe)
        if pol is not None and pol["policy"] not in getattr(e, "device_policies", (pol["policy"],)):
            pol = None      # (reco-gym-v0 runs uniform and table policies on the device; anything else through its gym protocol)
        if pol is not None:
            r = e.generate_logs_counters(number_of_users, pol["policy"], pol.get("table"), user_begin=user_begin,
                                         epsilon=pol.get("epsilon", 0.0))
            successes, failures = r["successes"], r["failures"]
        else:
            data = e.generate_logs(number_of_users, agent, user_begin=user_begin)
            bandits = data[data["z"] == "bandit"]
            successes = bandits[bandits["c"] == 1].shape[0]
            failures = bandits[bandits["c"] == 0].shape[0]
        stat["Agent"].append(agent_id)
        stat["0.025"].append(beta.ppf(0.025, successes + 1, failures + 1))
        stat["0.500"].append(beta.ppf(0.500, successes + 1, failures + 1))
        stat["0.975"].append(beta.ppf(0.975, successes + 1, failures + 1))
    return pd.DataFrame().from_dict(stat)


def _device_evolution_agent(agent):
    """The count agent behind `agent` when the whole evolution loop can stay on the GPU: OrganicCount / BanditCount,
    bare or wrapped in an EpsilonGreedy with the default selection rule (epsilon = 0: every action is 'greedy';
    0 < epsilon < 1: the device policy RG_POLICY_EPS_TABLE)."""
    from .agents import BanditCount, EpsilonGreedy, OrganicCount
    inner = agent
    if isinstance(agent, EpsilonGreedy):
        cfg = agent.config
        eps = float(cfg.epsilon)
        if cfg.with_ps_all or not (0.0 <= eps < 1.0) or (eps > 0.0 and (not cfg.epsilon_pure_new or cfg.epsilon_select_worse
                                                                           or int(cfg.num_products) < 2)):
            return None
        inner = agent.agent
    if type(inner) in (OrganicCount, BanditCount) and not inner.config.with_ps_all:
        return inner
    return None


def _evaluate_agent_device(env, agent, inner, num_initial_train_users, num_step_users, num_steps, training_approach,
                           math_mode=None, sliding_window_samples=10000, group=None, sharded=None):
    """evaluate_agent (evaluate_agent.py:51-146) with every phase on the GPU: the warm-up users are rolled out under
    the env's logging policy and the agent trains on their log (RG_TRAIN_EVOLUTION_INIT); each evolution step is ONE
    fused rollout of the frozen acting policy -- the table, or EpsilonGreedy(epsilon) around it (RG_POLICY_EPS_TABLE) --
    whose ordered log feeds the training copy (RG_TRAIN_EVOLUTION_STEP, all of its train calls or the ones the
    TrainingApproach selects: logs.MaskedTrainingLog) through the scatter-add kernels; the next acting table is a row
    arg-max kernel.  'greedy' of a logged action: it equals the acting table's (an exploring action never does).

    Multi-GPU (torch.distributed initialised, one process per GPU): the users of the warm-up and of
    every step are sharded over the ranks in contiguous id ranges; each rank rolls out and logs its shard, collects the
    shard's training statistics into a zero copy, and the sums over the ranks are added to the replicated statistics at
    the step boundary (`device_train_sharded`: one all-reduce per matrix, an all-gather of cells for the sparse tables),
    so every rank computes the same next acting table.  What the reference's strictly sequential loop threads through
    the users -- BanditCount's last viewed product and the running sample count of the sliding windows -- is chained
    across the shards with two small all-gathers (parallel.resolve_carry).  Counters are all-reduced per step.  The
    result equals the single-GPU one for any number of ranks (scripts/dist_evolution_check.py)."""
    import numpy as np
    import torch
    from . import _lib as L
    from .logs import MaskedTrainingLog, generate_logs_device
    from .parallel import allgather_ints, allreduce_counters, dist_info, resolve_carry, shard_range
    rank, world = dist_info(group)
    if sharded is None:
        sharded = world > 1
    if sharded and min(num_initial_train_users if num_initial_train_users > 0 else world, num_step_users) < world:
        sharded = False                  # fewer users than ranks: every rank computes the whole thing
    if not sharded:
        rank, world = 0, 1
    mm = L.MATH_TENSOR if math_mode is None else math_mode
    de = env.device_env
    P = env.config.num_products
    wrapped = inner is not agent         # under EpsilonGreedy the action dict carries 'greedy' (epsilon_greedy.py:44-62)
    eps = float(agent.config.epsilon) if wrapped else 0.0
    explored_only = training_approach in (TrainingApproach.ALL_EXPLORATION_DATA,
                                          TrainingApproach.SLIDING_WINDOW_EXPLORATION_DATA)
    window = int(sliding_window_samples) if training_approach in (TrainingApproach.SLIDING_WINDOW_ALL_DATA,
                                                                  TrainingApproach.SLIDING_WINDOW_EXPLORATION_DATA) else 0
    samples = 0
    lpv = lambda a: a.device_last_view() if hasattr(a, "device_last_view") else -1
    initial_agent = deepcopy(inner)
    log_pol = device_policy(env.agent, P) if env.agent else {"policy": "uniform"}
    def chain(view, n, prev0):
        """sharded logs are derived with prev_last_view = -1: point the rows that start from "no previous view" at the view
        the lower ranks' shards end with; -> last_product_viewed after the whole log"""
        carries = [c[0] for c in allgather_ints([int(view.carry.item())], de.torch_device, group)]
        prev, final = resolve_carry(carries, rank, prev0)
        if prev >= 0 and n:
            col = view.lvp[:n]
            col[col == -1] = prev
        return final

    if num_initial_train_users > 0:
        first, count = shard_range(num_initial_train_users, rank, world, 0)
        log = generate_logs_device(de, count, log_pol["policy"], log_pol.get("table"), None, first, 0, mm,
                                   prev_last_view=lpv(inner) if world == 1 else -1, train_mode=L.TRAIN_EVOLUTION_INIT,
                                   epsilon=log_pol.get("epsilon", 0.0))
        if world == 1:
            inner.device_train(log)
        else:
            inner.device_train_sharded(log, chain(log, log.rows, lpv(inner)), group)
    unique_user_id = num_initial_train_users
    rewards = {EvolutionCase.SUCCESS: [], EvolutionCase.SUCCESS_GREEDY: [], EvolutionCase.FAILURE: [],
               EvolutionCase.FAILURE_GREEDY: [], EvolutionCase.ACTIONS: dict()}
    acting = inner
    training = deepcopy(inner)
    per_action = [[0] for _ in range(P)]
    for step in range(num_steps):
        pol = acting.device_policy()
        if not (hasattr(pol["table"], "is_cuda") and pol["table"].is_cuda):
            pol = {**pol, "table": torch.as_tensor(np.ascontiguousarray(pol["table"], np.int32), device=de.torch_device)}
        first, count = shard_range(num_step_users, rank, world, unique_user_id)
        log = generate_logs_device(de, count, "eps_table" if eps > 0.0 else pol["policy"], pol.get("table"),
                                   pol.get("ps_table"), first, 0, mm, prev_last_view=lpv(training) if world == 1 else -1,
                                   train_mode=L.TRAIN_EVOLUTION_STEP, epsilon=eps)
        n = log.rows
        real = (log.z[:n] == 1) & ((log.flags[:n] & 1) == 0)
        is_greedy = real & (log.a[:n] == pol["table"][log.lv[:n].clamp(min=0).long()]) if wrapped else None
        clicked = log.a[:n][log.c[:n] == 1].long()
        # counters of the step: [bandit events, clicks, greedy clicks, greedy non-clicks | clicks per action]
        counters = torch.zeros(4 + P, dtype=torch.int64, device=de.torch_device)
        counters[0:2] = torch.stack([log.totals[3], log.totals[4]])
        if wrapped:
            counters[2] = (is_greedy & (log.c[:n] == 1)).sum()
            counters[3] = (is_greedy & (log.c[:n] != 1)).sum()
        counters[4:] = torch.bincount(clicked, minlength=P)
        n_bandit_local = int(log.totals[3].item())
        if world > 1:
            allreduce_counters(counters, group=group)
        counters = counters.cpu().numpy()
        successes, failures = int(counters[1]), int(counters[0] - counters[1])   # real bandit events only: the loop stops at done
        hist = counters[4:]
        if explored_only and not wrapped:
            raise KeyError("greedy")     # the reference reads action['greedy'], which only EpsilonGreedy provides
        if world > 1:
            # the running sample count before this rank's shard: the lower ranks' bandit events of this step
            before = samples + sum(c[0] for c in allgather_ints([n_bandit_local], de.torch_device, group)[:rank])
            view = MaskedTrainingLog(log, pol["table"], explored_only, window, before, -1) if explored_only or window else log
            training.device_train_sharded(view, chain(view, n, lpv(training)), group)
            samples += int(counters[0])
        elif explored_only or window:
            view = MaskedTrainingLog(log, pol["table"], explored_only, window, samples, lpv(training))
            samples = view.n_samples
            training.device_train(view)
        else:
            samples += int(counters[0])
            training.device_train(log)
        unique_user_id += num_step_users
        acting = training
        for a in range(P):
            per_action[a][-1] += int(hist[a])
            per_action[a].append(0)
        training = deepcopy(initial_agent if training_approach == TrainingApproach.LAST_STEP else acting)
        sg, fg = int(counters[2]), int(counters[3])
        rewards[EvolutionCase.SUCCESS].append(successes)
        rewards[EvolutionCase.SUCCESS_GREEDY].append(sg)
        rewards[EvolutionCase.FAILURE].append(failures)
        rewards[EvolutionCase.FAILURE_GREEDY].append(fg)
    rewards[EvolutionCase.ACTIONS] = {a: per_action[a] for a in range(P)}
    return rewards


def evaluate_agent(env, agent, num_initial_train_users=100, num_step_users=1000, num_steps=10,
                   training_approach=TrainingApproach.ALL_DATA, sliding_window_samples=10000, device=None, group=None):
    """Explore/exploit evolution (evaluate_agent.py:51-146): the acting policy is frozen inside a step and swapped at
    the step boundary.  device=None: count agents (bare or under a never-exploring EpsilonGreedy) with the ALL_DATA /
    LAST_STEP approaches run entirely on the GPU (`_evaluate_agent_device`); anything else goes user by user through
    the gym facade with the Python agents, exactly as the reference does.  With torch.distributed initialised (one
    process per GPU) the device loop shards every phase's users over the ranks of `group` and reduces the training
    statistics at the step boundaries; every rank returns the same dict."""
    inner = _device_evolution_agent(agent)
    if device is None:
        device = (inner is not None and getattr(env, "supports_device_training", True)
                  and (inner is not agent or training_approach in (TrainingApproach.ALL_DATA, TrainingApproach.LAST_STEP,
                                                                   TrainingApproach.SLIDING_WINDOW_ALL_DATA))
                  and (not env.agent or device_policy(env.agent, env.config.num_products) is not None))
    if device:
        if inner is None:
            raise ValueError("the device evolution loop needs a count agent (OrganicCount / BanditCount)")
        if not getattr(env, "supports_device_training", True):
            raise ValueError("the device evolution loop runs on reco-gym-v1; %s goes through the gym protocol" % type(env).__name__)
        return _evaluate_agent_device(env, agent, inner, num_initial_train_users, num_step_users, num_steps,
                                      training_approach, sliding_window_samples=sliding_window_samples, group=group)
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


def build_agents(agents_init_data, new_env_args):
    """evaluate_agent.py:284-296: agent = ctor(Configuration({**def_args, **env_args}))."""
    agents = dict()
    for agent_key, init in agents_init_data.items():
        agents[agent_key] = init[AgentInit.CTOR](Configuration({**init[AgentInit.DEF_ARGS], **new_env_args}))
    return agents


def gather_agent_stats(env, env_args, extra_env_args, agents_init_data,
                       user_samples=(100, 1000, 2000, 3000, 5000, 8000, 10000, 13000, 14000, 15000),
                       num_online_users=15000, num_epochs=1, epoch_with_random_reset=False,
                       num_organic_offline_users=100, with_cache=False, verbose=False, allow_probe=False):
    """CTR quantiles of every agent as a function of the number of training users
    (evaluate_agent.py:185-281).  Returns {SAMPLES: user_samples, AGENTS: {name: {Q0_025, Q0_500, Q0_975}}}."""
    from .bench_agents import test_agent
    new_env_args = {**env_args, **extra_env_args}
    new_env = deepcopy(env)
    new_env.init_gym(new_env_args)
    agents = build_agents(agents_init_data, new_env_args)
    agent_stats = {AgentStats.SAMPLES: user_samples, AgentStats.AGENTS: dict()}
    for agent_key, agent in agents.items():
        stats = {AgentStats.Q0_025: [], AgentStats.Q0_500: [], AgentStats.Q0_975: []}
        for num_offline_users in user_samples:
            start = time.time()
            q500, q025, q975 = test_agent(deepcopy(new_env), deepcopy(agent), num_offline_users, num_online_users,
                                          num_organic_offline_users, num_epochs, epoch_with_random_reset, with_cache,
                                          allow_probe=allow_probe)
            if verbose:
                print(f"Agent {agent_key}: {num_offline_users} offline users ({time.time() - start:.2f}s)")
            stats[AgentStats.Q0_025].append(q025)
            stats[AgentStats.Q0_500].append(q500)
            stats[AgentStats.Q0_975].append(q975)
        agent_stats[AgentStats.AGENTS][agent_key] = stats
    return agent_stats


def generate_epsilons(epsilon_step=0.02, iterations=6):
    return [0.00, 0.01, 0.02, 0.03, 0.05, 0.08]      # evaluate_agent.py:299-300 (the arguments are unused there too)


def format_epsilon(epsilon):
    return ("{0:." + f"{EpsilonPrecision}" + "f}").format(round(epsilon, EpsilonPrecision))


def gather_exploration_stats(env, env_args, extra_env_args, agents_init_data, training_approach,
                             num_initial_train_users=1000, num_step_users=1000, epsilons=EvolutionEpsilons,
                             num_evolution_steps=6):
    """Evolution of every agent under EpsilonGreedy(epsilon) for each epsilon (evaluate_agent.py:344-447).
    Returns {agent: {format_epsilon(eps): {SUCCESS, SUCCESS_GREEDY, FAILURE, FAILURE_GREEDY, ACTIONS}}}."""
    from .agents import EpsilonGreedy, epsilon_greedy_args
    new_env_args = {**env_args, **extra_env_args}
    new_env = deepcopy(env)
    new_env.init_gym(new_env_args)
    agents = build_agents(agents_init_data, new_env_args)
    out = dict()
    for agent_key, agent in agents.items():
        agent_stats = dict()
        for epsilon in epsilons:
            eg = EpsilonGreedy(Configuration({**epsilon_greedy_args, **new_env_args, "epsilon": epsilon}),
                               deepcopy(agent))
            rewards = evaluate_agent(deepcopy(new_env), eg, num_initial_train_users, num_step_users,
                                     num_evolution_steps, training_approach)
            assert len(rewards[EvolutionCase.SUCCESS]) == len(rewards[EvolutionCase.FAILURE]) == num_evolution_steps
            agent_stats[format_epsilon(epsilon)] = {k: rewards[k] for k in (
                EvolutionCase.SUCCESS, EvolutionCase.SUCCESS_GREEDY, EvolutionCase.FAILURE,
                EvolutionCase.FAILURE_GREEDY, EvolutionCase.ACTIONS)}
        out[agent_key] = agent_stats
    return out


# ---- off-policy estimators over a logged DataFrame (reference: recogym/evaluate_agent.py:753-908) ---------------------
def to_categorical(y, num_classes=None, dtype="float32"):
    """One-hot rows (the Keras helper the reference carries, evaluate_agent.py:36-48)."""
    import numpy as np
    y = np.array(y, dtype="int")
    input_shape = y.shape
    if input_shape and input_shape[-1] == 1 and len(input_shape) > 1:
        input_shape = tuple(input_shape[:-1])
    y = y.ravel()
    if not num_classes:
        num_classes = np.max(y) + 1
    out = np.zeros((y.shape[0], num_classes), dtype=dtype)
    out[np.arange(y.shape[0]), y] = 1
    return out.reshape(input_shape + (num_classes,))


def _replay_policy_probabilities(agent, reco_log):
    """Per bandit row of the users 0 .. max(u)-1 (the reference's `for u in range(max(reco_log.u))` leaves the last user
    out): (reward c, pi(a | context) of `agent`, logged ps, top-k helper).  Table agents (`device_policy()` is a
    last-view table) are evaluated in one vectorised pass -- pi = [a == table[last view]], the last view carried by
    `Agent.update_lpv` -- any other agent is replayed row by row through its own act() with 'ps-a', as the reference
    does (evaluate_agent.py:753-811)."""
    import numpy as np
    import pandas as pd
    from .envs.context import DefaultContext
    from .envs.observation import Observation
    from .envs.session import OrganicSessions
    u = reco_log["u"].to_numpy(dtype="int64")
    keep = u < int(u.max())
    z = (reco_log["z"].to_numpy() == "bandit")
    pol = agent.device_policy() if hasattr(agent, "device_policy") else None
    if pol is not None and pol.get("policy") == "table":
        table = np.asarray(pol["table"].cpu().numpy() if hasattr(pol["table"], "cpu") else pol["table"])
        v = reco_log["v"].to_numpy(dtype="float64", na_value=np.nan)
        lv = pd.Series(np.where(z, np.nan, v)).groupby(u).ffill().to_numpy()
        rows = np.flatnonzero(z & keep)
        a = reco_log["a"].to_numpy(dtype="float64", na_value=np.nan)[rows].astype(np.int64)
        pi = (table[lv[rows].astype(np.int64)] == a).astype(np.float64)
        return (reco_log["c"].to_numpy(dtype="float64")[rows], pi, reco_log["ps"].to_numpy(dtype="float64")[rows])
    rewards, probs, pss = [], [], []
    for user in range(int(u.max())):
        rows = reco_log[reco_log["u"] == user]
        session = OrganicSessions()
        agent.reset()
        for t, zz, v, a, c, ps in zip(rows["t"], rows["z"], rows["v"], rows["a"], rows["c"], rows["ps"]):
            if zz == "organic":
                session.next(DefaultContext(t, user), int(v))
            else:
                prob = agent.act(Observation(DefaultContext(t, user), session), 0, False)["ps-a"]
                if len(prob):
                    rewards.append(float(c)); probs.append(float(prob[int(a)])); pss.append(float(ps))
                session = OrganicSessions()
    return np.array(rewards), np.array(probs), np.array(pss)


def evaluate_IPS(agent, reco_log):
    """Per-row inverse-propensity terms c * pi(a) / ps (evaluate_agent.py:753-781)."""
    c, pi, ps = _replay_policy_probabilities(agent, reco_log)
    return list(c * pi / ps)


def evaluate_SNIPS(agent, reco_log):
    """-> (rewards, p_ratio) (evaluate_agent.py:784-811)."""
    c, pi, ps = _replay_policy_probabilities(agent, reco_log)
    return list(c), list(pi / ps)


def evaluate_recall_at_k(agent, reco_log, k=5):
    """Hit list of recall@k (evaluate_agent.py:852-893): for every un-clicked bandit row of the users 0 .. max(u)-1 that
    is directly followed by an organic row of the same user, 1 if the product viewed next is among the top k of the
    agent's action probabilities ('ps-a'; numpy.argpartition decides ties, as in the reference), else 0.  Table agents
    are evaluated in one vectorised pass, any other agent is replayed row by row through its own act()."""
    import numpy as np
    import pandas as pd
    from .envs.context import DefaultContext
    from .envs.observation import Observation
    from .envs.session import OrganicSessions
    from .logs import one_hot_top_k_table
    u = reco_log["u"].to_numpy(dtype="int64")
    z = (reco_log["z"].to_numpy() == "bandit")
    pol = agent.device_policy() if hasattr(agent, "device_policy") else None
    if pol is not None and pol.get("policy") == "table":
        table = np.asarray(pol["table"].cpu().numpy() if hasattr(pol["table"], "cpu") else pol["table"])
        v = reco_log["v"].to_numpy(dtype="float64", na_value=np.nan)
        c = reco_log["c"].to_numpy(dtype="float64")
        lv = pd.Series(np.where(z, np.nan, v)).groupby(u).ffill().to_numpy()
        nxt = np.arange(len(u)) + 1
        ok = z & (u < int(u.max()))
        ok[-1] = False
        rows = np.flatnonzero(ok)
        rows = rows[(u[rows + 1] == u[rows]) & ~z[rows + 1] & ~(c[rows] > 0)]
        topk = one_hot_top_k_table(len(table), k, table)
        g = table[lv[rows].astype(np.int64)]
        return [int(x) for x in (topk[g] == v[rows + 1].astype(np.int64)[:, None]).any(axis=1)]
    hits = []
    for user in range(int(u.max())):
        rows = reco_log[reco_log["u"] == user]
        zz, vv, cc, tt = list(rows["z"]), np.array(rows["v"]), np.array(rows["c"]), np.array(rows["t"])
        session = OrganicSessions()
        agent.reset()
        for j in range(len(zz)):
            if zz[j] == "organic":
                session.next(DefaultContext(tt[j], user), int(vv[j]))
                continue
            prob = agent.act(Observation(DefaultContext(tt[j], user), session), 0, False)["ps-a"]
            if j + 1 < len(zz) and zz[j + 1] == "organic" and not cc[j]:
                hits.append(1 if int(vv[j + 1]) in set(np.argpartition(prob, -k)[-k:]) else 0)
            session = OrganicSessions()
    return hits


def verify_agents_recall_at_k(reco_log, agents, k=5):
    """evaluate_agent.py:896-908: mean +- 2 standard errors of the recall@k hits per agent."""
    import numpy as np
    rows = []
    for agent_id, agent in agents.items():
        hits = np.asarray(evaluate_recall_at_k(agent, reco_log, k=k))
        rows.append((agent_id, np.mean(hits), np.std(hits) / np.sqrt(len(hits))))
    return _estimator_frame(rows)


def _estimator_frame(rows):
    import pandas as pd
    return pd.DataFrame().from_dict({"Agent": [r[0] for r in rows], "0.025": [r[1] - 2 * r[2] for r in rows],
                                     "0.500": [r[1] for r in rows], "0.975": [r[1] + 2 * r[2] for r in rows]})


def verify_agents_IPS(reco_log, agents):
    """evaluate_agent.py:813-828: mean +- 2 standard errors of the IPS terms per agent."""
    import numpy as np
    rows = []
    for agent_id, agent in agents.items():
        ee = np.asarray(evaluate_IPS(agent, reco_log))
        rows.append((agent_id, np.mean(ee), np.std(ee) / np.sqrt(len(ee))))
    return _estimator_frame(rows)


def verify_agents_SNIPS(reco_log, agents):
    """evaluate_agent.py:831-848: self-normalised estimate sum(c * ratio) / sum(ratio)."""
    import numpy as np
    rows = []
    for agent_id, agent in agents.items():
        rewards, ratio = (np.asarray(x) for x in evaluate_SNIPS(agent, reco_log))
        ee = rewards * ratio
        rows.append((agent_id, np.sum(ee) / np.sum(ratio), np.std(ee) / np.sqrt(len(ee))))
    return _estimator_frame(rows)
