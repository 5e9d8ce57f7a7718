import numpy as np
from numpy.random.mtrand import RandomState

from .abstract import Agent

# reference: recogym/agents/epsilon_greedy.py:6-18
epsilon_greedy_args = {
    "epsilon": 0.01,
    "random_seed": np.random.randint(2 ** 31 - 1),
    "epsilon_pure_new": True,       # the exploring action differs from the greedy one
    "epsilon_select_worse": False,  # explore in proportion to 1 - ps-a of the wrapped agent
    "with_ps_all": False,
}


class EpsilonGreedy(Agent):
    """Epsilon-greedy selection policy around another agent (reference: recogym/agents/epsilon_greedy.py:21-71).
    The action dict carries 'greedy' (and 'h0', the greedy action, when exploring): evaluate_agent's
    *_GREEDY counters and the ALL_EXPLORATION_DATA training approaches read them."""

    def __init__(self, config, agent):
        super().__init__(config)
        self.agent = agent
        self.rng = RandomState(self.config.random_seed)

    def train(self, observation, action, reward, done=False):
        self.agent.train(observation, action, reward, done)

    def act(self, observation, reward, done):
        cfg = self.config
        greedy_action = self.agent.act(observation, reward, done)
        if self.rng.choice([True, False], p=[cfg.epsilon, 1.0 - cfg.epsilon]):
            if cfg.epsilon_select_worse:
                probas = 1.0 - greedy_action["ps-a"]
            else:
                probas = np.ones(cfg.num_products)
            if cfg.epsilon_pure_new:
                probas[greedy_action["a"]] = 0.0
            probas = probas / np.sum(probas)
            a = self.rng.choice(cfg.num_products, p=probas)
            return {**super().act(observation, reward, done), "a": a, "ps": cfg.epsilon * probas[a],
                    "ps-a": (cfg.epsilon * probas if cfg.with_ps_all else ()), "greedy": False,
                    "h0": greedy_action["a"]}
        return {**greedy_action, "greedy": True, "ps": (1.0 - cfg.epsilon) * greedy_action["ps"],
                "ps-a": ((1.0 - cfg.epsilon) * greedy_action["ps-a"] if cfg.with_ps_all else ())}

    def reset(self):
        self.agent.reset()

    def device_policy(self):
        """Around an agent that compiles to a last-view table: epsilon == 0 never explores (the wrapped agent's policy is
        the policy); 0 < epsilon < 1 with the default selection (epsilon_pure_new, not epsilon_select_worse) is the device
        policy RG_POLICY_EPS_TABLE -- the coin and the exploring choice are drawn from the env's policy stream (counter
        RNG keyed by user and event time) instead of this object's RandomState, as RandomAgent's device form does."""
        cfg = self.config
        if not hasattr(self.agent, "device_policy") or cfg.with_ps_all:
            return None
        pol = self.agent.device_policy()
        if pol is None:
            return None
        eps = float(cfg.epsilon)
        if eps == 0.0:
            return pol
        if pol.get("policy") == "table" and 0.0 < eps < 1.0 and cfg.epsilon_pure_new and not cfg.epsilon_select_worse \
                and int(cfg.num_products) >= 2:
            return {**pol, "policy": "eps_table", "epsilon": eps}
        return None
