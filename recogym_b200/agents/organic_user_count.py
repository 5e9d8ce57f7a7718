import numpy as np
from numpy.random.mtrand import RandomState

from ..envs.configuration import Configuration
from .abstract import Model, ModelBasedAgent, ModelBuilder, ViewsFeaturesProvider

# reference: recogym/agents/organic_user_count.py:7-27
organic_user_count_args = {
    "num_products": 10,
    "random_seed": np.random.randint(2 ** 31 - 1),
    "select_randomly": True,      # sample the action in proportion to the user's views (else argmax)
    "epsilon": 0.0,
    "exploit_explore": True,
    "reverse_pop": False,
    "weight_history_function": None,
    "with_ps_all": False,
}


class OrganicUserEventCounterModel(Model):
    """Recommends one of the products THIS user has viewed, in proportion to the view counts
    (reference: recogym/agents/organic_user_count.py:36-98)."""

    def __init__(self, config):
        super().__init__(config)
        self.rng = RandomState(self.config.random_seed)

    def act(self, observation, features):
        cfg = self.config
        features = features.flatten()
        explore = False
        if cfg.exploit_explore:
            explore = self.rng.choice([True, False], p=[cfg.epsilon, 1 - cfg.epsilon])
            if explore:  # uniform over the products the user has NOT viewed
                unseen = features == 0
                features[unseen] = 1
                features[~unseen] = 0
            proba = features / np.sum(features)
        else:
            features = cfg.epsilon + features
            proba = features / np.sum(features)
            if cfg.reverse_pop:
                proba = 1 - proba
                proba = proba / np.sum(proba)
        if cfg.select_randomly:
            action = self.rng.choice(cfg.num_products, p=proba)
            if cfg.exploit_explore:
                ps = (cfg.epsilon if explore else 1 - cfg.epsilon) * proba[action]
            else:
                ps = proba[action]
            ps_all = proba if cfg.with_ps_all else ()
        else:
            action = np.argmax(proba)
            ps = 1.0
            ps_all = ()
            if cfg.with_ps_all:
                ps_all = np.zeros(cfg.num_products)
                ps_all[action] = 1.0
        return {**super().act(observation, features), "a": action, "ps": ps, "ps-a": ps_all}


class OrganicUserEventCounterModelBuilder(ModelBuilder):
    def build(self):
        return ViewsFeaturesProvider(self.config, is_sparse=False), OrganicUserEventCounterModel(self.config)


class OrganicUserEventCounterAgent(ModelBasedAgent):
    """Counts the organic views of the CURRENT user and recommends among them
    (reference: recogym/agents/organic_user_count.py:101-116).  With the default configuration
    (exploit_explore, epsilon = 0, select_randomly, no history weighting) the agent compiles to the
    device policy RG_POLICY_USER_VIEWS: the per-user counts live in a device list, the action is numpy's
    inverse-CDF choice on the policy stream of the counter RNG."""

    def __init__(self, config=Configuration(organic_user_count_args)):
        super().__init__(config, OrganicUserEventCounterModelBuilder(config))

    def device_policy(self):
        c = self.config
        if (c.exploit_explore and c.select_randomly and float(c.epsilon) == 0.0
                and getattr(c, "weight_history_function", None) is None and not c.with_ps_all):
            return {"policy": "user_views"}
        return None
