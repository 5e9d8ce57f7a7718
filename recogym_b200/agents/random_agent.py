import numpy as np
from numpy.random.mtrand import RandomState

from ..envs.configuration import Configuration
from .abstract import Agent

random_args = {
    "num_products": 10,
    "random_seed": np.random.randint(2 ** 31 - 1),
    "with_ps_all": False,
}


class RandomAgent(Agent):
    """Uniform random recommendations from the agent's own RandomState
    (reference: recogym/agents/random_agent.py:15-34).  Device equivalent: RG_POLICY_UNIFORM."""

    def __init__(self, config=Configuration(random_args)):
        super().__init__(config)
        self.rng = RandomState(config.random_seed)

    def act(self, observation, reward, done):
        P = self.config.num_products
        return {
            **super().act(observation, reward, done),
            "a": self.rng.choice(P),
            "ps": 1.0 / float(P),
            "ps-a": (np.ones(P) / P if self.config.with_ps_all else ()),
        }

    def device_policy(self):
        return {"policy": "uniform"}
