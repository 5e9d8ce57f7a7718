import numpy as np

from ..envs.configuration import Configuration
from .abstract import Agent

organic_count_args = {
    "num_products": 10,
    "with_ps_all": False,
}


class OrganicCount(Agent):
    """Recommends the product most often co-viewed with the user's last viewed product
    (reference: recogym/agents/organic_count.py:42-87).  `act` depends on the last viewed product
    only, so the agent compiles to a P-entry device table (`device_policy`)."""

    def __init__(self, config=Configuration(organic_count_args)):
        super().__init__(config)
        P = self.config.num_products
        self.co_counts = np.zeros((P, P))
        self.corr = None

    def act(self, observation, reward, done):
        self.update_lpv(observation)
        action = self.co_counts[self.last_product_viewed, :].argmax()
        if self.config.with_ps_all:
            ps_all = np.zeros(self.config.num_products)
            ps_all[action] = 1.0
        else:
            ps_all = ()
        return {**super().act(observation, reward, done), "a": action, "ps": 1.0, "ps-a": ps_all}

    def train(self, observation, action, reward, done=False):
        """co_counts += b b^T with b the view-count vector of the observation's sessions
        (organic_count.py:74-82); done as a sparse outer product over the distinct views."""
        sessions = observation.sessions()
        if sessions:
            views, cnt = np.unique(np.array([s["v"] for s in sessions], dtype=np.int64), return_counts=True)
            self.co_counts[np.ix_(views, views)] += np.outer(cnt, cnt).astype(self.co_counts.dtype)

    def update_lpv(self, observation):
        if observation.sessions():
            self.last_product_viewed = observation.sessions()[-1]["v"]

    def device_policy(self):
        return {"policy": "table", "table": self.co_counts.argmax(axis=1).astype(np.int32)}
