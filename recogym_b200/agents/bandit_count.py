import numpy as np

from ..envs.configuration import Configuration
from .abstract import Agent

bandit_count_args = {
    "num_products": 10,
    "with_ps_all": False,
}


class BanditCount(Agent):
    """Recommends the action with the best smoothed click rate given the last viewed product
    (reference: recogym/agents/bandit_count.py:14-79).  Compiles to a device table."""

    def __init__(self, config=Configuration(bandit_count_args)):
        super().__init__(config)
        P = self.config.num_products
        self.pulls_a = np.zeros((P, P))
        self.clicks_a = np.zeros((P, P))
        self.last_product_viewed = None
        self.ctr = (self.clicks_a + 1) / (self.pulls_a + 2)

    def act(self, observation, reward, done):
        self.update_lpv(observation)
        row = self.ctr[self.last_product_viewed, :]
        action = row.argmax()
        if self.config.with_ps_all:
            ps_all = np.zeros(self.config.num_products)
            ps_all[action] = 1.0
        else:
            ps_all = ()
        return {**super().act(observation, reward, done), "a": action, "ps": row[action], "ps-a": ps_all}

    def train(self, observation, action, reward, done=False):
        # NB the reference credits the pull to last_product_viewed BEFORE updating it from the
        # observation (bandit_count.py:54-57); kept.
        if action is not None and reward is not None:
            ix, jx = self.last_product_viewed, action["a"]
            self.update_lpv(observation)
            self.pulls_a[ix, jx] += 1
            self.clicks_a[ix, jx] += reward
            self.ctr[ix, jx] = (self.clicks_a[ix, jx] + 1) / (self.pulls_a[ix, jx] + 2)

    def update_lpv(self, observation):
        if observation.sessions():
            self.last_product_viewed = observation.sessions()[-1]["v"]

    def save(self, location):
        np.save(location + "pulls_a.npy", self.pulls_a)
        np.save(location + "clicks_a.npy", self.clicks_a)

    def load(self, location):
        self.pulls_a = np.load(location + "pulls_a.npy")
        self.clicks_a = np.load(location + "clicks_a.npy")

    def reset(self):
        pass

    def device_policy(self):
        table = self.ctr.argmax(axis=1).astype(np.int32)
        return {"policy": "table", "table": table, "ps_table": self.ctr[np.arange(len(table)), table]}
