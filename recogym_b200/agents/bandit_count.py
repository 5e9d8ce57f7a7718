import numpy as np

from ..envs.configuration import Configuration
from ._sparse import HostSparseCounts
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
        from ..logs import SPARSE_MIN_PRODUCTS
        # large catalogues (BASELINE configs[3]: P = 100 000 -> 3 x 80 GB dense): sparse statistics, same argmax
        self._sp = HostSparseCounts(P) if P >= SPARSE_MIN_PRODUCTS else None
        self.last_product_viewed = None
        if self._sp is None:
            self.pulls_a = np.zeros((P, P))
            self.clicks_a = np.zeros((P, P))
            self.ctr = (self.clicks_a + 1) / (self.pulls_a + 2)
        else:
            self.pulls_a = self.clicks_a = self.ctr = None

    def act(self, observation, reward, done):
        self.sync_host()
        self.update_lpv(observation)
        if self._sp is not None:
            action, ps = self._sp.bandit_argmax(self.last_product_viewed)
        else:
            row = self.ctr[self.last_product_viewed, :]
            action = row.argmax()
            ps = row[action]
        if self.config.with_ps_all:
            ps_all = np.zeros(self.config.num_products)
            ps_all[action] = 1.0
        else:
            ps_all = ()
        return {**super().act(observation, reward, done), "a": action, "ps": ps, "ps-a": ps_all}

    def train(self, observation, action, reward, done=False):
        # NB the reference credits the pull to last_product_viewed BEFORE updating it from the
        # observation (bandit_count.py:54-57); kept.
        self.sync_host()
        if action is not None and reward is not None:
            ix, jx = self.last_product_viewed, action["a"]
            self.update_lpv(observation)
            if self._sp is not None:
                if ix is None:          # numpy's pulls_a[None, jx] += 1 touches the whole ROW jx
                    self._sp.add_row(jx, 1, int(reward))
                else:
                    self._sp.add(ix, jx, 1, int(reward))
                return
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

    # ---- device residency (recogym_b200/logs.py) --------------------------------------------------------------
    _stats = None

    def _device_stats(self, device):
        from ..logs import CountStats
        import torch
        if self._stats is None:
            self._stats = CountStats(self.config.num_products, device, organic=False, bandit=True,
                                     sparse=self._sp is not None)
            if self._sp is None:
                self._stats.pulls.copy_(torch.as_tensor(self.pulls_a.astype(np.int64)))
                self._stats.clicks.copy_(torch.as_tensor(self.clicks_a.astype(np.int64)))
            elif self._sp.n_cells() or self._sp.row_a:
                self._stats.load_host(None, self._sp)
        return self._stats

    def device_train(self, log):
        """`train` for every bandit row of a DeviceLog (scatter-add kernel).  The log must have been generated with
        prev_last_view = self.device_last_view()."""
        stats = self._device_stats(log.env.device)
        stats.train(log)
        self.last_product_viewed = None if stats.last_view < 0 else int(stats.last_view)

    def device_train_sharded(self, log, final_view, group=None):
        """`log` holds THIS rank's shard of the users of a training log, its `lvp` column already chained to the lower
        ranks' shards (parallel.resolve_carry); `final_view` = last_product_viewed after the WHOLE log.  The shard's
        statistics go into a zero copy; the sum over the ranks is added to the replicated statistics."""
        stats = self._device_stats(log.env.device)
        delta = stats.zeros_like()
        delta.train(log)
        stats.add_reduced(delta, group)
        stats.last_view = int(final_view)
        self.last_product_viewed = None if final_view < 0 else int(final_view)

    def device_last_view(self):
        return -1 if self.last_product_viewed is None else int(self.last_product_viewed)

    def sync_host(self):
        if self._stats is not None:
            if self._sp is None:
                self.pulls_a = self._stats.pulls.cpu().numpy().astype(np.float64)
                self.clicks_a = self._stats.clicks.cpu().numpy().astype(np.float64)
                self.ctr = (self.clicks_a + 1) / (self.pulls_a + 2)
            else:
                self._sp = self._stats.to_host()[1]
            self._stats = None

    def device_policy(self):
        if self._stats is not None:
            return self._stats.bandit_count_policy()
        if self._sp is not None:
            table, ps = self._sp.bandit_table()
            return {"policy": "table", "table": table, "ps_table": ps}
        table = self.ctr.argmax(axis=1).astype(np.int32)
        return {"policy": "table", "table": table, "ps_table": self.ctr[np.arange(len(table)), table]}
