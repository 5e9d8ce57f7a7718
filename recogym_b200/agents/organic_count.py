import numpy as np

from ..envs.configuration import Configuration
from ._sparse import HostSparseCounts
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
        from ..logs import SPARSE_MIN_PRODUCTS
        # large catalogues: the P x P matrix (8 P^2 bytes) gives way to a sparse form with the same argmax (agents/_sparse.py)
        self._sp = HostSparseCounts(P) if P >= SPARSE_MIN_PRODUCTS else None
        self.co_counts = np.zeros((P, P)) if self._sp is None else None
        self.corr = None

    def act(self, observation, reward, done):
        self.sync_host()
        self.update_lpv(observation)
        if self._sp is not None:
            action = self._sp.organic_argmax(self.last_product_viewed)
        else:
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
        self.sync_host()
        sessions = observation.sessions()
        if sessions:
            views, cnt = np.unique(np.array([s["v"] for s in sessions], dtype=np.int64), return_counts=True)
            if self._sp is not None:
                for vi, ci in zip(views.tolist(), cnt.tolist()):
                    for vj, cj in zip(views.tolist(), cnt.tolist()):
                        self._sp.add(vi, vj, ci * cj)
                return
            self.co_counts[np.ix_(views, views)] += np.outer(cnt, cnt).astype(self.co_counts.dtype)

    def update_lpv(self, observation):
        if observation.sessions():
            self.last_product_viewed = observation.sessions()[-1]["v"]

    # ---- device residency (recogym_b200/logs.py): train() over a whole fused log, act() as a table ----------
    _stats = None

    def _device_stats(self, device):
        from ..logs import CountStats
        import torch
        if self._stats is None:
            self._stats = CountStats(self.config.num_products, device, organic=True, bandit=False,
                                     sparse=self._sp is not None)
            if self._sp is None:
                self._stats.co_counts.copy_(torch.as_tensor(self.co_counts.astype(np.int64)))
            elif self._sp.n_cells():
                self._stats.load_host(self._sp, None)
        return self._stats

    def device_train(self, log):
        """`train` for every observation of a DeviceLog (scatter-add kernel); the statistics then live on the GPU
        until `sync_host()`."""
        self._device_stats(log.env.device).train(log)

    def device_train_sharded(self, log, final_view, group=None):
        """`log` holds THIS rank's shard of the users of a training log: the shard's statistics go into a zero copy and
        the sum over the ranks is added to the (replicated) statistics -- one all-reduce per step boundary."""
        stats = self._device_stats(log.env.device)
        delta = stats.zeros_like()
        delta.train(log)
        stats.add_reduced(delta, group)

    def sync_host(self):
        if self._stats is not None:
            if self._sp is None:
                self.co_counts = self._stats.co_counts.cpu().numpy().astype(np.float64)
            else:
                self._sp = self._stats.to_host()[0]
            self._stats = None

    def device_policy(self):
        if self._stats is not None:
            pol = self._stats.organic_count_policy()
            return {"policy": "table", "table": pol["table"]}
        if self._sp is not None:
            return {"policy": "table", "table": self._sp.organic_table()}
        return {"policy": "table", "table": self.co_counts.argmax(axis=1).astype(np.int32)}
