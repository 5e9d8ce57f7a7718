import numpy as np
import torch
from torch import nn, optim

from ..envs.configuration import Configuration
from .abstract import Agent

# reference: recogym/agents/bandit_mf.py:10-18
bandit_mf_square_args = {
    "num_products": 10,
    "embed_dim": 5,
    "mini_batch_size": 32,
    "loss_function": nn.BCEWithLogitsLoss(),
    "optim_function": optim.RMSprop,
    "learning_rate": 0.01,
    "with_ps_all": False,
}


class BanditMFSquare(nn.Module, Agent):
    """Matrix-factorisation bandit: logit(a | last viewed v) = <product_embedding[a], user_embedding[v]>,
    trained on (v, a, click) mini-batches, acts greedily (reference: recogym/agents/bandit_mf.py:22-123).

    `act` depends on the last viewed product only, so the trained agent compiles to a P-entry device table
    (`device_policy`): table[v] = argmax_a logit(a | v), 'ps' = that logit, as the reference reports it."""

    def __init__(self, config=Configuration(bandit_mf_square_args)):
        nn.Module.__init__(self)
        Agent.__init__(self, config)
        P, D = self.config.num_products, self.config.embed_dim
        self.product_embedding = nn.Embedding(P, D)
        self.user_embedding = nn.Embedding(P, D)
        self.optimizer = self.config.optim_function(self.parameters(), lr=self.config.learning_rate)
        self.last_product_viewed = None
        self.curr_step = 0
        self.train_data = ([], [], [])
        self.all_products = np.arange(P)

    def forward(self, products, users=None):
        if users is None:
            users = np.full(products.shape[0], self.last_product_viewed)
        a = self.product_embedding(torch.as_tensor(np.asarray(products), dtype=torch.long))
        b = self.user_embedding(torch.as_tensor(np.asarray(users), dtype=torch.long))
        return torch.sum(a * b, dim=1)

    def get_logits(self):
        return self.forward(self.all_products)

    def update_lpv(self, observation):
        assert observation is not None and observation.sessions() is not None
        if observation.sessions():
            self.last_product_viewed = observation.sessions()[-1]["v"]

    def act(self, observation, reward, done):
        with torch.no_grad():
            self.update_lpv(observation)
            logits = self.get_logits()
            action = logits.argmax().item()
            ps_all = ()
            if self.config.with_ps_all:
                ps_all = np.zeros(self.config.num_products)
                ps_all[action] = 1.0
            return {**Agent.act(self, observation, reward, done), "a": action, "ps": logits[action], "ps-a": ps_all}

    def update_weights(self):
        lpvs, actions, rewards = self.train_data
        if len(lpvs) == 0:
            return
        assert len(lpvs) == len(actions) == len(rewards)
        self.optimizer.zero_grad()
        logit = self.forward(np.array(actions), np.array(lpvs))
        loss = self.config.loss_function(logit, torch.as_tensor(np.array(rewards), dtype=torch.float32))
        loss.backward()
        self.optimizer.step()

    def train(self, observation, action, reward, done=False):
        # as the reference: the sample that lands on a mini-batch boundary triggers the update and is dropped
        self.update_lpv(observation)
        self.curr_step += 1
        if self.curr_step % self.config.mini_batch_size == 0:
            self.update_weights()
            self.train_data = ([], [], [])
        elif action is not None and reward is not None:
            self.train_data[0].append(self.last_product_viewed)
            self.train_data[1].append(action["a"])
            self.train_data[2].append(reward)

    def device_policy(self, rows_per_block=256):
        """table[v] = argmax_a <product_embedding[a], user_embedding[v]> with the same fp32 expression as `act`
        (sum over the embedding dimension of the elementwise product), evaluated block-wise for all v."""
        P = self.config.num_products
        table = np.empty(P, np.int32)
        ps = np.empty(P, np.float32)
        with torch.no_grad():
            a = self.product_embedding.weight                      # (P, D)
            for v0 in range(0, P, rows_per_block):
                b = self.user_embedding.weight[v0:v0 + rows_per_block]          # (B, D)
                logits = torch.sum(a.unsqueeze(0) * b.unsqueeze(1), dim=2)      # (B, P)
                best = logits.argmax(dim=1)
                table[v0:v0 + b.shape[0]] = best.numpy()
                ps[v0:v0 + b.shape[0]] = logits.gather(1, best[:, None])[:, 0].numpy()
        return {"policy": "table", "table": table, "ps_table": ps}
