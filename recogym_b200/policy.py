"""Compiles Python agents into device policies for the fused rollout kernel.

An agent whose `act()` depends only on the user's last viewed product (OrganicCount, BanditCount,
BanditMFSquare of the reference) is a P-entry lookup table; it is obtained by probing the agent's
UNCHANGED `act()` once per product on a synthetic observation.  Agents whose action depends on the
whole history stay on the RecoEnv1 facade (one user at a time).
"""
from copy import deepcopy

import numpy as np

from .envs.context import DefaultContext
from .envs.observation import Observation
from .envs.session import OrganicSessions


def probe_last_view_table(agent, num_products):
    """table[v] = agent.act(observation whose last view is v)['a'] on a deep copy of the agent."""
    probe = deepcopy(agent)
    table = np.zeros(num_products, np.int32)
    for v in range(num_products):
        s = OrganicSessions()
        s.next(DefaultContext(0, 0), v)
        table[v] = int(probe.act(Observation(DefaultContext(1, 0), s), 0, False)["a"])
    return table


def device_policy(agent, num_products, allow_probe=False):
    """-> dict(policy='uniform') | dict(policy='table', table=int32[P]) | None (not compilable)."""
    if agent is None:
        return {"policy": "uniform"}
    if hasattr(agent, "device_policy"):
        return agent.device_policy()
    if allow_probe:
        return {"policy": "table", "table": probe_last_view_table(agent, num_products)}
    return None
