"""recogym_b200 -- B200-native reco-gym-v1 simulation hot path behind the reference's Python API.

Exports the names a user of `recogym` imports for this path (reference: recogym/__init__.py:1-45):
value types, env args, enums, `test_agent`, `verify_agents`, `evaluate_agent`, `build_agent_init`,
and `make('reco-gym-v1')` / `make('reco-gym-v0')` in place of gym's registry (gym is not a dependency).
"""
from .envs import (Configuration, Context, DefaultContext, Observation, Session, OrganicSessions, RecoEnv1, RecoEnv0,
                   env_args, env_1_args, env_0_args, DeviceEnv, DeviceEnv0, BatchState, rollout_log, DefaultTimeGenerator,
                   NormalTimeGenerator)
from .constants import AgentStats, AgentInit, EvolutionCase, TrainingApproach, RoiMetrics
from .bench_agents import test_agent
from .evaluate_agent import (evaluate_agent, build_agent_init, build_agents, verify_agents, gather_agent_stats,
                             gather_exploration_stats, verify_agents_IPS, verify_agents_SNIPS, evaluate_IPS,
                             evaluate_SNIPS, evaluate_recall_at_k, verify_agents_recall_at_k, to_categorical)
from .competition import competition_score
from . import agents
from ._lib import MATH_EXACT, MATH_FAST, MATH_TENSOR, RecoGymB200Error

_REGISTRY = {"reco-gym-v1": RecoEnv1, "reco-gym-v0": RecoEnv0}      # recogym/__init__.py:36-45


def make(env_id):
    """gym.make stand-in for the ids this package provides (recogym/__init__.py:42-45)."""
    try:
        return _REGISTRY[env_id]()
    except KeyError:
        raise KeyError("unknown environment id %r (recogym_b200 provides %s)" % (env_id, sorted(_REGISTRY)))
