from .abstract import (Agent, Model, ModelBuilder, FeatureProvider, ViewsFeaturesProvider, ModelBasedAgent)
from .random_agent import RandomAgent, random_args
from .organic_count import OrganicCount, organic_count_args
from .bandit_count import BanditCount, bandit_count_args
from .organic_user_count import OrganicUserEventCounterAgent, organic_user_count_args
from .bandit_mf import BanditMFSquare, bandit_mf_square_args
from .epsilon_greedy import EpsilonGreedy, epsilon_greedy_args
