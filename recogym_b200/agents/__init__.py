from .abstract import Agent
from .random_agent import RandomAgent, random_args
from .organic_count import OrganicCount, organic_count_args
from .bandit_count import BanditCount, bandit_count_args
