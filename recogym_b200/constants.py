"""Enums keyed into the result dicts of the drivers (reference: recogym/constants.py:4-61)."""
from enum import Enum


class AgentStats(Enum):
    Q0_025 = 0
    Q0_500 = 1
    Q0_975 = 2
    SAMPLES = 3
    AGENTS = 4
    SUCCESSES = 5
    FAILURES = 6


class AgentInit(Enum):
    CTOR = 0
    DEF_ARGS = 1


class TrainingApproach(Enum):
    ALL_DATA = 0
    SLIDING_WINDOW_ALL_DATA = 1
    ALL_EXPLORATION_DATA = 2
    SLIDING_WINDOW_EXPLORATION_DATA = 3
    MOST_VALUABLE = 4
    LAST_STEP = 5


class EvolutionCase(Enum):
    SUCCESS = 0
    FAILURE = 1
    ACTIONS = 2
    SUCCESS_GREEDY = 3
    FAILURE_GREEDY = 4


class RoiMetrics(Enum):
    ROI_MEAN = 0
    ROI_0_025 = 1
    ROI_0_975 = 2
    ROI_SUCCESS = 3
    ROI_FAILURE = 4
