"""Event clocks (reference: recogym/envs/features/time/*.py)."""
import numpy as np
from numpy.random.mtrand import RandomState


class TimeGenerator:
    def __init__(self, config):
        self.config = config

    def new_time(self):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError


class DefaultTimeGenerator(TimeGenerator):
    """t = 0, 1, 2, ... per Markov event (default_time_generator.py:4-16).  This is the clock the
    CUDA kernels implement: event index == current_time."""

    def __init__(self, config):
        super().__init__(config)
        self.current_time = 0

    def new_time(self):
        now = self.current_time
        self.current_time += 1
        return now

    def reset(self):
        self.current_time = 0


class NormalTimeGenerator(TimeGenerator):
    """t += |N(mu, sigma)| (normal_time_generator.py:7-31).  As the env's `time_generator` it selects the kernels'
    normal clock (rg_env_desc.time_generator = 1): the steps are then drawn from the env's counter RNG (Philox slot 2
    keyed by user and event index), not from this object's RandomState -- `new_time()` here only serves code that
    uses the generator on its own."""

    def __init__(self, config):
        super().__init__(config)
        self.current_time = 0
        self.normal_time_mu = getattr(config, "normal_time_mu", 0)
        self.normal_time_sigma = getattr(config, "normal_time_sigma", 1)
        self.rng = RandomState(config.random_seed)

    def new_time(self):
        now = self.current_time
        self.current_time += np.abs(self.rng.normal(self.normal_time_mu, self.normal_time_sigma))
        return now

    def reset(self):
        self.current_time = 0
