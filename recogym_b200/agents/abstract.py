class Agent:
    """Agent protocol of the reference (recogym/agents/abstract.py:9-34): `act` returns at least
    {'t','u'}; subclasses add 'a', 'ps', 'ps-a'.  The env reads only action['a'] (abstract.py:236)."""

    def __init__(self, config):
        self.config = config

    def act(self, observation, reward, done):
        return {"t": observation.context().time(), "u": observation.context().user()}

    def train(self, observation, action, reward, done=False):
        pass

    def reset(self):
        pass
