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


class Model:
    """Acting half of a model-based agent (reference: recogym/agents/abstract.py:114-130)."""

    def __init__(self, config):
        self.config = config

    def act(self, observation, features):
        return {"t": observation.context().time(), "u": observation.context().user()}

    def reset(self):
        pass


class FeatureProvider:
    """Turns the observations of ONE user into the feature set a Model acts on
    (reference: recogym/agents/abstract.py:133-170)."""

    def __init__(self, config):
        self.config = config

    def observe(self, observation):
        raise NotImplementedError

    def features(self, observation):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError


class ModelBuilder:
    """Collects (observation, action, reward) rows until `build()` turns them into a
    (FeatureProvider, Model) pair (reference: recogym/agents/abstract.py:37-111)."""

    def __init__(self, config):
        self.config = config
        self.data = None
        self.reset()

    def train(self, observation, action, reward, done):
        assert observation is not None and observation.sessions() is not None
        d = self.data
        for s in observation.sessions():
            d["t"].append(s["t"]); d["u"].append(s["u"]); d["z"].append("organic"); d["v"].append(s["v"])
            d["a"].append(None); d["c"].append(None); d["ps"].append(None)
        if action:
            d["t"].append(action["t"]); d["u"].append(action["u"]); d["z"].append("bandit"); d["v"].append(None)
            d["a"].append(action["a"]); d["c"].append(reward); d["ps"].append(action["ps"])

    def build(self):
        raise NotImplementedError

    def reset(self):
        self.data = {k: [] for k in ("t", "u", "z", "v", "a", "c", "ps")}


class ViewsFeaturesProvider(FeatureProvider):
    """Per-user product view counts, shape (1, P) (reference: recogym/agents/abstract.py:318-409).
    Without `weight_history_function` the features are the raw int16 counts; with it every view is
    weighted by f(now - time of the view) and summed in float32."""

    def __init__(self, config, is_sparse=False):
        super().__init__(config)
        self.is_sparse = is_sparse
        self.with_history = getattr(self.config, "weight_history_function", None) is not None
        self.reset()

    def observe(self, observation):
        assert observation is not None and observation.sessions() is not None
        for s in observation.sessions():
            v = int(s["v"])
            if self.with_history:
                self._viewed.append(v)
                self._when.append(int(s["t"]))
            else:
                self.views[0, v] += 1

    def features(self, observation):
        if not self.with_history:
            return self.views
        import numpy as np
        now = int(observation.context().time())
        w = np.asarray(self.config.weight_history_function(now - np.array(self._when)), dtype=np.float32)
        out = np.zeros((1, self.config.num_products), dtype=np.float32)
        np.add.at(out[0], np.array(self._viewed, dtype=np.int64), w)
        return out

    def reset(self):
        import numpy as np
        self._viewed, self._when = [], []
        self.views = np.zeros((1, self.config.num_products), dtype=np.int16)


class ModelBasedAgent(Agent):
    """Agent = ModelBuilder (training) + FeatureProvider/Model (acting), built lazily on the first
    `act` (reference: recogym/agents/abstract.py:283-315)."""

    def __init__(self, config, model_builder):
        super().__init__(config)
        self.model_builder = model_builder
        self.feature_provider = None
        self.model = None

    def train(self, observation, action, reward, done=False):
        self.model_builder.train(observation, action, reward, done)

    def act(self, observation, reward, done):
        if self.model is None:
            assert self.feature_provider is None
            self.feature_provider, self.model = self.model_builder.build()
        self.feature_provider.observe(observation)
        return {**super().act(observation, reward, done),
                **self.model.act(observation, self.feature_provider.features(observation))}

    def reset(self):
        if self.model is not None:
            assert self.feature_provider is not None
            self.feature_provider.reset()
            self.model.reset()
