class Observation:
    """What an agent sees: a context and the organic sessions since the last action
    (reference: recogym/envs/observation.py:1-10; agents also read .current_sessions directly,
    recogym/agents/organic_mf.py:60-61)."""

    def __init__(self, context, sessions):
        self.current_context = context
        self.current_sessions = sessions

    def context(self):
        return self.current_context

    def sessions(self):
        return self.current_sessions
