class Configuration:
    """Immutable attribute bag built from a dict (reference: recogym/envs/configuration.py:1-24).

    Values may be any object. Assignment after construction is silently ignored and deep copies
    return the same instance, exactly as in the reference, because drivers deepcopy envs and
    agents freely and expect the configuration to be shared.
    """

    def __init__(self, args):
        for key, value in dict(args).items():
            object.__setattr__(self, key, value)

    def __setattr__(self, key, value):
        return None

    def __deepcopy__(self, memo=None):
        return self

    def __reduce__(self):
        return (Configuration, (dict(self.__dict__),))
