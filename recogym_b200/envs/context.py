class Context:
    """reference: recogym/envs/context.py:1-7"""

    def time(self):
        raise NotImplementedError

    def user(self):
        raise NotImplementedError


class DefaultContext(Context):
    """(time, user) pair attached to observations (reference: recogym/envs/context.py:10-21)."""

    def __init__(self, current_time, current_user_id):
        self.current_time = current_time
        self.current_user_id = current_user_id

    def time(self):
        return self.current_time

    def user(self):
        return self.current_user_id
