class Session(list):
    """reference: recogym/envs/session.py:1-15"""

    def get_type(self):
        raise NotImplementedError

    def to_strings(self, user_id, session_id):
        kind = self.get_type()
        return [",".join([str(user_id), kind, str(session_id), event, str(product)])
                for event, product in self]


class OrganicSessions(Session):
    """List of {'t','u','z':'pageview','v'} dicts (reference: recogym/envs/session.py:18-36)."""

    def next(self, context, product):
        self.append({"t": context.time(), "u": context.user(), "z": "pageview", "v": product})

    def get_type(self):
        return "organic"

    def get_views(self):
        return [p for _, _, e, p in self if e == "pageview"]
