"""Import stub for the reference's text normalizer dependency (never called on the hot path)."""


class engine:
    def __getattr__(self, name):
        raise RuntimeError("inflect is not available in this environment")
