"""No-op pyplot: every attribute is a callable that does nothing."""


class Axes:
    pass


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
