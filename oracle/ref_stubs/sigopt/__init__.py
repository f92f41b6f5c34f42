from . import objects  # noqa: F401


class Connection:
    def __init__(self, *a, **k):
        raise NotImplementedError('sigopt is a remote service; not available')
