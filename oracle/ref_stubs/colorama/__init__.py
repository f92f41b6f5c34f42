"""Presentation-only stand-in so the reference imports in this container (no arithmetic)."""


class _Blank:
    def __getattr__(self, name):
        return ''


Fore = Style = Back = _Blank()


def init(*a, **k):
    pass
