"""Presentation-only stand-in so the reference imports in this container (no arithmetic)."""


def use(*a, **k):
    pass
