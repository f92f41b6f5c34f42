"""OptimisationGoals: the metrics an evaluation reports (reference core/optimisation_goals.py:6-48)."""
from types import SimpleNamespace


class OptimisationGoals(SimpleNamespace):
    """`OptimisationGoals(fval=..., test_error=..., validation_error=...)`; optimisers min/max a function of it."""

    def __init__(self, **kwargs):
        if 'validation_error' not in kwargs or 'test_error' not in kwargs:
            print("WARNING: validation_error or test_error or both are not included in the Optimization Goals")
        super().__init__(**kwargs)

    def goals_to_str(self, goals_to_print=()):
        names = goals_to_print or tuple(self.__dict__)
        goals = {g: self.__dict__[g] for g in names if g in self.__dict__}
        if not goals:
            return ""
        width = max(len(g) for g in goals) + 1
        return "\n".join(["> Optimization goals:"] + [f"  {g}:{' ' * (width - len(g))}{v}" for g, v in goals.items()])

    def __str__(self):
        return self.goals_to_str()
