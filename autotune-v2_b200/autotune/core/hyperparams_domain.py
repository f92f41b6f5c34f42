"""Domain: named hyperparameter ranges (reference core/hyperparams_domain.py:13-38)."""
from types import SimpleNamespace


class Domain(SimpleNamespace):
    """`Domain(x=Param(...), y=Param(...))`; attribute and item access, insertion order = draw order."""

    def __getitem__(self, name):
        return getattr(self, name)

    def hyperparams_names(self):
        return self.__dict__.keys()
