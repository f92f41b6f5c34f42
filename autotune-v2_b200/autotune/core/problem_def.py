"""Problem definitions (reference core/problem_def.py:18-105)."""
from abc import abstractmethod
from pprint import PrettyPrinter

import numpy as np


class HyperparameterOptimisationProblem:
    """A hyperparameter domain + which of its hyperparameters are optimised; subclasses hand out evaluators."""

    def __init__(self, hyperparams_domain, hyperparams_to_opt=(), dataset_loader=None, output_dir=None):
        self.domain = hyperparams_domain
        names = tuple(self.domain.hyperparams_names())
        if hyperparams_to_opt:
            self.hyperparams_to_opt = tuple(set(hyperparams_to_opt) & set(names))
        else:                                     # () means: optimise everything in the domain
            self.hyperparams_to_opt = names
        print(f"\n> Hyperparameters to optimise:\n    {'' if hyperparams_to_opt else 'ALL:'} {self.hyperparams_to_opt}")
        self.dataset_loader, self.output_dir = dataset_loader, output_dir

    def log_domain(self):
        print(f"\n> Problem {type(self).__name__} hyperparameters domain:")
        print(PrettyPrinter(indent=4).pformat(self.domain.__dict__))

    @abstractmethod
    def get_evaluator(self, arm=None):
        """Evaluator for `arm` (a random arm when None)."""

    def get_hyperopt_space_from_hyperparams_to_opt(self):
        """{name: (kind, name, low, high[, q])} with hyperopt's kind names (problem_def.py:59-78).  hyperopt itself is
        not a dependency here: the TPE sampler in autotune.b200 consumes this description directly."""
        space = {}
        for name in self.hyperparams_to_opt:
            prm = self.domain[name]
            if prm.scale == 'log':
                assert prm.logbase == np.e
                kind = 'qloguniform' if prm.interval else 'loguniform'
            else:
                kind = 'quniform' if prm.interval else 'uniform'
            space[name] = (kind, prm.name, prm.min_val, prm.max_val) + ((prm.interval,) if prm.interval else ())
        return space


class SimulationProblem:
    """Marker base: problems whose evaluators are instant simulations (problem_def.py:101-105)."""

    @abstractmethod
    def get_evaluator(self, arm=None, should_plot=True):
        pass
