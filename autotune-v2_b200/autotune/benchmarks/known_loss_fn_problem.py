"""Closest-known-loss-function approximation (reference benchmarks/known_loss_fn_problem.py) - filled in with kernel (d)."""
from autotune.core import HyperparameterOptimisationProblem, SimulationProblem


class KnownFnProblem(HyperparameterOptimisationProblem, SimulationProblem):

    def __init__(self, known_fs, real_problem):
        HyperparameterOptimisationProblem.__init__(self, real_problem.domain, real_problem.hyperparams_to_opt)
        self.known_fs = known_fs

    def get_evaluator(self, arm=None, should_plot=False):
        raise NotImplementedError
