"""Optimiser base class (reference core/optimiser.py:16-151): stop conditions, evaluation history, wall-clock metrics."""
import time
from abc import abstractmethod

import numpy as np

from autotune.core.evaluation import Evaluation
from autotune.core.problem_def import SimulationProblem


class Optimiser:
    """Results go into `eval_history`; the optimum is min/max of `optimisation_func` over it (first optimum wins)."""
    checkpoints = []

    @staticmethod
    def default_optimisation_func(opt_goals):
        """validation_error (Default optimisation_func)"""
        return opt_goals.validation_error

    def __init__(self, max_iter=None, max_time=None, min_or_max=min, optimisation_func=default_optimisation_func,
                 is_simulation=False, scheduler=None, plot_simulation=False):
        if max_iter is None and max_time is None:
            raise ValueError("max_iter and max_time cannot be None simultaneously")
        self.max_time = np.inf if max_time is None else max_time
        self.max_iter = np.inf if max_iter is None else max_iter
        if min_or_max not in [min, max]:
            raise ValueError(f"optimization must be a built in function: min or max, instead {min_or_max} was supplied")
        self.min_or_max, self.optimisation_func = min_or_max, optimisation_func
        self.eval_history = []
        self.is_simulation, self.scheduler, self.plot_simulation = is_simulation, scheduler, plot_simulation

    def _update_evaluation_history(self, evaluator, opt_goals):
        self.eval_history.append(Evaluation(evaluator, opt_goals))

    def _get_best_evaluation(self):
        return self.min_or_max(self.eval_history, key=lambda e: self.optimisation_func(e.optimisation_goals))

    @abstractmethod
    def run_optimisation(self, problem, verbosity):
        """Run to the stop condition and return the Evaluation of the best arm."""

    def init_optimiser_metrics(self):
        self.time_zero = time.time()
        self.cum_time = 0
        self.num_iterations = 0
        self.checkpoints = []

    def _update_optimiser_metrics(self):
        self.cum_time = time.time() - self.time_zero
        self.num_iterations += 1
        self.checkpoints.append(self.cum_time)

    def _needs_to_stop(self):
        timed_out, spent = self.max_time <= self.cum_time, self.max_iter <= self.num_iterations
        if timed_out:
            print("\nFINISHED: Time limit exceeded")
        elif spent:
            print("\nFINISHED: Exceeded maximum number of iterations")
        return timed_out or spent

    def __str__(self):
        return (f"\n> Starting optimisation\n  Stop when:\n    Max iterations          = {self.max_iter}\n"
                f"    Max time                = {self.max_time} seconds\n"
                f"  Optimizing ({self.min_or_max.__name__}) {self.optimisation_func.__name__}"
                f" {self.optimisation_func.__doc__}")

    def _print_evaluation(self, opt_func_value):
        best = self.min_or_max([self.optimisation_func(e.optimisation_goals) for e in self.eval_history])
        gap = 8 * ' '
        print(f"\n> SUMMARY: iteration number: {self.num_iterations},{gap}time elapsed: {self.cum_time:.2f}s,{gap}"
              f"current {self.optimisation_func.__doc__}: {opt_func_value:.5f},{gap}"
              f" best {self.optimisation_func.__doc__} so far: {best:.5f}")


def optimisation_metric_user(run_optimisation):
    """Decorator for run_optimisation: resets the metrics and rejects a simulation run on a non-simulation problem."""
    def wrapper(self, problem, verbosity=False):
        self.init_optimiser_metrics()
        if self.is_simulation and not isinstance(problem, SimulationProblem):
            raise ValueError("You are trying to run a simulation but you have not provided a shape family schedule")
        return run_optimisation(self, problem, verbosity)
    wrapper.__doc__ = run_optimisation.__doc__
    return wrapper
