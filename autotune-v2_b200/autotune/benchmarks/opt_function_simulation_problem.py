"""Simulated loss curves around a test function (reference benchmarks/opt_function_simulation_problem.py).

An arm's curve starts at f(x,y) + noise - start_shift, is driven towards f(x,y) - end_shift by the Gamma process of
`SimulationEvaluator`, and `evaluate(n_resources)` reads it at position int(n_resources).
"""
import numpy as np

from autotune.benchmarks.opt_function_problem import (
    OPT_FUNCTIONS, OptFunctionBuilder, OptFunctionEvaluator, OptFunctionProblem)
from autotune.core import OptimisationGoals, SimulationEvaluator, SimulationProblem


class OptFunctionSimulationEvaluator(OptFunctionEvaluator, SimulationEvaluator):
    EPSILON = 0.0001

    def __init__(self, func_name, model_builder, ml_aggressiveness=0, necessary_aggressiveness=0, up_spikiness=0,
                 is_smooth=False, start_shift=0, end_shift=200, max_resources=81, init_noise=0, should_plot=False):
        OptFunctionEvaluator.__init__(self, func_name=func_name, model_builder=model_builder)
        SimulationEvaluator.__init__(self, ml_aggressiveness, necessary_aggressiveness, up_spikiness, max_resources,
                                     is_smooth)
        self.init_noise, self.start_shift, self.end_shift = init_noise, start_shift, end_shift
        self.should_plot = should_plot

    def evaluate(self, n_resources):
        """Loss of the simulated curve after int(n_resources) resources (opt_function_simulation_problem.py:50-86)."""
        time = int(n_resources)
        n = self.max_resources
        func = OPT_FUNCTIONS[self.func_name]
        value = func(self.arm.x, self.arm.y, noise_variance=0, scaled_noise=True)
        noisy_value = func(self.arm.x, self.arm.y, self.init_noise, scaled_noise=False)
        f_n = value - self.end_shift
        f_1 = noisy_value - self.start_shift
        if not self.non_smooth_fs or time == 1:     # first read, or a read at time 1: (re)simulate the whole curve
            self.non_smooth_fs = [f_1]
            self.simulate(n, n, f_n)
        if time == self.max_resources and self.necessary_aggressiveness != np.inf:
            assert self.non_smooth_fs[n - 1] - f_n < self.EPSILON
        return OptimisationGoals(fval=self.fs[time - 1], test_error=-1, validation_error=-1)


class OptFunctionSimulationProblem(OptFunctionProblem, SimulationProblem):

    def get_evaluator(self, arm=None, ml_aggressiveness=0.9, necessary_aggressiveness=10, up_spikiness=0.1,
                      is_smooth=False, start_shift=0, end_shift=200, max_resources=81, init_noise=0,
                      should_plot=False):
        """Positional order matches `EvaluatorParams`, so `get_evaluator(*scheduler.get_family())` works."""
        return OptFunctionSimulationEvaluator(
            func_name=self.func_name, model_builder=OptFunctionBuilder(self._arm_or_random(arm)),
            ml_aggressiveness=ml_aggressiveness, necessary_aggressiveness=necessary_aggressiveness,
            up_spikiness=up_spikiness, is_smooth=is_smooth, start_shift=start_shift, end_shift=end_shift,
            max_resources=max_resources, init_noise=init_noise, should_plot=should_plot)
