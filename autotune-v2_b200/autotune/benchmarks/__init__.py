from autotune.benchmarks.known_loss_fn_problem import KnownFnProblem
from autotune.benchmarks.opt_function_problem import AVAILABLE_OPT_FUNCTIONS, OptFunctionProblem
from autotune.benchmarks.opt_function_simulation_problem import OptFunctionSimulationProblem

__all__ = ['OptFunctionSimulationProblem', 'OptFunctionProblem', 'AVAILABLE_OPT_FUNCTIONS', 'KnownFnProblem']
