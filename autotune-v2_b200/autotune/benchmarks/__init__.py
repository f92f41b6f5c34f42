from autotune.benchmarks.cifar_problem import CifarProblem, MrbiProblem, SvhnProblem
from autotune.benchmarks.known_loss_fn_problem import KnownFnProblem
from autotune.benchmarks.mnist_problem import MnistProblem
from autotune.benchmarks.opt_function_problem import AVAILABLE_OPT_FUNCTIONS, OptFunctionProblem
from autotune.benchmarks.opt_function_simulation_problem import OptFunctionSimulationProblem

# CifarProblem / MnistProblem / MrbiProblem / SvhnProblem carry their hyperparameter DOMAINS only (what the
# closest-known-loss-function path needs); training the real models is outside this package's scope.
__all__ = ['CifarProblem', 'MnistProblem', 'MrbiProblem', 'SvhnProblem',
           'OptFunctionSimulationProblem', 'OptFunctionProblem', 'AVAILABLE_OPT_FUNCTIONS', 'KnownFnProblem']
