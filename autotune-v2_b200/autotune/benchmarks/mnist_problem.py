"""MNIST problem - DOMAIN ONLY (reference benchmarks/mnist_problem.py:11-19,62-75).

Training logistic regression on MNIST is outside the hot path this package rebuilds (real-ML problems are out
of scope).  What the closest-known-loss-function path needs from `MnistProblem` is its hyperparameter domain, which
normalises proposed and recorded arms (experiments/run_closest_loss_fn_approximation.py:80,186).  The class keeps the
reference's constructor signature so `KnownFnProblem(known_fs, MnistProblem(input_dir, output_dir))` reads as it does
there; `get_evaluator` raises."""
from autotune.benchmarks.mnist_like import MNIST_DOMAIN, RecordedDomainProblem

HYPERPARAMS_DOMAIN = MNIST_DOMAIN


class MnistProblem(RecordedDomainProblem):

    def __init__(self, data_dir=None, output_dir=None, hyperparams_domain=HYPERPARAMS_DOMAIN, hyperparams_to_opt=()):
        super().__init__('mnist', hyperparams_domain, hyperparams_to_opt)
        self.data_dir, self.output_dir = data_dir, output_dir
