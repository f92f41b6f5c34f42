"""Hyperparameter domains of the reference's recorded-arm databases and synthetic databases of the same shape.

The real MNIST / CIFAR problems (training runs, datasets) are out of scope; the closest-known-function path only needs
their *domains* (benchmarks/mnist_problem.py:11-19, benchmarks/cifar_problem.py:11-33) to normalise arms, so they are
restated here together with `RecordedDomainProblem`, the stand-in for `MnistProblem(...)` / `CifarProblem(...)` that
experiments/run_closest_loss_fn_approximation.py passes to `KnownFnProblem` as `real_problem`.

Synthetic recorded-arm database of the MNIST shape: the reference's real database
(`loss_functions/`, experiments/run_closest_loss_fn_approximation.py:93-129) is git-ignored upstream, so benchmarks and
examples use A arms drawn from the MNIST hyperparameter domain (benchmarks/mnist_problem.py:11-19) with curves
c_a(t) = b_a + (1 - b_a) exp(-t / tau_a), b ~ U[0.07, 0.57], tau ~ U[5, 55]."""
import numpy as np

from autotune.core import Arm, Domain, HyperparameterOptimisationProblem, Param

MNIST_DOMAIN = Domain(
    learning_rate=Param('learning_rate', np.log(10 ** -6), np.log(10 ** 0), distrib='uniform', scale='log'),
    weight_decay=Param('weight_decay', np.log(10 ** -6), np.log(10 ** -1), distrib='uniform', scale='log'),
    momentum=Param('momentum', 0.3, 0.999, distrib='uniform', scale='linear'),
    batch_size=Param('batch_size', 20, 2000, distrib='uniform', scale='linear', interval=1))


def synthetic_mnist_database(n_arms=425, curve_len=300, seed=0):
    """-> (known_fs {Arm: [loss]}, domain, hyperparams_to_opt): the arguments of autotune.b200.engine.KnownFnSpec."""
    state = np.random.get_state()
    np.random.seed(seed)
    names = tuple(MNIST_DOMAIN.hyperparams_names())
    arms = []
    for _ in range(n_arms):
        arm = Arm()
        arm.draw_hp_val(domain=MNIST_DOMAIN, hyperparams_to_opt=names)
        arms.append(arm)
    np.random.set_state(state)
    rng = np.random.default_rng(seed)
    b, tau = rng.uniform(0.07, 0.57, n_arms), rng.uniform(5, 55, n_arms)
    t = np.arange(1, curve_len + 1)
    curves = b[:, None] + (1 - b[:, None]) * np.exp(-t[None, :] / tau[:, None])
    return {a: list(c) for a, c in zip(arms, curves)}, MNIST_DOMAIN, names


_UNITS = dict(min_val=np.log(2 ** 4), max_val=np.log(2 ** 8), distrib='uniform', scale='log', interval=1)
CIFAR_DOMAIN = Domain(   # benchmarks/cifar_problem.py:11-31
    learning_rate=Param('learning_rate', np.log(10 ** -6), np.log(10 ** 0), distrib='uniform', scale='log'),
    n_units_1=Param('n_units_1', **_UNITS), n_units_2=Param('n_units_2', **_UNITS), n_units_3=Param('n_units_3', **_UNITS),
    batch_size=Param('batch_size', 32, 512, distrib='uniform', scale='linear', interval=1),
    lr_step=Param('lr_step', 1, 5, distrib='uniform', init_val=1, scale='linear', interval=1),
    gamma=Param('gamma', np.log(10 ** -3), np.log(10 ** -1), distrib='uniform', init_val=0.1, scale='log'),
    weight_decay=Param('weight_decay', np.log(10 ** -6), np.log(10 ** -1), init_val=0.004, distrib='uniform', scale='log'),
    momentum=Param('momentum', 0.3, 0.999, init_val=0.9, distrib='uniform', scale='linear'))
CIFAR_HYPERPARAMS_TO_OPTIMIZE = ('learning_rate', 'n_units_1', 'n_units_2', 'n_units_3', 'batch_size')   # cifar_problem.py:33


class RecordedDomainProblem(HyperparameterOptimisationProblem):
    """Carries the domain (and the optimised subset) of a real problem whose recorded arms are being replayed."""

    def __init__(self, name, domain, hyperparams_to_opt=()):
        super().__init__(domain, hyperparams_to_opt)
        self.name = name

    def get_evaluator(self, arm=None):
        raise NotImplementedError(f'{self.name}: the real training problem is not part of this package; wrap it in '
                                  'KnownFnProblem to replay recorded arms')


def mnist_domain_problem():
    from autotune.benchmarks.mnist_problem import MnistProblem
    return MnistProblem()


def cifar_domain_problem():
    from autotune.benchmarks.cifar_problem import CifarProblem
    return CifarProblem()


def synthetic_database(problem, n_arms, curve_len=300, seed=0):
    """{Arm: [loss]} of `n_arms` arms drawn from `problem`'s domain with the curve family of the module docstring."""
    state = np.random.get_state()
    np.random.seed(seed)
    arms = []
    for _ in range(n_arms):
        arm = Arm()
        arm.draw_hp_val(domain=problem.domain, hyperparams_to_opt=problem.hyperparams_to_opt)
        arms.append(arm)
    np.random.set_state(state)
    rng = np.random.default_rng(seed)
    b, tau = rng.uniform(0.07, 0.57, n_arms), rng.uniform(5, 55, n_arms)
    t = np.arange(1, curve_len + 1)
    curves = b[:, None] + (1 - b[:, None]) * np.exp(-t[None, :] / tau[:, None])
    return {a: list(c) for a, c in zip(arms, curves)}
