"""Synthetic recorded-arm database of the MNIST shape: the reference's real database
(`loss_functions/`, experiments/run_closest_loss_fn_approximation.py:93-129) is git-ignored upstream, so benchmarks and
examples use A arms drawn from the MNIST hyperparameter domain (benchmarks/mnist_problem.py:11-19) with curves
c_a(t) = b_a + (1 - b_a) exp(-t / tau_a), b ~ U[0.07, 0.57], tau ~ U[5, 55]."""
import numpy as np

from autotune.core import Arm, Domain, Param

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
