"""CIFAR-10 / SVHN / MRBI problems - DOMAIN ONLY (reference benchmarks/cifar_problem.py:11-33,92-120, svhn_problem.py,
mrbi_problem.py).  See mnist_problem.py: the closest-known-loss-function path needs the domain, nothing is trained."""
from autotune.benchmarks.mnist_like import CIFAR_DOMAIN, CIFAR_HYPERPARAMS_TO_OPTIMIZE, RecordedDomainProblem

HYPERPARAMS_DOMAIN = CIFAR_DOMAIN
HYPERPARAMETERS_TO_OPTIMIZE = CIFAR_HYPERPARAMS_TO_OPTIMIZE


class CifarProblem(RecordedDomainProblem):
    _name = 'cifar'

    def __init__(self, data_dir=None, output_dir=None, dataset_loader=None, hyperparams_domain=HYPERPARAMS_DOMAIN,
                 hyperparams_to_opt=HYPERPARAMETERS_TO_OPTIMIZE, in_channels=3):
        super().__init__(self._name, hyperparams_domain, hyperparams_to_opt)
        # the reference intersects sets, which leaves the order unspecified; arms are drawn in domain order
        self.hyperparams_to_opt = tuple(n for n in hyperparams_domain.hyperparams_names() if n in self.hyperparams_to_opt)
        self.data_dir, self.output_dir = data_dir, output_dir


class SvhnProblem(CifarProblem):     # reference svhn_problem.py:8 - same domain, SVHN data
    _name = 'svhn'


class MrbiProblem(CifarProblem):     # reference mrbi_problem.py:8 - same domain, one input channel
    _name = 'mrbi'
