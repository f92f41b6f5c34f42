"""SigOpt optimisers - NOT PART OF THIS PACKAGE (reference optimisers/sequential/sigopt_optimiser.py,
hybrid_hyperband_sigopt_optimiser.py).  They are clients of a remote proprietary service; the names exist so that the
reference's import lines keep working, constructing one raises."""
from autotune.optimisers.sequential.hyperband_optimiser import HyperbandOptimiser
from autotune.core import Optimiser


class SigOptimiser(Optimiser):

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('SigOpt is a remote proprietary service; SigOptimiser is not part of this package')


class HybridHyperbandSigoptOptimiser(HyperbandOptimiser):

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('SigOpt is a remote proprietary service; HybridHyperbandSigoptOptimiser is not part of '
                                  'this package')
