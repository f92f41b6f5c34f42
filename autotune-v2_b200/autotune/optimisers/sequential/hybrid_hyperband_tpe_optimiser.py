"""Hyperband whose first round of every bracket is a TPE run (reference
optimisers/sequential/hybrid_hyperband_tpe_optimiser.py; adapted there from https://arxiv.org/pdf/1801.01596.pdf).

Simulation problems only: the TPE phase, the simulation and the halving run fused on the device
(torch.ops.at2.tpe_sim).  TPE on real-ML problems needs hyperopt, which is not a dependency of this package.
"""
from autotune.core import Optimiser, optimisation_metric_user
from autotune.optimisers.sequential.hyperband_optimiser import HyperbandOptimiser


class HybridHyperbandTpeOptimiser(HyperbandOptimiser):
    _tpe_transfer = 'none'

    def __init__(self, eta, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 plot_simulation=False):
        super().__init__(eta, max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler,
                         plot_simulation)

    @optimisation_metric_user
    def run_optimisation(self, problem, verbosity=False):
        if not self.is_simulation:
            raise NotImplementedError('Hyperband-TPE hybrids run on simulation problems only (TPE on real problems '
                                      'would need the hyperopt package)')
        return self._run_on_device(problem)
