"""Hyperband-TPE hybrids with history transfer between brackets (reference
optimisers/sequential/hybrid_hyperband_tpe_with_transfer.py).

Before a bracket's TPE run, evaluations of earlier brackets are injected as TPE history when the resource level r at
which the arm was LAST evaluated satisfies the variant's predicate against the new bracket's first level r_b:
  NoTransfer: never          TransferLongest: r >= max_iter       TransferAll: r >= r_b
  TransferSame: r == r_b     TransferThreshold: r >= threshold and r >= r_b
"""
from abc import abstractmethod

from autotune.core import Optimiser
from autotune.optimisers.sequential.hybrid_hyperband_tpe_optimiser import HybridHyperbandTpeOptimiser


class HybridHyperbandTpeWithTransferOptimiser(HybridHyperbandTpeOptimiser):

    def __init__(self, eta, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 plot_simulation=False):
        super().__init__(eta, max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler,
                         plot_simulation)
        self.evaluations_by_resources = {}

    @abstractmethod
    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        pass


class HybridHyperbandTpeNoTransferOptimiser(HybridHyperbandTpeWithTransferOptimiser):
    _tpe_transfer = 'none'

    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        return False


class HybridHyperbandTpeTransferLongestOptimiser(HybridHyperbandTpeWithTransferOptimiser):
    _tpe_transfer = 'longest'

    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        return evaluator_n_evaluated_resources >= self.max_iter


class HybridHyperbandTpeTransferAllOptimiser(HybridHyperbandTpeWithTransferOptimiser):
    _tpe_transfer = 'all'

    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        return evaluator_n_evaluated_resources >= bracket_n_resources


class HybridHyperbandTpeTransferSameOptimiser(HybridHyperbandTpeWithTransferOptimiser):
    _tpe_transfer = 'same'

    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        return evaluator_n_evaluated_resources == bracket_n_resources


class HybridHyperbandTpeTransferThresholdOptimiser(HybridHyperbandTpeWithTransferOptimiser):
    _tpe_transfer = 'threshold'

    def __init__(self, eta, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 n_res_transfer_threshold=None):
        super().__init__(eta, max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler)
        self.n_res_transfer_threshold = n_res_transfer_threshold

    def _is_transferable(self, evaluator_n_evaluated_resources, bracket_n_resources):
        return (evaluator_n_evaluated_resources >= self.n_res_transfer_threshold and
                evaluator_n_evaluated_resources >= bracket_n_resources)
