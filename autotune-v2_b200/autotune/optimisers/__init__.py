from autotune.optimisers.sequential.hybrid_hyperband_tpe_optimiser import HybridHyperbandTpeOptimiser
from autotune.optimisers.sequential.hybrid_hyperband_tpe_with_transfer import (
    HybridHyperbandTpeNoTransferOptimiser, HybridHyperbandTpeTransferAllOptimiser,
    HybridHyperbandTpeTransferLongestOptimiser, HybridHyperbandTpeTransferSameOptimiser,
    HybridHyperbandTpeTransferThresholdOptimiser)
from autotune.optimisers.sequential.hyperband_optimiser import HyperbandOptimiser
from autotune.optimisers.sequential.random_optimiser import RandomOptimiser
from autotune.optimisers.sequential.sigopt_optimiser import HybridHyperbandSigoptOptimiser, SigOptimiser
from autotune.optimisers.sequential.tpe_optimiser import TpeOptimiser

__all__ = [
    'HybridHyperbandTpeOptimiser', 'HybridHyperbandSigoptOptimiser', 'HyperbandOptimiser', 'RandomOptimiser',
    'SigOptimiser', 'TpeOptimiser',
    'HybridHyperbandTpeTransferAllOptimiser', 'HybridHyperbandTpeTransferLongestOptimiser',
    'HybridHyperbandTpeTransferThresholdOptimiser', 'HybridHyperbandTpeNoTransferOptimiser',
    'HybridHyperbandTpeTransferSameOptimiser']
