"""Drop-in for the `autotune` package of cristianmatache/autotune-v2, restricted to its Monte-Carlo
optimiser-evaluation path (simulation problems, closest-known-loss-function problem, Hyperband / TPE / hybrids,
EPDF-OFE), with the arithmetic executed by hand-written sm_100a CUDA kernels (see `autotune.b200`).

Same class names, constructor arguments and `run_optimisation` / `evaluate` surface as the reference
(autotune/core, autotune/benchmarks, autotune/optimisers); there is no CPU fallback.
"""
__version__ = '0.1.0'
