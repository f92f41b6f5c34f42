"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

CPU restatement (plain Python floats / numpy, float64) of the reference's Monte-Carlo optimiser-evaluation path, written
from the behaviour documented in SURVEY.md section 8a; every function cites the reference file:line it follows.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this
package - as the checker or as the timed CPU baseline, never as something the product path falls back to.

Parity pins (see tests/test_oracle_golden.py):
  * Hyperband + Gamma-process simulation: pinned bit-for-bit against outputs of the reference itself, both on
    pre-drawn inputs (tests/golden/hb_*.npz) and on seeded global-RNG runs (tests/golden/seeded_ofes.npz).
  * Epanechnikov KDE: pinned against sklearn's KernelDensity (the reference's call, plot_results.py:11-17).
  * Closest-known-loss-function lookup: pinned against the reference's KnownFnEvaluator (tests/golden/knownfn_*.npz).
  * TPE (hyperopt 0.2.4, third-party, absent from /root/reference and from this image): PARITY UNPINNED at the
    arithmetic level - see oracle/tpe.py.
"""
