"""Oracle as the timed CPU baseline (bench.py's `cpu_baseline` leg and `--impl reference` arm only).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Times `oracle.hyperband.run_global_rng`, the restatement that follows
the reference's own per-arm call pattern (lazy simulation, np.random.gamma per epoch, re-filter per read,
experiments/run_simulation.py:147-157) and is pinned bit-for-bit to the reference under seeded RNG
(tests/golden/seeded_ofes.npz) - i.e. it does the work the reference's CPU path does.
"""
import multiprocessing as mp
import os
import random
import time

import numpy as np

from . import hyperband as ohb


def _worker(args):
    func, families, R, eta, noise, n_reps, seed = args
    np.random.seed(seed)
    random.seed(seed)
    t0 = time.perf_counter()
    ofes = [ohb.run_global_rng(func, families, R, eta, noise) for _ in range(n_reps)]
    return time.perf_counter() - t0, ofes


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class HyperbandCpuPool:
    """A fork pool that is created BEFORE any CUDA initialisation and reused for every timed sample."""

    def __init__(self, procs=None):
        self.procs = procs or host_cores()
        self.pool = mp.get_context('fork').Pool(self.procs) if self.procs > 1 else None

    def run(self, func, families, R, eta, noise, reps_per_proc, seed0=1000):
        """Wall-clock throughput (runs/s) of procs x reps_per_proc repetitions; returns (runs_per_s, seconds, ofes)."""
        jobs = [(func, families, R, eta, noise, reps_per_proc, seed0 + w) for w in range(self.procs)]
        t0 = time.perf_counter()
        res = self.pool.map(_worker, jobs) if self.pool else [_worker(jobs[0])]
        wall = time.perf_counter() - t0
        ofes = [o for _, part in res for o in part]
        return len(ofes) / wall, wall, ofes

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
