"""The UNMODIFIED reference as the timed CPU arm (bench.py `--impl reference` and the `cpu_baseline` leg only).

TEST / MEASUREMENT INFRASTRUCTURE - see oracle/__init__.py.  Worker processes import the reference's own `autotune`
package from baseline/_ref/ (placed there verbatim by scripts/install_reference.py; presentation-only stand-ins for
colorama / matplotlib / hyperopt / sigopt come from oracle/ref_stubs/ - none touches arithmetic on the Hyperband path)
and drive it exactly as the reference's driver does (experiments/run_simulation.py:147-157): per repetition a fresh
problem, scheduler and HyperbandOptimiser, `run_optimisation(problem, verbosity=True)`, the optimum's `fval` collected.
The parent process never imports the reference (it may hold this repo's own `autotune` package).
"""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')
_STATE = {}


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'autotune', 'optimisers'))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _init_worker():
    """Runs once in every pool process: reference on the path (ahead of anything else called `autotune`), stdout to
    /dev/null (the reference driver prints every round)."""
    for name in [m for m in sys.modules if m == 'autotune' or m.startswith('autotune.')]:
        del sys.modules[name]
    sys.path[:] = [REF_DIR, os.path.join(HERE, 'ref_stubs')] + [p for p in sys.path if 'autotune-v2_b200' not in p]
    sys.stdout = open(os.devnull, 'w')
    from autotune.benchmarks import OptFunctionSimulationProblem
    from autotune.core import RoundRobinShapeFamilyScheduler, ShapeFamily, UniformShapeFamilyScheduler
    from autotune.optimisers import HyperbandOptimiser
    import autotune
    assert os.path.realpath(autotune.__file__).startswith(os.path.realpath(REF_DIR)), autotune.__file__
    _STATE.update(Problem=OptFunctionSimulationProblem, Uniform=UniformShapeFamilyScheduler,
                  RoundRobin=RoundRobinShapeFamilyScheduler, ShapeFamily=ShapeFamily, Hyperband=HyperbandOptimiser)


def _fval(opt_goals):
    """fval."""
    return opt_goals.fval


def _worker(job):
    func, families, R, eta, noise, n_reps, seed = job
    import random

    import numpy as np
    np.random.seed(seed)
    random.seed(seed)
    S = _STATE
    fams = tuple(S['ShapeFamily'](None, *f) for f in families)
    t0 = time.perf_counter()
    ofes = []
    for _ in range(n_reps):                                   # run_simulation.py:147-157
        problem = S['Problem'](func_name=func)
        problem.log_domain()
        scheduler = S['Uniform'](shape_families=fams, max_resources=R, init_noise=noise)
        optimiser = S['Hyperband'](eta=eta, max_iter=R, max_time=None, min_or_max=min, optimisation_func=_fval,
                                   is_simulation=True, scheduler=scheduler, plot_simulation=False)
        optimum = optimiser.run_optimisation(problem, verbosity=True)
        ofes.append(_fval(optimum.optimisation_goals))
    return time.perf_counter() - t0, ofes


class ReferencePool:
    """Fork pool of reference workers; create it BEFORE any CUDA initialisation."""

    def __init__(self, procs=None):
        if not available():
            raise RuntimeError(f'{REF_DIR}/autotune is missing: run scripts/install_reference.py where /root/reference exists')
        self.procs = procs or host_cores()
        self.pool = mp.get_context('fork').Pool(self.procs, initializer=_init_worker)

    def run(self, func, families, R, eta, noise, reps_per_proc, seed0=1000):
        """(runs_per_s, wall seconds, ofes) of procs x reps_per_proc repetitions of the reference's own loop."""
        jobs = [(func, [tuple(f) for f in families], R, eta, noise, reps_per_proc, seed0 + w) for w in range(self.procs)]
        t0 = time.perf_counter()
        res = self.pool.map(_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        ofes = [o for _, part in res for o in part]
        return len(ofes) / wall, wall, ofes

    def close(self):
        self.pool.close()
        self.pool.join()
