"""Oracle, vectorised: the same Hyperband-on-simulation repetitions as oracle/hyperband.py, written the way a NumPy user
would batch them - all arms of a block of repetitions advance together, one vectorised Gamma draw per epoch, each arm
simulated once, the Savitzky-Golay read-outs as rows of the filter's linear operator, stable argsort for the halving.

TEST INFRASTRUCTURE - see oracle/__init__.py.  This is NOT the reference's code shape (the reference loops over python
objects, experiments/run_simulation.py:147-157); BASELINE.md asks for such a figure next to the reference-shaped one "so
that the GPU speed-up is not overstated".  It uses its own Generator, so it is pinned distributionally only: two-sample KS
against the reference's stored OFE vectors (tests/test_oracle_golden.py).  Reference behaviour followed:
core/simulation_evaluator.py:36-48,84-116 (Gamma draw, recurrence), :76-82 (read-out),
benchmarks/opt_function_simulation_problem.py:61-64 (f_1, f_n), optimisers/sequential/hyperband_optimiser.py:46-57,89-106.
"""
import numpy as np

from . import gamma_process as gp
from . import hyperband as ohb
from .opt_functions import DOMAINS, FUNCS


def run_batch(func, families, R, eta, noise, n_reps, rng, min_or_max='min'):
    """OFEs [n_reps] of `n_reps` independent Hyperband(R, eta) repetitions (uniform family scheduler)."""
    plan = ohb.bracket_plan(R, eta)
    A = ohb.plan_arms(plan)
    ck = ohb.plan_checkpoints(plan)
    x0, x1, y0, y1 = DOMAINS[func]
    fam = np.asarray(families, dtype=np.float64)              # [F, 6]: ml, nec, up, smooth, start, end
    x = rng.uniform(x0, x1, (n_reps, A))
    y = rng.uniform(y0, y1, (n_reps, A))
    fid = rng.integers(0, len(families), (n_reps, A))
    ml, nec, up, sm, s0, s1 = (fam[fid, j] for j in range(6))
    f = FUNCS[func](x, y)
    f_n = f - s1
    cur = f + noise * rng.standard_normal((n_reps, A)) - s0
    any_smooth = bool(fam[:, 3].any())
    traj = np.empty((R,) + cur.shape) if any_smooth else None
    pos = {t - 1: i for i, t in enumerate(ck)}
    at_ck = np.empty((len(ck),) + cur.shape)
    if 0 in pos:
        at_ck[pos[0]] = cur
    if any_smooth:
        traj[0] = cur
    for t in range(1, R):
        shape, scale = gp.gamma_shape_scale(t, R)
        g = rng.gamma(shape, scale, cur.shape)
        frac = t / (R - 1)
        m = cur + (g - 2) * ml * (f_n - cur) / 100
        down = m + (f_n - m) * frac ** nec
        u = cur + up / (1 + g)
        upv = f_n if R - t == 1 else u + (f_n - u) * frac ** (1.1 * nec)
        cur = np.where(g > 2, down, np.where(g < 2, upv, cur))
        if t in pos:
            at_ck[pos[t]] = cur
        if any_smooth:
            traj[t] = cur
    if any_smooth:
        W = gp.savgol_operator(R)[[t - 1 for t in ck]]       # [C, R]
        sm_ck = np.tensordot(W, traj, axes=(1, 0))
        at_ck = np.where(sm.astype(bool)[None], sm_ck, at_ck)
    desc = min_or_max == 'max'
    best = None
    first = 0
    for b in plan:
        alive = np.broadcast_to(np.arange(first, first + b['n']), (n_reps, b['n']))
        first += b['n']
        for rd in b['rounds']:
            losses = np.take_along_axis(at_ck[ck.index(rd['time'])], alive, axis=1)
            key = -losses if desc else losses
            order = np.argsort(key, axis=1, kind='stable')
            rb = np.take_along_axis(losses, order[:, :1], axis=1)[:, 0]
            best = rb if best is None else (np.maximum(best, rb) if desc else np.minimum(best, rb))
            alive = np.take_along_axis(alive, order[:, :rd['keep']], axis=1)
    return best


def _worker(args):
    import time
    func, families, R, eta, noise, n_reps, block, seed = args
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    out = [run_batch(func, families, R, eta, noise, min(block, n_reps - i), rng) for i in range(0, n_reps, block)]
    return time.perf_counter() - t0, np.concatenate(out)
