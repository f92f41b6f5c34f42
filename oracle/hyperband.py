"""Oracle: Hyperband over simulated arms (reference autotune/optimisers/sequential/hyperband_optimiser.py,
autotune/benchmarks/opt_function_simulation_problem.py, autotune/core/shape_family_scheduler.py).

TEST INFRASTRUCTURE - see oracle/__init__.py.
"""
import random
from math import ceil

import mpmath
import numpy as np

from . import gamma_process as gp
from .opt_functions import DOMAINS, base_value

_MP = mpmath.mp.clone()
_MP.dps = 64  # hyperband_optimiser.py:13


def bracket_plan(R, eta):
    """Bracket/round table exactly as the reference's float arithmetic produces it (hyperband_optimiser.py:66-97).

    Returns a list of brackets, each dict(s, n, rounds=[dict(n_eval, resource (float), time (int), keep)]).
    n_eval of round 0 is n; of later rounds the previous keep count (the evaluators list is what survived).
    """
    s_max = int(_MP.log(R) / _MP.log(eta))                 # :68-69
    s_min = 2 if s_max >= 2 else 0                         # :70
    B = (s_max + 1) * R                                    # :71
    plan = []
    for s in reversed(range(s_min, s_max + 1)):            # :74
        n = int(ceil(int(B / R / (s + 1)) * eta ** s))     # :75
        r = R * eta ** (-s)                                # :76
        rounds, alive = [], n
        for i in range(s + 1):                             # :89
            n_i = n * eta ** (-i)                          # :90
            r_i = r * eta ** i                             # :91
            keep = int(n_i / eta)                          # :96
            rounds.append(dict(n_eval=alive, resource=r_i, time=int(r_i), keep=keep))
            alive = min(keep, alive)                       # sorted(...)[:keep] of a list of `alive`
        plan.append(dict(s=s, n=n, rounds=rounds))
    return plan


def plan_arms(plan):
    return sum(b['n'] for b in plan)


def plan_checkpoints(plan):
    return sorted({rd['time'] for b in plan for rd in b['rounds']})


def halve(ids, losses, keep, descending=False):
    """Survivors of one round in sorted order + the round-best (hyperband_optimiser.py:46-57,96-99).

    Stable sort (python `sorted`, reverse=True keeps the original order of equal keys); round-best is the FIRST
    min (or max)."""
    order = sorted(range(len(ids)), key=lambda j: losses[j], reverse=descending)
    surv = [ids[j] for j in order[:keep]]
    pick = max if descending else min
    jbest = pick(range(len(ids)), key=lambda j: losses[j])
    return surv, ids[jbest], losses[jbest]


def run_predrawn(func, families, R, eta, noise, xy, fam, z, gamma, min_or_max='min'):
    """One Hyperband repetition on pre-drawn inputs (the layout of tests/golden/hb_*.npz).

    families[f] = (ml_agg, nec_agg, up_spik, is_smooth, start_shift, end_shift) (shape_family_scheduler.py:10-17).
    f_1 = f(x,y) + noise*z - start_shift, f_n = f(x,y) - end_shift (opt_function_simulation_problem.py:61-64).
    """
    plan = bracket_plan(R, eta)
    A = plan_arms(plan)
    assert len(z) == A
    desc = min_or_max == 'max'
    raw, smooth = [], []
    for a in range(A):
        ml, nec, up, is_s, s0, s1 = families[int(fam[a])]
        f = base_value(func, float(xy[a, 0]), float(xy[a, 1]))
        f_n = f - s1
        f_1 = (f + noise * z[a] * 1) - s0
        tr = gp.simulate(f_1, f_n, R, ml, nec, up, lambda t, a=a: float(gamma[a, t - 1]))
        raw.append(tr)
        smooth.append(gp.readout(tr, bool(is_s), R))
    eval_ids, eval_f, round_off, surv_ids, surv_off, keeps, rb, rb_id = [], [], [0], [], [0], [], [], []
    first = 0
    for b in plan:
        alive = list(range(first, first + b['n']))
        first += b['n']
        for rd in b['rounds']:
            losses = [smooth[a][rd['time'] - 1] for a in alive]
            surv, best_id, best = halve(alive, losses, rd['keep'], desc)
            eval_ids += alive
            eval_f += losses
            round_off.append(len(eval_ids))
            surv_ids += surv
            surv_off.append(len(surv_ids))
            keeps.append(rd['keep'])
            rb.append(best)
            rb_id.append(best_id)
            alive = surv
    pick = max if desc else min
    j = pick(range(len(rb)), key=lambda i: rb[i])          # core/optimiser.py:67-68 (first optimum)
    ck = plan_checkpoints(plan)
    return dict(traj_raw=np.asarray(raw), traj=np.asarray(smooth, dtype=np.float64),
                ckpts=np.asarray(ck, np.int32), ckpt_loss=np.asarray(smooth, dtype=np.float64)[:, np.asarray(ck) - 1],
                eval_ids=np.asarray(eval_ids, np.int32), eval_fvals=np.asarray(eval_f, np.float64),
                round_off=np.asarray(round_off, np.int32), surv_ids=np.asarray(surv_ids, np.int32),
                surv_off=np.asarray(surv_off, np.int32), keep=np.asarray(keeps, np.int32),
                round_best=np.asarray(rb, np.float64), round_best_id=np.asarray(rb_id, np.int32),
                ofe=np.float64(rb[j]), ofe_id=np.int32(rb_id[j]))


class _LazyArm:
    """An arm as the reference's OptFunctionSimulationEvaluator treats it in the global-RNG path: it simulates on its
    first evaluate(), re-simulates when first read at time 1, and re-filters on every read
    (opt_function_simulation_problem.py:50-86).  Draws come from numpy's legacy global RandomState in the reference's
    order so a seeded run reproduces the reference's OFEs bit-for-bit."""
    __slots__ = ('func', 'x', 'y', 'fam', 'R', 'noise', 'raw')

    def __init__(self, func, x, y, fam, R, noise):
        self.func, self.x, self.y, self.fam, self.R, self.noise, self.raw = func, x, y, fam, R, noise, None

    def _simulate(self, f_1, f_n):
        ml, nec, up = self.fam[0], self.fam[1], self.fam[2]
        R = self.R

        def draw(t):
            shape, scale = gp.gamma_shape_scale(t, R)
            return np.random.gamma(shape, scale)           # simulation_evaluator.py:44
        self.raw = gp.simulate(f_1, f_n, R, ml, nec, up, draw)

    def evaluate(self, n_resources):
        time = int(n_resources)                            # :58
        f = base_value(self.func, self.x, self.y)
        base = f + 0 * np.random.randn(1)[0] * (f - 0.0)   # :61 - consumes one N(0,1); contributes 0
        noisy = f + self.noise * np.random.randn(1)[0] * 1  # :62
        f_n = base - self.fam[5]
        f_1 = noisy - self.fam[4]
        if self.raw is None:                               # :67-69
            self._simulate(f_1, f_n)
        if time == 1:                                      # :71-74
            self._simulate(f_1, f_n)
            return gp.readout(self.raw, bool(self.fam[3]), self.R)[0]
        if time == self.R and self.fam[1] != np.inf:       # :83-84
            assert self.raw[self.R - 1] - f_n < 0.0001
        return gp.readout(self.raw, bool(self.fam[3]), self.R)[time - 1]


def run_global_rng(func, families, R, eta, noise, min_or_max='min', scheduler='uniform', _rr_state=None):
    """One repetition the way run_simulation.py:147-157 does it: arms drawn from numpy's global RNG
    (arm.py:44-55, params.py:36), family by python `random.choice` (shape_family_scheduler.py:75) or round-robin,
    lazy simulation per arm.  Returns the OFE."""
    x0, x1, y0, y1 = DOMAINS[func]
    desc = min_or_max == 'max'
    plan = bracket_plan(R, eta)
    rb = []
    rr = _rr_state if _rr_state is not None else [0]
    for b in plan:
        alive = []
        for _ in range(b['n']):
            if scheduler == 'uniform':
                fam = random.choice(families)
            else:
                fam = families[rr[0] % len(families)]
                rr[0] += 1
            x = (np.random.rand(1) * (x1 - x0) + x0)[0]
            y = (np.random.rand(1) * (y1 - y0) + y0)[0]
            alive.append(_LazyArm(func, x, y, fam, R, noise))
        for rd in b['rounds']:
            losses = [arm.evaluate(rd['resource']) for arm in alive]
            surv, _, best = halve(alive, losses, rd['keep'], desc)
            rb.append(best)
            alive = surv
    return (max if desc else min)(rb)
