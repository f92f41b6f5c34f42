"""Oracle: Hyperband whose first round of every bracket is a TPE run, with history transfer between brackets
(reference optimisers/sequential/hybrid_hyperband_tpe_optimiser.py, hybrid_hyperband_tpe_with_transfer.py,
tpe_optimiser.py) on the simulation problem, plus stand-alone TPE / random search.

TEST INFRASTRUCTURE - see oracle/__init__.py.  TPE arithmetic is the unpinned restatement in oracle/tpe.py.
"""
import random

import numpy as np

from . import hyperband as ohb
from .opt_functions import DOMAINS
from .tpe import TpeRun

TRANSFER_MODES = ('none', 'all', 'longest', 'same', 'threshold')


def is_transferable(mode, r_arm, r_bracket, R, threshold=None):
    """hybrid_hyperband_tpe_with_transfer.py:161-209."""
    if mode == 'none':
        return False
    if mode == 'longest':
        return r_arm >= R
    if mode == 'all':
        return r_arm >= r_bracket
    if mode == 'same':
        return r_arm == r_bracket
    if mode == 'threshold':
        return r_arm >= threshold and r_arm >= r_bracket
    raise ValueError(mode)


def _new_arm(func, families, R, noise, scheduler, rr, x, y):
    fam = random.choice(families) if scheduler == 'uniform' else families[rr[0] % len(families)]
    rr[0] += 1
    return ohb._LazyArm(func, x, y, fam, R, noise)


def run_hybrid(func, families, R, eta, noise, transfer='none', min_or_max='min', scheduler='uniform', threshold=None,
               rstate=None):
    """One repetition of HybridHyperbandTpe*Optimiser.run_optimisation; returns the OFE.

    Per bracket (hybrid_hyperband_tpe_with_transfer.py:66-104): round 0 = TpeOptimiser(n_resources=r_i,
    max_iter=n_i + injected) whose objective builds an evaluator for the suggested arm and evaluates it at r_i
    (tpe_optimiser.py:62-89, loss sign-flipped for max); later rounds are plain successive halving; every
    evaluated arm's latest (int(r_i), loss) is remembered for transfer into later brackets."""
    x0, x1, y0, y1 = DOMAINS[func]
    desc = min_or_max == 'max'
    sign = -1.0 if desc else 1.0
    plan = ohb.bracket_plan(R, eta)
    history = []                                   # [arm, latest_r, latest_loss] in first-evaluation order
    rb, rr = [], [0]
    for b in plan:
        rounds = b['rounds']
        r0 = rounds[0]['resource']
        injected = []
        if rb:                                     # _get_trials: nothing is injected into the very first bracket
            injected = [((h[0].x, h[0].y), sign * h[2]) for h in history
                        if is_transferable(transfer, h[1], r0, R, threshold)]
        tpe = TpeRun([(x0, x1), (y0, y1)], injected, rstate)
        alive, entries = [], []
        for _ in range(int(b['n'])):               # fmin: max_evals - injected = n_i new sequential evaluations
            x, y = tpe.suggest()
            arm = _new_arm(func, families, R, noise, scheduler, rr, x, y)
            loss = arm.evaluate(r0)
            tpe.observe((x, y), sign * loss)
            entry = [arm, int(r0), loss]
            history.append(entry)
            alive.append(arm)
            entries.append(entry)
        by_arm = {id(e[0]): e for e in entries}
        losses = [by_arm[id(a)][2] for a in alive]
        for i, rd in enumerate(rounds):
            if i > 0:
                losses = [a.evaluate(rd['resource']) for a in alive]
                for a, l in zip(alive, losses):
                    by_arm[id(a)][1], by_arm[id(a)][2] = int(rd['resource']), l
            surv, _, best = ohb.halve(alive, losses, rd['keep'], desc)
            rb.append(best)
            alive = surv
    return (max if desc else min)(rb)


def run_tpe(func, families, R, noise, n_resources, max_iter, min_or_max='min', scheduler='uniform', rstate=None):
    """Stand-alone TpeOptimiser(n_resources, max_iter) (tpe_optimiser.py:40-60): OFE = best of max_iter evaluations."""
    x0, x1, y0, y1 = DOMAINS[func]
    sign = -1.0 if min_or_max == 'max' else 1.0
    tpe = TpeRun([(x0, x1), (y0, y1)], (), rstate)
    rr, losses = [0], []
    for _ in range(int(max_iter)):
        x, y = tpe.suggest()
        loss = _new_arm(func, families, R, noise, scheduler, rr, x, y).evaluate(n_resources)
        tpe.observe((x, y), sign * loss)
        losses.append(loss)
    return (max if sign < 0 else min)(losses)


def run_random(func, families, R, noise, n_resources, max_iter, min_or_max='min', scheduler='uniform'):
    """RandomOptimiser (random_optimiser.py:39-58): max_iter random arms, each evaluated at n_resources."""
    x0, x1, y0, y1 = DOMAINS[func]
    rr, losses = [0], []
    for _ in range(int(max_iter)):
        fam = random.choice(families) if scheduler == 'uniform' else families[rr[0] % len(families)]
        rr[0] += 1
        x = (np.random.rand(1) * (x1 - x0) + x0)[0]
        y = (np.random.rand(1) * (y1 - y0) + y0)[0]
        losses.append(ohb._LazyArm(func, x, y, fam, R, noise).evaluate(n_resources))
    return (max if min_or_max == 'max' else min)(losses)
