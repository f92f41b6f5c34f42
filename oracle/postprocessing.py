"""Oracle: post-processing statistics - loss-curve family profiles (reference
experiments/simulation_evaluation/profiles.py:17-81) and the k-fold cross-validation of the closest-known-function
approximation (experiments/known_functions_evaluation/cross_validate_approximation.py:36-74).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Pinned to the reference's own functions by tests/golden/postprocessing.npz.
"""
import numpy as np

from . import known_fn as okf


def order_agreement(col_a, col_b):
    """Mean over curves i of the fraction of other curves j whose strict order relative to i is the same in both columns
    (profiles.py:29-53 for consecutive epochs, :56-81 for the two ends)."""
    a, b = np.asarray(col_a)[:, None], np.asarray(col_b)[:, None]
    same = ((a.T < a) & (b.T < b)) | ((a.T > a) & (b.T > b))
    n = len(col_a)
    return same.sum(axis=1).astype(np.float64).__truediv__(n - 1).sum() / n


def dynamic_order_profile(curves):
    curves = np.asarray(curves)
    return np.asarray([order_agreement(curves[:, t - 1], curves[:, t]) for t in range(1, curves.shape[1])])


def ends_order_profile(curves):
    curves = np.asarray(curves)
    return order_agreement(curves[:, 0], curves[:, -1])


def cross_validate(n_folds, db_arms, curves, domain):
    """cross_validate_approximation.py:48-74: every arm is looked up in the database minus its own fold; the absolute
    error between its curve and the nearest arm's curve is divided, epoch by epoch, by the range of the recorded losses."""
    n = len(db_arms)
    fold = n // n_folds
    min_len = min(len(c) for c in curves)
    cols = np.asarray([c[:min_len] for c in curves])
    norm = cols.max(axis=0) - cols.min(axis=0)
    res = []
    for i in range(0, n, fold):
        keep = list(range(0, i)) + list(range(i + fold, n))
        sub_arms = [db_arms[k] for k in keep]
        for t in range(i, min(i + fold, n)):
            j, _ = okf.nearest(db_arms[t], sub_arms, domain)
            res.append(np.abs(cols[keep[j]] - cols[t]) / norm)
    return np.asarray(res), min_len
