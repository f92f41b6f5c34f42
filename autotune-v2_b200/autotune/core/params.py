"""Hyperparameter range description (reference core/params.py:12-66; originally from L. Li's Hyperband code).

Only the continuous `Param` is on the simulation / closest-known-function path; the categorical, pair, conditional and
factored parameter kinds of the reference belong to its real-ML problems and are out of scope here.
"""
import numpy as np


class Param:
    """A continuous hyperparameter.  For `scale='log'`, `min_val`/`max_val` are exponents of `logbase`."""

    def __init__(self, name, min_val, max_val, init_val=None, distrib='uniform', scale='log', logbase=np.e,
                 interval=None):
        self.name, self.min_val, self.max_val, self.init_val = name, min_val, max_val, init_val
        self.distrib, self.scale, self.logbase, self.interval = distrib, scale, logbase, interval
        self.param_type = 'continuous'

    def __repr__(self):
        return '%s (%f,%f,%s)' % (self.name, self.min_val, self.max_val, self.scale)

    def _finish(self, val):
        if self.scale == 'log':
            val = np.array([self.logbase ** v for v in val])
        if self.interval:
            return (np.floor(val / self.interval) * self.interval).astype(int)
        return val

    def get_param_range(self, num_vals, stochastic=False):
        """`num_vals` values: random draws (numpy global RNG, as the reference) or an even grid."""
        if stochastic:
            if self.distrib == 'normal':   # reference convention: min_val is the mean, max_val the sigma
                val = np.random.normal(self.min_val, self.max_val, num_vals)
            else:
                val = np.random.rand(num_vals) * (self.max_val - self.min_val) + self.min_val
            return self._finish(val)
        if self.scale == 'log':
            val = np.logspace(self.min_val, self.max_val, num_vals, base=self.logbase)
        else:
            val = np.linspace(self.min_val, self.max_val, num_vals)
        if self.interval:
            return (np.floor(val / self.interval) * self.interval).astype(int)
        return val

    def get_transformed_param(self, x):
        if self.distrib == 'normal':
            print('not implemented')
            return None
        val = self.logbase ** x if self.scale == 'log' else x
        if self.interval:
            val = (np.floor(val / self.interval) * self.interval).astype(int)
        return val

    def get_min(self):
        return self.min_val

    def get_max(self):
        return self.max_val

    def get_type(self):
        return 'integer' if self.interval else 'continuous'
