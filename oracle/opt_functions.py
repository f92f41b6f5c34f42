"""Oracle: 2-D test functions and their domains (reference autotune/benchmarks/opt_function_problem.py:15-147).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Operation order follows the reference expressions exactly so that
float64 results are bit-identical to the reference's.
"""
import math

import numpy as np

# opt_function_problem.py:15-46 - (x_min, x_max, y_min, y_max)
DOMAINS = {
    'egg': (-512, 512, -512, 512),
    'branin': (-5, 10, 1, 15),
    'camel': (-3, 3, -2, 2),
    'wave': (-5.12, 5.12, -5.12, 5.12),
    'rastrigin': (-5.12, 5.12, -5.12, 5.12),
}
FUNC_IDS = {'egg': 0, 'branin': 1, 'camel': 2, 'wave': 3, 'rastrigin': 4}


def egg(x, y):  # opt_function_problem.py:60
    return -(y + 47) * np.sin(np.sqrt(abs(y + 0.5 * x + 47))) - x * np.sin(np.sqrt(abs(x - y - 47)))


def branin(x, y):  # opt_function_problem.py:75-82
    b = 5.1 / (4 * np.pi ** 2)
    c = 5 / np.pi
    t = 1 / (8 * np.pi)
    return 1 * (y - b * x ** 2 + c * x - 6) ** 2 + 10 * (1 - t) * np.cos(x) + 10


def camel(x, y):  # opt_function_problem.py:97-99
    f = (4 - (2.1 * (x ** 2)) + ((x ** 4) / 3)) * (x ** 2)
    f += x * y
    f += (-4 + 4 * (y ** 2)) * (y ** 2)
    return f


def wave(x, y):  # opt_function_problem.py:114-116
    f1 = 1 + np.cos(12 * np.sqrt(x ** 2 + y ** 2))
    f2 = 0.5 * (x ** 2 + y ** 2) + 2
    return -f1 / f2


def rastrigin(x, y):  # opt_function_problem.py:131-137 (python sum() starts from int 0)
    tx = x ** 2 - 10 * np.cos(2 * np.pi * x)
    ty = y ** 2 - 10 * np.cos(2 * np.pi * y)
    return 10 * 2 + ((0 + tx) + ty)


def egg4(x0, x1, x2, x3):
    """EXTENSION - no reference equivalent (the reference's Egg-holder is 2-D; BASELINE.json config 5 asks
    for a 4-D one): the n-dimensional Egg-holder, the 2-D form (:60) summed over consecutive coordinate pairs.  This is
    the definition the device follows (include/at2.h AT2_FUNC_EGG4), not a restatement of reference code."""
    return (egg(x0, x1) + egg(x1, x2)) + egg(x2, x3)


FUNCS = {'egg': egg, 'branin': branin, 'camel': camel, 'wave': wave, 'rastrigin': rastrigin}


def base_value(func, x, y):
    """Noise-free value as the reference's evaluate() sees it: x, y are numpy float64 scalars there
    (params.py:36 returns an ndarray element), so numpy scalar arithmetic applies."""
    return FUNCS[func](np.float64(x), np.float64(y))


def _isfinite(v):
    return not (math.isinf(v) or math.isnan(v))
