"""Canonical 2-D optimisation test functions as instant "problems" (reference benchmarks/opt_function_problem.py).

Functions follow https://www.sfu.ca/~ssurjano/optimization.html with the reference's constants and expression order
(opt_function_problem.py:49-138), so float64 values are bit-identical to the reference's.  They accept scalars or
numpy arrays; every call consumes one N(0,1) draw from numpy's global RNG for its noise term, as the reference does.
"""
import numpy as np

from autotune.core import Arm, Domain, Evaluator, HyperparameterOptimisationProblem, ModelBuilder, OptimisationGoals, Param

AVAILABLE_OPT_FUNCTIONS = ('rastrigin', 'wave', 'branin', 'egg', 'camel')


def _square_domain(lo_x, hi_x, lo_y, hi_y):
    return Domain(x=Param('x', lo_x, hi_x, distrib='uniform', scale='linear'),
                  y=Param('y', lo_y, hi_y, distrib='uniform', scale='linear'))


HYPERPARAMS_DOMAIN_EGGHOLDER = _square_domain(-512, 512, -512, 512)
HYPERPARAMS_DOMAIN_BRANIN = _square_domain(-5, 10, 1, 15)
HYPERPARAMS_DOMAIN_CAMEL = _square_domain(-3, 3, -2, 2)
HYPERPARAMS_DOMAIN_WAVE = _square_domain(-5.12, 5.12, -5.12, 5.12)

GLOBAL_MIN_EGGHOLDER = -959.6407
GLOBAL_MIN_BRANIN = 0.397887
GLOBAL_MIN_CAMEL = -1.0316
GLOBAL_MIN_WAVE = -1
GLOBAL_MIN_RASTRIGIN = 0

DOMAINS = {'egg': HYPERPARAMS_DOMAIN_EGGHOLDER, 'branin': HYPERPARAMS_DOMAIN_BRANIN, 'camel': HYPERPARAMS_DOMAIN_CAMEL,
           'wave': HYPERPARAMS_DOMAIN_WAVE, 'rastrigin': HYPERPARAMS_DOMAIN_WAVE}


def _noisy(f, global_min, noise_variance, scaled_noise):
    """f + noise_variance * N(0,1) * (f - global_min if scaled_noise else 1)."""
    return f + noise_variance * np.random.randn(1)[0] * ((f - global_min) if scaled_noise else 1)


def egg_holder_value(x1, x2):
    return -(x2 + 47) * np.sin(np.sqrt(abs(x2 + 0.5 * x1 + 47))) - x1 * np.sin(np.sqrt(abs(x1 - x2 - 47)))


def branin_value(x1, x2):
    b, c, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 1 / (8 * np.pi)
    return 1 * (x2 - b * x1 ** 2 + c * x1 - 6) ** 2 + 10 * (1 - t) * np.cos(x1) + 10


def six_hump_camel_value(x1, x2):
    f = (4 - (2.1 * (x1 ** 2)) + ((x1 ** 4) / 3)) * (x1 ** 2)
    f += x1 * x2
    f += (-4 + 4 * (x2 ** 2)) * (x2 ** 2)
    return f


def drop_wave_value(x1, x2):
    r2 = x1 ** 2 + x2 ** 2
    return -(1 + np.cos(12 * np.sqrt(r2))) / (0.5 * r2 + 2)


def rastrigin_value(x1, x2):
    t1 = x1 ** 2 - 10 * np.cos(2 * np.pi * x1)
    t2 = x2 ** 2 - 10 * np.cos(2 * np.pi * x2)
    return 10 * 2 + ((0 + t1) + t2)


NOISE_FREE_FUNCTIONS = {'egg': egg_holder_value, 'branin': branin_value, 'camel': six_hump_camel_value,
                        'wave': drop_wave_value, 'rastrigin': rastrigin_value}


def egg_holder(x1, x2, noise_variance=0, scaled_noise=False):
    return _noisy(egg_holder_value(x1, x2), GLOBAL_MIN_EGGHOLDER, noise_variance, scaled_noise)


def branin(x1, x2, noise_variance=0, scaled_noise=False):
    return _noisy(branin_value(x1, x2), GLOBAL_MIN_BRANIN, noise_variance, scaled_noise)


def six_hump_camel(x1, x2, noise_variance=0, scaled_noise=False):
    return _noisy(six_hump_camel_value(x1, x2), GLOBAL_MIN_CAMEL, noise_variance, scaled_noise)


def drop_wave(x1, x2, noise_variance=0, scaled_noise=False):
    return _noisy(drop_wave_value(x1, x2), GLOBAL_MIN_WAVE, noise_variance, scaled_noise)


def rastrigin(x1, x2, noise_variance=0, scaled_noise=False):
    return _noisy(rastrigin_value(x1, x2), GLOBAL_MIN_RASTRIGIN, noise_variance, scaled_noise)


OPT_FUNCTIONS = {'egg': egg_holder, 'branin': branin, 'camel': six_hump_camel, 'wave': drop_wave,
                 'rastrigin': rastrigin}


class OptFunctionBuilder(ModelBuilder):
    """A test function has no model; the builder only carries the arm."""

    def __init__(self, arm):
        super().__init__(arm)


class OptFunctionEvaluator(Evaluator):

    def __init__(self, func_name, model_builder):
        super().__init__(model_builder)
        self.func_name = func_name

    def evaluate(self, n_resources):
        """The function value at the arm; `n_resources` is ignored (every optimiser passes it)."""
        return OptimisationGoals(fval=OPT_FUNCTIONS[self.func_name](self.arm.x, self.arm.y, noise_variance=0),
                                 test_error=-1, validation_error=-1)

    def _train(self, epoch, max_batches, batch_size):
        raise TypeError('Cannot call _train on well defined loss functions')

    def _test(self, is_validation):
        raise TypeError('Cannot call _test on well defined loss functions')

    def _save_checkpoint(self, epoch, val_error, test_error):
        raise TypeError('Cannot call _save_checkpoint on well defined loss functions')


class OptFunctionProblem(HyperparameterOptimisationProblem):

    def __init__(self, func_name, hyperparams_to_opt=()):
        super().__init__(DOMAINS[func_name], hyperparams_to_opt)
        self.func_name = func_name

    def _arm_or_random(self, arm):
        if arm is None:
            arm = Arm()
            arm.draw_hp_val(domain=self.domain, hyperparams_to_opt=self.hyperparams_to_opt)
        return arm

    def get_evaluator(self, arm=None):
        return OptFunctionEvaluator(self.func_name, OptFunctionBuilder(self._arm_or_random(arm)))
