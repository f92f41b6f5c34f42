"""Glue between the reference-shaped optimiser classes and the batched device entry points (autotune.b200.engine).

An optimiser run on a simulation problem is ONE repetition of the Monte-Carlo job; the same kernels serve both, so
`optimiser.run_optimisation(problem)` and `run_repetitions(optimiser, problem, n)` agree by construction.
"""
import numpy as np

from autotune.core import Arm, Evaluation, OptimisationGoals

DEFAULT_FAMILY = (0.9, 10, 0.1, False, 0, 200)   # defaults of OptFunctionSimulationProblem.get_evaluator


class _Probe(OptimisationGoals):
    pass


def require_fval_objective(optimisation_func):
    """The device kernels rank arms by the simulated loss (`fval`).  Accept any callable that returns exactly that
    attribute; anything else would need arbitrary python per arm and is refused rather than silently run on the CPU."""
    for v in (0.123456789, -987.654321):
        try:
            out = optimisation_func(_Probe(fval=v, test_error=-1, validation_error=-1))
        except Exception as e:  # noqa: BLE001
            raise ValueError(f'optimisation_func failed on a simulation OptimisationGoals: {e!r}') from e
        if out != v:
            raise ValueError('on simulation problems the device path optimises `fval` only: optimisation_func must '
                             'return opt_goals.fval (the reference drivers pass exactly that, run_simulation.py:60-62)')


def simulation_spec(optimiser, problem):
    """HyperbandSimSpec for (optimiser, OptFunctionSimulationProblem): families/noise/R from the scheduler."""
    from autotune.b200.engine import HyperbandSimSpec
    sched = optimiser.scheduler
    if sched is None:
        families, R, noise, kind = [DEFAULT_FAMILY], 81, 0, 'uniform'
    else:
        kind = getattr(sched, 'device_kind', None)
        if kind is None:
            raise NotImplementedError(f'{type(sched).__name__}: only the uniform and round-robin schedulers exist on the '
                                      'device path')
        for fam in sched.shape_families:
            if fam[0] is not None:
                raise NotImplementedError('shape families with a fixed default arm are not supported on the device path')
        families, R, noise = list(sched.shape_families), sched.max_resources, sched.init_noise
    if R != optimiser.max_iter and hasattr(optimiser, 'eta'):
        raise ValueError(f'scheduler.max_resources ({R}) must equal the optimiser\'s max_iter ({optimiser.max_iter})')
    eta = getattr(optimiser, 'eta', 3)
    return HyperbandSimSpec(problem.func_name, families, R, eta, noise, kind, optimiser.min_or_max)


def fresh_seed():
    """Seed for one run, drawn from numpy's global RNG so that np.random.seed() makes runs reproducible."""
    return int(np.random.randint(0, 2 ** 31 - 1))


def history_from_rounds(optimiser, problem, spec, out, seed, rep=0, rep_offset=0):
    """Fill optimiser.eval_history with the best evaluation of every round (hyperband_optimiser.py:98-99), built from
    the device outputs of repetition `rep` of the job keyed by (seed, rep_offset); returns the best Evaluation
    (core/optimiser.py:67-68).  Every evaluator is rebuilt with the shape family its arm really drew on the device
    (the reference keeps the evaluator object itself, hyperband_optimiser.py:99-100)."""
    from autotune.b200.engine import arm_families
    rb = out['round_best'][rep].tolist()
    rb_arm = out['round_best_arm'][rep].tolist()
    xy = out['arm_xy'][rep].tolist()
    fams = spec.families
    fam_of = arm_families(spec, seed, rep_offset + rep, rb_arm).tolist()
    for loss, arm_id, fam_id in zip(rb, rb_arm, fam_of):
        x, y = xy[arm_id]
        evaluator = problem.get_evaluator(Arm(x=x, y=y), *fams[fam_id], spec.R, spec.init_noise)
        optimiser.eval_history.append(Evaluation(evaluator, OptimisationGoals(fval=loss, test_error=-1,
                                                                               validation_error=-1)))
        optimiser._update_optimiser_metrics()
    return optimiser._get_best_evaluation()
