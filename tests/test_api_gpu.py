"""The reference-shaped Python API on top of the device kernels: same classes, arguments and error behaviour as
autotune/core, autotune/benchmarks, autotune/optimisers and the experiments/run_simulation.py driver."""
import io
import pickle
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import autotune.b200  # noqa: F401
from autotune.benchmarks import KnownFnProblem, OptFunctionSimulationProblem
from autotune.core import (
    Arm, Domain, Evaluation, HyperparameterOptimisationProblem, OptimisationGoals, Param, RoundRobinShapeFamilyScheduler,
    ShapeFamily, UniformShapeFamilyScheduler)
from autotune.optimisers import (
    HybridHyperbandTpeNoTransferOptimiser, HybridHyperbandTpeOptimiser, HybridHyperbandTpeTransferAllOptimiser,
    HybridHyperbandTpeTransferThresholdOptimiser, HyperbandOptimiser, RandomOptimiser, TpeOptimiser)

pytestmark = pytest.mark.gpu

FAMILIES = (ShapeFamily(None, 1.5, 10, 15, False), ShapeFamily(None, 0.5, 7, 10, False), ShapeFamily(None, 0.2, 4, 7, True),
            ShapeFamily(None, 1.5, 10, 15, False, 200, 400), ShapeFamily(None, 0.5, 7, 10, False, 200, 400),
            ShapeFamily(None, 0.2, 4, 7, True, 200, 400))


def fval(goals):
    """fval."""
    return goals.fval


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _problem(name='rastrigin'):
    return _quiet(OptFunctionSimulationProblem, name)


def test_hyperband_run_optimisation_surface():
    problem = _problem()
    sched = UniformShapeFamilyScheduler(FAMILIES, max_resources=81, init_noise=10)
    opt = HyperbandOptimiser(eta=3, max_iter=81, min_or_max=min, optimisation_func=fval, is_simulation=True, scheduler=sched)
    np.random.seed(4)
    best = _quiet(opt.run_optimisation, problem, verbosity=True)
    assert isinstance(best, Evaluation) and isinstance(best.optimisation_goals, OptimisationGoals)
    assert len(opt.eval_history) == 12 and opt.num_iterations == 12 and len(opt.checkpoints) == 12
    losses = [fval(e.optimisation_goals) for e in opt.eval_history]
    assert fval(best.optimisation_goals) == min(losses)
    assert -5.12 <= best.evaluator.arm.x <= 5.12 and -5.12 <= best.evaluator.arm.y <= 5.12
    assert best.optimisation_goals.test_error == -1 and best.optimisation_goals.validation_error == -1
    # np.random.seed makes a run reproducible
    opt2 = HyperbandOptimiser(eta=3, max_iter=81, optimisation_func=fval, is_simulation=True, scheduler=sched)
    np.random.seed(4)
    best2 = _quiet(opt2.run_optimisation, problem)
    assert fval(best2.optimisation_goals) == fval(best.optimisation_goals)
    # max: descending halving
    opt3 = HyperbandOptimiser(eta=3, max_iter=27, min_or_max=max, optimisation_func=fval, is_simulation=True,
                              scheduler=RoundRobinShapeFamilyScheduler(FAMILIES, 27, 10))
    best3 = _quiet(opt3.run_optimisation, problem)
    assert fval(best3.optimisation_goals) == max(fval(e.optimisation_goals) for e in opt3.eval_history)


def test_error_behaviour_matches_reference():
    problem = _problem()
    sched = UniformShapeFamilyScheduler(FAMILIES, 81, 10)
    with pytest.raises(ValueError, match='max_iter cannot be None'):
        HyperbandOptimiser(eta=3, max_time=10)
    with pytest.raises(ValueError, match='cannot be None simultaneously'):
        HyperbandOptimiser(eta=3)
    with pytest.raises(ValueError, match='min or max'):
        HyperbandOptimiser(eta=3, max_iter=81, min_or_max=sum)

    class NotASimulation(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise AssertionError
    plain = _quiet(NotASimulation, Domain(x=Param('x', 0, 1, distrib='uniform', scale='linear')))
    opt = HyperbandOptimiser(eta=3, max_iter=81, optimisation_func=fval, is_simulation=True, scheduler=sched)
    with pytest.raises(ValueError, match='shape family schedule'):
        opt.run_optimisation(plain)
    # the default optimisation_func (validation_error) is not the simulated loss: refused, not run on the CPU
    with pytest.raises(ValueError, match='fval'):
        HyperbandOptimiser(eta=3, max_iter=81, is_simulation=True, scheduler=sched).run_optimisation(problem)
    with pytest.raises(ValueError, match='window_length'):    # R=6 with a smooth family: scipy's error in the reference
        HyperbandOptimiser(eta=3, max_iter=6, optimisation_func=fval, is_simulation=True,
                           scheduler=UniformShapeFamilyScheduler(FAMILIES, 6, 10)).run_optimisation(problem)


def test_single_evaluator_api():
    problem = _problem('branin')
    sched = RoundRobinShapeFamilyScheduler(FAMILIES, max_resources=81, init_noise=10)
    for k in range(3):                       # families 0, 1 (raw) and 2 (smooth)
        ev = problem.get_evaluator(*sched.get_family())
        goals = ev.evaluate(n_resources=27.0)
        assert len(ev.non_smooth_fs) == 81 and len(ev.fs) == 81
        assert goals.fval == ev.fs[26] and goals.test_error == -1
        from autotune.benchmarks.opt_function_problem import branin_value
        f_n = branin_value(ev.arm.x, ev.arm.y) - 200
        assert abs(ev.non_smooth_fs[-1] - f_n) < 1e-3            # the trajectory lands on its target
        assert ev.evaluate(81).fval == ev.fs[80]
        first = list(ev.non_smooth_fs)
        assert ev.evaluate(9).fval == ev.fs[8] and ev.non_smooth_fs == first     # no re-simulation above time 1
        ev.evaluate(1)                                                           # time 1 re-simulates (reference :71-74)
        assert ev.non_smooth_fs != first
        assert ev.is_smooth == (k == 2)
    given = problem.get_evaluator(Arm(x=1.0, y=2.0), 0, 1, 0, False, 0, 0, 81, 0)   # flat family: constant curve
    assert given.evaluate(50).fval == pytest.approx(branin_value(1.0, 2.0), rel=1e-6)


def test_trajectory_kernel_agrees_with_fused_kernel():
    """The full-trajectory kernel and the fused kernel consume the same Philox streams: at the Hyperband checkpoints
    the values coincide bit-for-bit (raw families) / through the Savitzky-Golay operator (smooth families)."""
    from autotune.b200 import trajectories
    from autotune.b200.engine import HyperbandSimSpec, run_optimiser_repetitions, tpe_config
    fams = [tuple(f[1:]) for f in FAMILIES]
    spec = HyperbandSimSpec('rastrigin', fams, 81, 3, 10)
    n, seed, off = 3, 17, 5
    out = run_optimiser_repetitions(spec, n, tpe_config(False), seed=seed, rep_offset=off, details=True)
    A = spec.n_arms
    ck = out['ckpt_loss'].cpu().numpy()
    xy = out['arm_xy'].cpu().numpy().astype(np.float64)
    # family and noise of every arm are not exported; recover (f_1, f_n, family) from the checkpoint at time 1 / R:
    # f_1 = loss at position 0 (raw arms), f_n = f(x,y) - end_shift; try each family and keep the consistent one
    from autotune.benchmarks.opt_function_problem import rastrigin_value
    ids = (np.arange(n)[:, None] + off) * A + np.arange(A)[None, :]
    pos = np.asarray(spec.checkpoints) - 1
    checked = 0
    for f_idx, fam in enumerate(fams):
        if fam[3]:
            continue                                     # raw families: position 0 is f_1 itself
        f = rastrigin_value(xy[..., 0], xy[..., 1])
        f_n = (f - fam[5]).astype(np.float32)
        f_1 = ck[..., 0]
        raw = trajectories.simulate_arms(spec, f_1.ravel(), f_n.ravel(), np.full(n * A, f_idx), ids.ravel(), seed=seed)
        raw = raw.cpu().numpy().reshape(n, A, 81)
        match = np.all(raw[..., pos] == ck, axis=-1)     # arms that really belong to this family match at every ckpt
        checked += int(match.sum())
    assert checked > n * A * 0.5                         # 4 of the 6 families are raw: ~2/3 of the arms


@pytest.mark.parametrize('cls,kw', [(HybridHyperbandTpeOptimiser, {}), (HybridHyperbandTpeNoTransferOptimiser, {}),
                                    (HybridHyperbandTpeTransferAllOptimiser, {}),
                                    (HybridHyperbandTpeTransferThresholdOptimiser, {'n_res_transfer_threshold': 9})])
def test_hybrid_optimisers(cls, kw):
    problem = _problem()
    sched = UniformShapeFamilyScheduler(FAMILIES, 81, 10)
    opt = cls(eta=3, max_iter=81, optimisation_func=fval, is_simulation=True, scheduler=sched, **kw)
    best = _quiet(opt.run_optimisation, problem)
    assert len(opt.eval_history) == 12
    assert fval(best.optimisation_goals) == min(fval(e.optimisation_goals) for e in opt.eval_history)
    with pytest.raises(NotImplementedError):
        cls(eta=3, max_iter=81, optimisation_func=fval, **kw).run_optimisation(problem)      # real problems need hyperopt


def test_tpe_and_random_optimisers():
    problem = _problem()
    sched = UniformShapeFamilyScheduler(FAMILIES, 81, 10)
    for cls in (TpeOptimiser, RandomOptimiser):
        opt = cls(n_resources=24, max_iter=81, optimisation_func=fval, is_simulation=True, scheduler=sched)
        best = _quiet(opt.run_optimisation, problem)
        assert len(opt.eval_history) == 81
        assert fval(best.optimisation_goals) == min(fval(e.optimisation_goals) for e in opt.eval_history)


def test_run_repetitions_and_cli_pickle(tmp_path):
    from autotune.experiments import run_simulation
    from autotune.experiments.monte_carlo import run_repetitions
    from autotune.experiments.simulation_evaluation.plot_results import plot_epdf_ofe
    problem = _problem()
    sched = UniformShapeFamilyScheduler(FAMILIES, 81, 10)
    opt = HyperbandOptimiser(eta=3, max_iter=81, optimisation_func=fval, is_simulation=True, scheduler=sched)
    ofe = run_repetitions(opt, problem, 5000, seed=1)
    assert ofe.shape == (5000,) and ofe.is_cuda and abs(ofe.mean().item() + 385) < 1.0      # reference: -384.95
    dens = plot_epdf_ofe((ofe,), ['Hyperband'], start=-400, end=-370, bandwidth=3)[0]
    assert dens.shape == (1000,) and dens.max() > 0.02
    for method in ('sim(hb)', 'sim(hb+tpe+transfer+all)', 'sim(tpe)', 'sim(random)', 'sim(hb+tpe)'):
        stats = _quiet(run_simulation.main, ['-m', method, '-n', '400', '--output-dir', str(tmp_path), '--seed', '3'])
        with open(tmp_path / f'hist-sim-rastrigin-{method}-400.pkl', 'rb') as f:
            saved = pickle.load(f)
        assert set(saved) == {'method', 'n_simulations', 'all_optimums', 'avg_optimum', 'sample_std_dev', 'avg_top_25%',
                              'top_quartile'}                       # run_simulation.py:172-182
        assert saved['n_simulations'] == 400 and len(saved['all_optimums']) == 400 and len(saved['top_quartile']) == 100
        assert saved['avg_optimum'] == pytest.approx(np.mean(saved['all_optimums']))
        assert saved['sample_std_dev'] == pytest.approx(np.std(saved['all_optimums'], ddof=1))
        assert stats['avg_top_25%'] == pytest.approx(np.mean(sorted(saved['all_optimums'])[:100]))
    with pytest.raises(ValueError):
        run_simulation.main(['-m', 'sim(nope)', '-n', '10'])


def test_knownfn_problem_through_hyperband_api():
    from oracle import known_fn as okf
    odom = okf.MNIST_DOMAIN
    arms, curves = okf.synthetic_db(odom, 425, 300, seed=0)
    dom = Domain(**{p['name']: Param(p['name'], p['min_val'], p['max_val'], init_val=p['init_val'], distrib='uniform',
                                     scale=p['scale'], interval=p['interval']) for p in odom})
    known_fs = {Arm(**a): list(c) for a, c in zip(arms, curves)}

    class Real(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise NotImplementedError
    problem = _quiet(KnownFnProblem, known_fs, _quiet(Real, dom))
    opt = HyperbandOptimiser(eta=3, max_iter=81, optimisation_func=fval, is_simulation=True)
    best = _quiet(opt.run_optimisation, problem)
    assert len(opt.eval_history) == 12
    assert fval(best.optimisation_goals) == min(fval(e.optimisation_goals) for e in opt.eval_history)
    assert fval(best.optimisation_goals) in curves
    # the winning arm's nearest recorded curve really carries that loss
    again = problem.get_evaluator(arm=best.evaluator.arm).evaluate(81)
    assert min(again.fvals) <= fval(best.optimisation_goals) <= max(again.fvals)
    from autotune.experiments.monte_carlo import run_repetitions
    ofe = run_repetitions(opt, problem, 2000, seed=5)
    assert ofe.shape == (2000,) and ofe.dtype == torch.float64 and np.isin(ofe.cpu().numpy(), curves).all()


def test_closest_known_function_driver_cli_and_pickle(tmp_path):
    """experiments/run_closest_loss_fn_approximation.py: flags, synthetic fallback for the absent recorded arms, the
    reference's result schema (:207-217) - and the OFE distribution against the oracle's python loop on the same DB."""
    from scipy.stats import ks_2samp

    from autotune.experiments import run_closest_loss_fn_approximation as drv
    from oracle import known_fn as okf
    for problem, n_arms in (('mnist', 425), ('cifar', 300)):
        stats = _quiet(drv.main, ['-p', problem, '-m', 'sim(hb)', '-iter', '81', '-n', '3000', '-o', str(tmp_path), '--seed', '4',
                                  '--known-functions-dir', str(tmp_path / 'absent')])
        with open(tmp_path / 'known-fns-hist-sim(hb)-3000.pkl', 'rb') as f:
            saved = pickle.load(f)
        assert set(saved) == {'method', 'n_simulations', 'all_norm_optimums', 'avg_optimum', 'sample_std_dev', 'avg_top_25%'}
        assert saved['n_simulations'] == 3000 and len(saved['all_norm_optimums']) == 3000
        assert saved['avg_optimum'] == pytest.approx(np.mean(saved['all_norm_optimums']))
        assert stats['avg_top_25%'] == pytest.approx(np.mean(sorted(saved['all_norm_optimums'])[:750]))
        # every OFE is a recorded loss of the synthetic database the driver built
        args = drv._get_args(['-p', problem, '--seed', '4'])
        known = _quiet(drv.get_known_functions, args, str(tmp_path / 'absent'))
        assert len(known) == n_arms
        recorded = np.unique(np.asarray(list(known.values()), np.float64))
        assert np.isin(np.asarray(saved['all_norm_optimums'], np.float64), recorded).all()      # exact float64 losses
        # distribution vs the oracle's restatement of the reference loop on the same database
        odom = okf.MNIST_DOMAIN if problem == 'mnist' else okf.CIFAR_DOMAIN
        order = okf.attribute_order(odom)
        arms = [{n: a[n] for n in order} for a in known]
        curves = np.asarray(list(known.values()), np.float64)
        np.random.seed(11)
        want = [okf.run_hyperband(arms, curves, odom, 81, 3)['ofe'] for _ in range(60)]
        assert ks_2samp(np.asarray(saved['all_norm_optimums']), np.asarray(want)).pvalue > 0.01
    # the result pickles feed the comparison script of the reference (plot_run_known_functions_results.py)
    from autotune.experiments.simulation_evaluation import plot_run_known_functions_results as cmp_drv
    cmp_out = _quiet(cmp_drv.main, ['--dir', str(tmp_path), '-n', '3000'])
    assert cmp_out['labels'] == ['Hyperband'] and len(cmp_out['epdf']) == 3 and len(cmp_out['histograms'][0]) == 19
    with pytest.raises(NotImplementedError):
        _quiet(drv.main, ['-p', 'mnist', '-m', 'sim(hb+tpe+transfer+none)', '-n', '10', '-o', str(tmp_path)])
    with pytest.raises(ValueError):
        _quiet(drv.main, ['-p', 'nope', '-n', '10'])
    # test-function problems: 1000 random arms with constant curves (reference :118-123)
    stats = _quiet(drv.main, ['-p', 'branin', '-m', 'sim(hb)', '-iter', '6', '-n', '500', '-o', str(tmp_path)])
    assert stats['n_simulations'] == 500 and stats['avg_optimum'] > 0.39
