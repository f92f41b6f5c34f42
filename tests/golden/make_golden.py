#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What it does
------------
The reference (`/root/reference`, pure Python) is imported with the presentation-only stand-ins in
`oracle/ref_stubs/` (colorama / matplotlib / hyperopt / sigopt names; none touches arithmetic).  A repetition of
`HyperbandOptimiser.run_optimisation` is then made deterministic *from the outside* (no reference edits):

  * a `ShapeFamilyScheduler` subclass hands out `Arm(x, y)` + family for arm ids 0,1,2,... in creation order;
  * `OptFunctionSimulationEvaluator.evaluate` is wrapped to note which arm is being evaluated;
  * `numpy.random.randn` returns that arm's pre-drawn `z` (`opt_function_problem.py:138`);
  * `simulation_evaluator.get_aggressiveness_from_gamma_distrib` returns `gamma[arm, t-1]`
    (`simulation_evaluator.py:36-48,98`) - the same values for the discarded first pass and the effective second
    pass of arms first evaluated at r=1 (`opt_function_simulation_problem.py:67-74`);
  * `HyperbandOptimiser._get_best_n_evaluators` is wrapped to record every round's evaluation list and survivors.

Outputs are what the reference itself computed (float64): full raw and smoothed trajectories, checkpoint losses,
per-round evaluation lists, survivor ids in sorted order, round-bests, the OFE and the winning arm.

Also written: `bracket_tables.npz` (n, r_i, keep counts captured from reference Hyperband runs on a constant-loss
problem, cf. `tests/hyperband_test.py`) and `epdf_ofes.npz` (the `all_optimums` vectors of the reference's stored
EPDF-OFE pickles, used as distributional known-answer tests).
"""
import io
import os
import pickle
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('AT2_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))

import autotune.core.simulation_evaluator as ref_sim  # noqa: E402
from autotune.benchmarks import OptFunctionSimulationProblem  # noqa: E402
from autotune.benchmarks.opt_function_problem import DOMAINS  # noqa: E402
from autotune.benchmarks.opt_function_simulation_problem import OptFunctionSimulationEvaluator  # noqa: E402
from autotune.core import (  # noqa: E402
    Arm, Domain, Evaluator, EvaluatorParams, HyperparameterOptimisationProblem, OptimisationGoals, Optimiser, Param,
    ShapeFamily, ShapeFamilyScheduler)
from autotune.optimisers import HyperbandOptimiser  # noqa: E402

FAMILIES_GENERAL6 = (  # run_simulation.py:41-47
    (1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
    (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400))
FAMILIES_EGG = (  # run_simulation.py:36-40
    (1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000))
FAMILIES_FLAT = ((0, 1, 0, False, 0, 0),)  # run_simulation.py:49 (commented "flat")
# R=6 cannot take a smooth family: the reference's savgol window int(0.17*6+6)=7 exceeds the 6 points -> ValueError
FAMILIES_NONSMOOTH4 = tuple(f for f in FAMILIES_GENERAL6 if not f[3])


def fval(goals):
    """fval."""
    return goals.fval


class PreDrawnScheduler(ShapeFamilyScheduler):
    def __init__(self, families, max_resources, init_noise, xy, fam):
        super().__init__(tuple(ShapeFamily(None, *f) for f in families), max_resources, init_noise)
        self.xy, self.fam, self.next_id = xy, fam, 0

    def get_family(self, arm=None):
        i = self.next_id
        self.next_id += 1
        _, *rest = self.shape_families[int(self.fam[i])]
        return EvaluatorParams(Arm(x=float(self.xy[i, 0]), y=float(self.xy[i, 1])), *rest,
                               self.max_resources, self.init_noise)


def draw_inputs(rng, func, n_arms, R, n_families):
    """Pre-drawn inputs of one repetition (float64).  gamma[a, t-1] ~ Gamma(alpha_t, 1/beta_t), t = 1..R-1, with
    (alpha_t, beta_t) as in simulation_evaluator.py:37-42 called with n = R + 1, k = 2."""
    dom = DOMAINS[func]
    xy = np.stack([rng.uniform(dom.x.min_val, dom.x.max_val, n_arms),
                   rng.uniform(dom.y.min_val, dom.y.max_val, n_arms)], axis=1)
    fam = rng.integers(0, n_families, n_arms).astype(np.int32)
    z = rng.standard_normal(n_arms)
    t = np.arange(1, R)
    m = (R + 1) - t
    beta = (2 + np.sqrt(2 ** 2 + 4 * m)) / (2 * m)
    alpha = 2 * beta + 1
    gamma = rng.gamma(shape=alpha[None, :], scale=(1 / beta)[None, :], size=(n_arms, R - 1))
    return xy, fam, z, gamma


def run_reference_repetition(func, families, R, eta, noise, xy, fam, z, gamma, min_or_max=min):
    """One full reference Hyperband run on pre-drawn inputs; returns everything it computed."""
    state = {'cur': -1, 'created': [], 'res': set()}
    rounds = []

    orig_get_eval = OptFunctionSimulationProblem.get_evaluator
    orig_evaluate = OptFunctionSimulationEvaluator.evaluate
    orig_best_n = HyperbandOptimiser._get_best_n_evaluators
    orig_randn = np.random.randn
    orig_gamma_fn = ref_sim.get_aggressiveness_from_gamma_distrib

    def get_evaluator(self, *a, **k):
        ev = orig_get_eval(self, *a, **k)
        ev.golden_id = len(state['created'])
        state['created'].append(ev)
        return ev

    def evaluate(self, n_resources):
        state['cur'] = self.golden_id
        state['res'].add(int(n_resources))
        return orig_evaluate(self, n_resources)

    def best_n(self, n, evaluations):
        out = orig_best_n(self, n, evaluations)
        rounds.append(dict(ids=[e.evaluator.golden_id for e in evaluations],
                           fvals=[e.optimisation_goals.fval for e in evaluations],
                           keep=n, surv=[ev.golden_id for ev in out]))
        return out

    OptFunctionSimulationProblem.get_evaluator = get_evaluator
    OptFunctionSimulationEvaluator.evaluate = evaluate
    HyperbandOptimiser._get_best_n_evaluators = best_n
    np.random.randn = lambda *a: np.array([z[state['cur']]])
    ref_sim.get_aggressiveness_from_gamma_distrib = lambda t, n, k: float(gamma[state['cur'], t - 1])
    try:
        with redirect_stdout(io.StringIO()):
            problem = OptFunctionSimulationProblem(func)
            sched = PreDrawnScheduler(families, R, noise, xy, fam)
            opt = HyperbandOptimiser(eta=eta, max_iter=R, min_or_max=min_or_max, optimisation_func=fval,
                                     is_simulation=True, scheduler=sched)
            best = opt.run_optimisation(problem, verbosity=False)
    finally:
        OptFunctionSimulationProblem.get_evaluator = orig_get_eval
        OptFunctionSimulationEvaluator.evaluate = orig_evaluate
        HyperbandOptimiser._get_best_n_evaluators = orig_best_n
        np.random.randn = orig_randn
        ref_sim.get_aggressiveness_from_gamma_distrib = orig_gamma_fn

    evs = state['created']
    assert len(evs) == len(z), (len(evs), len(z))
    traj_raw = np.array([ev.non_smooth_fs for ev in evs], dtype=np.float64)
    traj = np.array([ev.fs for ev in evs], dtype=np.float64)
    ckpts = np.asarray(sorted({int(c) for c in state['res']}), np.int32)
    ckpt_loss = traj[:, ckpts - 1]
    ids = np.concatenate([np.asarray(r['ids'], np.int32) for r in rounds])
    fvals = np.concatenate([np.asarray(r['fvals'], np.float64) for r in rounds])
    round_off = np.cumsum([0] + [len(r['ids']) for r in rounds]).astype(np.int32)
    surv = np.concatenate([np.asarray(r['surv'], np.int32) for r in rounds]) if rounds else np.zeros(0, np.int32)
    surv_off = np.cumsum([0] + [len(r['surv']) for r in rounds]).astype(np.int32)
    keep = np.asarray([r['keep'] for r in rounds], np.int32)
    round_best = np.asarray([fval(e.optimisation_goals) for e in opt.eval_history], np.float64)
    round_best_id = np.asarray([e.evaluator.golden_id for e in opt.eval_history], np.int32)
    return dict(traj_raw=traj_raw, traj=traj, ckpts=ckpts, ckpt_loss=ckpt_loss, eval_ids=ids, eval_fvals=fvals, round_off=round_off, surv_ids=surv,
                surv_off=surv_off, keep=keep, round_best=round_best, round_best_id=round_best_id,
                ofe=np.float64(fval(best.optimisation_goals)), ofe_id=np.int32(best.evaluator.golden_id))


def arms_per_repetition(R, eta):
    """Number of arms a reference Hyperband run creates (captured by actually running it on a constant problem)."""
    return int(sum(reference_bracket_table(R, eta)['n']))


class _ConstEvaluator(Evaluator):
    calls = []

    def evaluate(self, n_resources):
        _ConstEvaluator.calls.append(n_resources)
        return OptimisationGoals(validation_error=1, test_error=1)

    def _train(self, *a, **k):
        pass

    def _test(self, *a, **k):
        pass

    def _save_checkpoint(self, *a, **k):
        pass


class _Builder:
    def __init__(self, arm):
        self.arm = arm


class _ConstProblem(HyperparameterOptimisationProblem):
    def __init__(self):
        super().__init__(Domain(x=Param('x', 0, 1, distrib='uniform', scale='linear')))

    def get_evaluator(self, arm=None):
        arm = Arm()
        arm.draw_hp_val(domain=self.domain, hyperparams_to_opt=self.hyperparams_to_opt)
        return _ConstEvaluator(_Builder(arm))


_TABLE_CACHE = {}


def reference_bracket_table(R, eta):
    """Bracket table as the reference executes it: per bracket n; per round the n_resources float it passes to
    evaluate(), the number of evaluations and the keep count (hyperband_optimiser.py:69-97)."""
    if (R, eta) in _TABLE_CACHE:
        return _TABLE_CACHE[(R, eta)]
    recs = []
    orig_best_n = HyperbandOptimiser._get_best_n_evaluators

    def best_n(self, n, evaluations):
        recs.append((len(evaluations), n, _ConstEvaluator.calls[-1]))
        return orig_best_n(self, n, evaluations)

    HyperbandOptimiser._get_best_n_evaluators = best_n
    _ConstEvaluator.calls = []
    try:
        with redirect_stdout(io.StringIO()):
            opt = HyperbandOptimiser(eta=eta, max_iter=R, min_or_max=min,
                                     optimisation_func=Optimiser.default_optimisation_func)
            opt.run_optimisation(_ConstProblem(), verbosity=False)
    finally:
        HyperbandOptimiser._get_best_n_evaluators = orig_best_n
    n_eval = np.asarray([r[0] for r in recs], np.int32)
    keep = np.asarray([r[1] for r in recs], np.int32)
    res = np.asarray([r[2] for r in recs], np.float64)
    # a bracket starts wherever the evaluation count is not the previous keep count
    starts = [0] + [i for i in range(1, len(recs)) if n_eval[i] != keep[i - 1]]
    # (equal consecutive sizes cannot be told apart that way; disambiguate by the resource going down)
    starts = sorted(set(starts) | {i for i in range(1, len(recs)) if res[i] <= res[i - 1]})
    out = dict(n=n_eval[starts], round_n=n_eval, round_keep=keep, round_res=res,
               bracket_start=np.asarray(starts + [len(recs)], np.int32))
    _TABLE_CACHE[(R, eta)] = out
    return out


CASES = [
    # name, func, families, R, eta, noise, n_reps, seed, min_or_max
    ('rastrigin6_R81', 'rastrigin', FAMILIES_GENERAL6, 81, 3, 10, 6, 0, 'min'),
    ('branin_flat_R81', 'branin', FAMILIES_FLAT, 81, 3, 0, 2, 1, 'min'),
    ('wave_flat_R81', 'wave', FAMILIES_FLAT, 81, 3, 0, 2, 2, 'min'),
    ('egg3_R243', 'egg', FAMILIES_EGG, 243, 3, 10, 1, 3, 'min'),
    ('camel6_R27', 'camel', FAMILIES_GENERAL6, 27, 3, 10, 3, 4, 'min'),
    ('rastrigin6_R27_max', 'rastrigin', FAMILIES_GENERAL6, 27, 3, 10, 2, 5, 'max'),
    ('branin4_R6', 'branin', FAMILIES_NONSMOOTH4, 6, 3, 10, 3, 6, 'min'),
    ('wave6_R16_eta2', 'wave', FAMILIES_GENERAL6, 16, 2, 1, 3, 7, 'min'),
    ('rastrigin6_R100', 'rastrigin', FAMILIES_GENERAL6, 100, 3, 10, 1, 8, 'min'),
]


def make_hyperband_cases():
    for name, func, families, R, eta, noise, n_reps, seed, mom in CASES:
        rng = np.random.default_rng(seed)
        A = arms_per_repetition(R, eta)
        out = dict(func=np.array(func), families=np.asarray(families, np.float64), R=np.int32(R), eta=np.int32(eta),
                   noise=np.float64(noise), min_or_max=np.array(mom), n_reps=np.int32(n_reps), n_arms=np.int32(A))
        for rep in range(n_reps):
            xy, fam, z, gamma = draw_inputs(rng, func, A, R, len(families))
            if name == 'branin_flat_R81' and rep == 1:
                xy[:] = xy[0]  # constant-loss repetition: the tie-break case of tests/hyperband_test.py
            res = run_reference_repetition(func, families, R, eta, noise, xy, fam, z, gamma,
                                           min_or_max=min if mom == 'min' else max)
            if rep > 0:
                res.pop('traj')  # full smoothed trajectories only for the first repetition (fixture size)
            for k, v in dict(xy=xy, fam=fam, z=z, gamma=gamma, **res).items():
                out[f'rep{rep}_{k}'] = v
        path = os.path.join(HERE, f'hb_{name}.npz')
        np.savez_compressed(path, **out)
        print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_bracket_tables():
    out = {}
    for R, eta in [(6, 3), (9, 3), (16, 2), (27, 3), (64, 4), (81, 3), (100, 3), (125, 5), (243, 3), (256, 2),
                   (729, 3), (1000, 10), (1121, 3), (2187, 3)]:
        tab = reference_bracket_table(R, eta)
        for k, v in tab.items():
            out[f'R{R}_eta{eta}_{k}'] = v
    path = os.path.join(HERE, 'bracket_tables.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


def make_epdf_vectors():
    out = {}
    base = os.path.join(REF, 'epdf_ofes')
    for d in sorted(os.listdir(base)):
        for f in sorted(os.listdir(os.path.join(base, d))):
            with open(os.path.join(base, d, f), 'rb') as fh:
                obj = pickle.load(fh)
            key = 'all_optimums' if 'all_optimums' in obj else 'all_norm_optimums'
            out[f'{d}/{f[:-4]}'] = np.asarray(obj[key], np.float64)
    path = os.path.join(HERE, 'epdf_ofes.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', len(out), 'vectors')


SEEDED = [
    # name, func, families, R, eta, noise, scheduler, seed, n_reps, min_or_max
    ('rastrigin6_R81_uniform', 'rastrigin', FAMILIES_GENERAL6, 81, 3, 10, 'uniform', 1000, 5, 'min'),
    ('branin_flat_R81_uniform', 'branin', FAMILIES_FLAT, 81, 3, 0, 'uniform', 1001, 5, 'min'),
    ('egg3_R243_uniform', 'egg', FAMILIES_EGG, 243, 3, 10, 'uniform', 1002, 1, 'min'),
    ('wave6_R27_roundrobin', 'wave', FAMILIES_GENERAL6, 27, 3, 10, 'round-robin', 1003, 4, 'min'),
    ('camel6_R27_uniform_max', 'camel', FAMILIES_GENERAL6, 27, 3, 10, 'uniform', 1004, 4, 'max'),
]


def make_seeded_ofes():
    """OFEs of the reference's own repetition loop (run_simulation.py:147-157: fresh problem + optimiser per
    repetition) under np.random.seed(seed); random.seed(seed) - pins the RNG consumption order end to end."""
    import random

    from autotune.core import RoundRobinShapeFamilyScheduler, UniformShapeFamilyScheduler
    out = {}
    for name, func, families, R, eta, noise, sched, seed, n_reps, mom in SEEDED:
        np.random.seed(seed)
        random.seed(seed)
        ofes = []
        with redirect_stdout(io.StringIO()):
            for _ in range(n_reps):
                problem = OptFunctionSimulationProblem(func)
                cls = UniformShapeFamilyScheduler if sched == 'uniform' else RoundRobinShapeFamilyScheduler
                scheduler = cls(shape_families=tuple(ShapeFamily(None, *f) for f in families), max_resources=R,
                                init_noise=noise)
                opt = HyperbandOptimiser(eta=eta, max_iter=R, min_or_max=min if mom == 'min' else max,
                                         optimisation_func=fval, is_simulation=True, scheduler=scheduler)
                best = opt.run_optimisation(problem, verbosity=True)
                ofes.append(fval(best.optimisation_goals))
        out[name] = np.asarray(ofes, np.float64)
        out[name + '_cfg'] = np.array(repr(dict(func=func, families=families, R=R, eta=eta, noise=noise,
                                                scheduler=sched, seed=seed, n_reps=n_reps, min_or_max=mom)))
    path = os.path.join(HERE, 'seeded_ofes.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


def make_knownfn():
    """Closest-known-loss-function goldens from the reference's own KnownFnProblem / KnownFnEvaluator on a synthetic
    recorded-arm database (the real one is git-ignored upstream): nearest-arm index + loss for individual queries,
    and seeded Hyperband repetitions (R=81 and the script default R=6)."""
    import random

    from autotune.benchmarks import KnownFnProblem
    from autotune.benchmarks.cifar_problem import HYPERPARAMETERS_TO_OPTIMIZE as CIFAR_OPT
    from autotune.benchmarks.cifar_problem import HYPERPARAMS_DOMAIN as CIFAR_DOM
    from autotune.benchmarks.known_loss_fn_problem import KnownFnEvaluator
    from autotune.benchmarks.mnist_problem import HYPERPARAMS_DOMAIN as MNIST_DOM

    sys.path.insert(0, ROOT)
    from oracle import known_fn as okf

    class RealProblem(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise NotImplementedError

    out = {}
    for tag, dom, to_opt, odom, n_arms in [('mnist', MNIST_DOM, (), okf.MNIST_DOMAIN, 425),
                                           ('cifar', CIFAR_DOM, CIFAR_OPT, okf.CIFAR_DOMAIN, 300)]:
        with redirect_stdout(io.StringIO()):
            real = RealProblem(dom, to_opt)
        # hyperparams_to_opt goes through a set() upstream; the draw order is the domain order regardless
        db_dicts, curves = okf.synthetic_db(odom, n_arms, 300, seed=0)
        names = okf.attribute_order(odom)
        known_fs = {Arm(**{n: d[n] for n in names}): list(c) for d, c in zip(db_dicts, curves)}
        assert len(known_fs) == n_arms
        out[f'{tag}_db'] = np.asarray([[float(d[n]) for n in names] for d in db_dicts], np.float64)
        out[f'{tag}_names'] = np.array(names)
        # (1) individual queries through the reference evaluator
        np.random.seed(7)
        with redirect_stdout(io.StringIO()):
            problem = KnownFnProblem(known_fs=known_fs, real_problem=real)
        q_vals, q_idx, q_loss = [], [], []
        keys = list(known_fs.keys())
        for k in range(200):
            with redirect_stdout(io.StringIO()):
                ev = problem.get_evaluator()
                goals = ev.evaluate(n_resources=(1, 3, 9, 27, 81)[k % 5])
            assert list(ev.arm.__dict__.keys()) == names
            q_vals.append([float(ev.arm[n]) for n in names])
            q_loss.append(goals.fval)
            q_idx.append(next(i for i, a in enumerate(keys) if known_fs[a] is goals.fvals))
        out[f'{tag}_query'] = np.asarray(q_vals, np.float64)
        out[f'{tag}_query_idx'] = np.asarray(q_idx, np.int32)
        out[f'{tag}_query_loss'] = np.asarray(q_loss, np.float64)
        # (2) seeded Hyperband repetitions
        for R in (81, 6):
            np.random.seed(100 + R)
            random.seed(100 + R)
            ofes, arms_all, best_ids = [], [], []
            n_reps = 3 if R == 81 else 6
            for _ in range(n_reps):
                created = []
                orig = KnownFnProblem.get_evaluator

                def get_evaluator(self, *a, **k):
                    ev = orig(self, *a, **k)
                    ev.golden_id = len(created)
                    created.append(ev)
                    return ev
                KnownFnProblem.get_evaluator = get_evaluator
                try:
                    with redirect_stdout(io.StringIO()):
                        problem = KnownFnProblem(known_fs=known_fs, real_problem=real)
                        opt = HyperbandOptimiser(eta=3, max_iter=R, min_or_max=min, optimisation_func=fval,
                                                 is_simulation=True)
                        best = opt.run_optimisation(problem, verbosity=False)
                finally:
                    KnownFnProblem.get_evaluator = orig
                ofes.append(fval(best.optimisation_goals))
                best_ids.append(best.evaluator.golden_id)
                arms_all.append([[float(ev.arm[n]) for n in names] for ev in created])
            out[f'{tag}_hb_R{R}_ofe'] = np.asarray(ofes, np.float64)
            out[f'{tag}_hb_R{R}_ofe_id'] = np.asarray(best_ids, np.int32)
            out[f'{tag}_hb_R{R}_arms'] = np.asarray(arms_all, np.float64)
    path = os.path.join(HERE, 'knownfn.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_postprocessing():
    """Goldens of the post-processing layer: loss-curve family profiles (simulation_evaluation/profiles.py) and the k-fold
    cross-validation of the closest-known-function approximation (known_functions_evaluation/cross_validate_approximation.py),
    both computed by the reference's own functions."""
    from autotune.benchmarks.mnist_problem import HYPERPARAMS_DOMAIN as MNIST_DOM
    from autotune.experiments.known_functions_evaluation.cross_validate_approximation import cross_validate
    from autotune.experiments.simulation_evaluation import profiles as ref_prof

    sys.path.insert(0, ROOT)
    from oracle import known_fn as okf
    out = {}
    rng = np.random.default_rng(42)
    base = np.linspace(100, -100, 27)[None, :]
    curves = (base + rng.normal(0, 25, (40, 27))).astype(np.float32).astype(np.float64)
    curves[3] = curves[5]                                   # exact ties must count as neither less nor greater
    out['prof_curves'] = curves
    out['prof_avg'] = ref_prof.get_avg_profile(curves)
    out['prof_med'] = ref_prof.get_med_profile(curves)
    out['prof_std'] = ref_prof.get_std_profile(curves)
    out['prof_dynamic'] = np.asarray(ref_prof.get_dynamic_order_profile(curves))
    out['prof_ends'] = np.float64(ref_prof.get_ends_order_profile(curves))

    class RealProblem(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise NotImplementedError
    with redirect_stdout(io.StringIO()):
        real = RealProblem(MNIST_DOM, ())
        db_dicts, db_curves = okf.synthetic_db(okf.MNIST_DOMAIN, 103, 60, seed=1)   # 103 arms: a ragged last fold
        names = okf.attribute_order(okf.MNIST_DOMAIN)
        known_fs = {Arm(**{n: d[n] for n in names}): list(c) for d, c in zip(db_dicts, db_curves)}
        res, min_len = cross_validate(10, known_fs, real)
    out['cv_res'] = res
    out['cv_min_len'] = np.int32(min_len)
    path = os.path.join(HERE, 'postprocessing.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', res.shape)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    make_seeded_ofes()
    make_knownfn()
    make_postprocessing()
    make_bracket_tables()
    make_hyperband_cases()
    make_epdf_vectors()
