"""Hyperband (reference optimisers/sequential/hyperband_optimiser.py).

On simulation problems (`is_simulation=True` with an OptFunctionSimulationProblem or a KnownFnProblem) a run is one
launch of the fused simulation + successive-halving kernel; on any other problem the optimiser drives the user's
evaluators round by round exactly as the reference does (that control flow is inherently host side).
"""
from math import ceil

import mpmath

from autotune.core import Evaluation, Optimiser, optimisation_metric_user
from autotune.optimisers import _device

_MP = mpmath.mp.clone()
_MP.dps = 64


def bracket_table(R, eta):
    """[(s, n, r)] per bracket with the reference's arithmetic: s_max from a 64-digit log (the float log gets 243 = 3**5
    wrong), brackets below s = 2 skipped, n = int(ceil(int(B/R/(s+1)) * eta**s)), r = R * eta**(-s)."""
    s_max = int(_MP.log(R) / _MP.log(eta))
    s_min = 2 if s_max >= 2 else 0
    B = (s_max + 1) * R
    return [(s, int(ceil(int(B / R / (s + 1)) * eta ** s)), R * eta ** (-s)) for s in reversed(range(s_min, s_max + 1))]


class HyperbandOptimiser(Optimiser):
    """Examples of resources: 1 resource = 10 000 training examples, or 1 epoch."""

    def __init__(self, eta, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 plot_simulation=False):
        super().__init__(max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler, plot_simulation)
        if max_iter is None:
            raise ValueError("For Hyperband max_iter cannot be None")
        self.eta = eta

    def _get_optimisation_func_val(self, evaluation):
        return self.optimisation_func(evaluation.optimisation_goals)

    def _get_best_n_evaluators(self, n, evaluations):
        """Stable sort (ascending for min, descending for max), first n evaluators."""
        ranked = sorted(evaluations, key=self._get_optimisation_func_val, reverse=self.min_or_max == max)
        return [e.evaluator for e in ranked[:n]]

    # ---- device path -------------------------------------------------------------------------------------------------
    _tpe_transfer = None      # None: plain Hyperband; hybrids set a transfer mode

    def _tpe_config(self):
        from autotune.b200.engine import tpe_config
        if self._tpe_transfer is None:
            return tpe_config(False)
        return tpe_config(True, self._tpe_transfer, getattr(self, 'n_res_transfer_threshold', None))

    def _run_on_device(self, problem):
        from autotune.b200.engine import run_knownfn_hyperband, run_optimiser_repetitions
        from autotune.benchmarks.known_loss_fn_problem import KnownFnProblem
        from autotune.benchmarks.opt_function_simulation_problem import OptFunctionSimulationProblem
        _device.require_fval_objective(self.optimisation_func)
        if isinstance(problem, KnownFnProblem):
            if self._tpe_transfer is not None:
                raise NotImplementedError('Hyperband-TPE hybrids on the closest-known-function problem are not on the '
                                          'device path yet')
            return self._run_knownfn(problem, run_knownfn_hyperband)
        if not isinstance(problem, OptFunctionSimulationProblem):
            raise NotImplementedError(f'{type(problem).__name__}: no device path for this simulation problem')
        spec = _device.simulation_spec(self, problem)
        seed = _device.fresh_seed()
        out = run_optimiser_repetitions(spec, 1, self._tpe_config(), seed=seed, details=True)
        return _device.history_from_rounds(self, problem, spec, out, seed)

    def _run_knownfn(self, problem, run_knownfn_hyperband):
        from autotune.core import Arm, OptimisationGoals
        db = problem.device_db()
        out = run_knownfn_hyperband(db, self.max_iter, self.eta, 1, seed=_device.fresh_seed(),
                                    min_or_max=self.min_or_max, details=True)
        # details give the nearest index of every arm; rebuild the round-best evaluations from them
        nn, rb, rb_arm = out['nn_idx'][0].tolist(), out['round_best'][0].tolist(), out['round_best_arm'][0].tolist()
        best_vals = out['best_arm_vals'][0].tolist()
        best_arm_id = int(out['ofe_arm'][0])
        for loss, arm_id in zip(rb, rb_arm):
            curve = problem.known_fs[db.arms[nn[arm_id]]]
            arm = Arm(**dict(zip(db.names, best_vals))) if arm_id == best_arm_id else db.arms[nn[arm_id]]
            evaluator = problem.get_evaluator(arm=arm)
            self.eval_history.append(Evaluation(evaluator, OptimisationGoals(fval=loss, test_error=-1,
                                                                             validation_error=-1, fvals=curve)))
            self._update_optimiser_metrics()
        return self._get_best_evaluation()

    # ---- host control flow for problems whose evaluators are user code ------------------------------------------------
    def _new_evaluators(self, problem, n, bracket_index, r):
        return [problem.get_evaluator() for _ in range(n)], None

    def _run_on_host(self, problem, verbosity):
        for bracket_index, (s, n, r) in enumerate(bracket_table(self.max_iter, self.eta)):
            evaluators, evaluations = self._new_evaluators(problem, n, bracket_index, r)
            print(f"\n{'=' * 73}\n>> Generated {n} evaluators each with a random arm")
            for i in range(s + 1):
                n_i = n * self.eta ** (-i)
                r_i = r * self.eta ** i
                if evaluations is None:
                    evaluations = [Evaluation(ev, ev.evaluate(n_resources=r_i)) for ev in evaluators]
                print(f"** Evaluated {int(n_i)} arms, each with {r_i:.2f} resources")
                evaluators = self._get_best_n_evaluators(n=int(n_i / self.eta), evaluations=evaluations)
                best_in_round = self.min_or_max(evaluations, key=self._get_optimisation_func_val)
                self._update_evaluation_history(*best_in_round)
                self._update_optimiser_metrics()
                if verbosity:
                    self._print_evaluation(self.optimisation_func(best_in_round.optimisation_goals))
                evaluations = None
        return self._get_best_evaluation()

    @optimisation_metric_user
    def run_optimisation(self, problem, verbosity=False):
        """Run Hyperband and return the Evaluation of the best arm; `eval_history` holds the best of every round."""
        if self.is_simulation:
            return self._run_on_device(problem)
        return self._run_on_host(problem, verbosity)

    def __str__(self):
        return (f"\n> Starting Hyperband optimisation\n    Max iterations (R)      = {self.max_iter}\n"
                f"    Halving rate (eta)      = {self.eta}\n"
                f"  Optimizing ({self.min_or_max.__name__}) {self.optimisation_func.__doc__}")
