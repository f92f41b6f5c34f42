"""Stand-alone TPE (reference optimisers/sequential/tpe_optimiser.py, a wrapper over hyperopt.fmin there).

On simulation problems the max_iter sequential suggestions + evaluations at `n_resources` run fused on the device.
"""
from autotune.core import Evaluation, OptimisationGoals, Optimiser, optimisation_metric_user
from autotune.optimisers import _device


class TpeOptimiser(Optimiser):
    _tpe_enabled = True

    def __init__(self, n_resources, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 trials_to_inject=None, plot_simulation=False):
        super().__init__(max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler, plot_simulation)
        if trials_to_inject is not None:
            raise NotImplementedError('injecting hyperopt Trials is internal to the hybrids on the device path')
        self.n_injected_trials = 0
        self.sign = -1 if min_or_max == max else 1
        self.n_resources = n_resources

    @optimisation_metric_user
    def run_optimisation(self, problem, verbosity=False):
        from autotune.b200.engine import run_optimiser_repetitions, single_round_plan, tpe_config
        from autotune.benchmarks.opt_function_simulation_problem import OptFunctionSimulationProblem
        from autotune.core import Arm
        if not self.is_simulation or not isinstance(problem, OptFunctionSimulationProblem):
            raise NotImplementedError(f'{type(self).__name__} runs on OptFunctionSimulationProblem simulations only')
        if self.max_iter is None or self.max_iter == float('inf'):
            raise ValueError('max_iter is required on the device path')
        _device.require_fval_objective(self.optimisation_func)
        spec = _device.simulation_spec(self, problem)
        plan = single_round_plan(spec.R, int(self.max_iter), int(self.n_resources))
        from autotune.b200.engine import arm_families
        seed = _device.fresh_seed()
        out = run_optimiser_repetitions(spec, 1, tpe_config(self._tpe_enabled), seed=seed, details=True, plan=plan)
        losses = out['ckpt_loss'][0, :, 0].tolist()
        xy = out['arm_xy'][0].tolist()
        fam_of = arm_families(spec, seed, 0, range(len(losses))).tolist()   # the family each arm drew on the device
        for (x, y), loss, fam_id in zip(xy, losses, fam_of):   # every evaluation goes into the history (tpe_optimiser.py:56-58)
            evaluator = problem.get_evaluator(Arm(x=x, y=y), *spec.families[fam_id], spec.R, spec.init_noise)
            self.eval_history.append(Evaluation(evaluator, OptimisationGoals(fval=loss, test_error=-1, validation_error=-1)))
            self._update_optimiser_metrics()
        return self._get_best_evaluation()
