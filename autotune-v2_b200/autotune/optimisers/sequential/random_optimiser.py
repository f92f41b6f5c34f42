"""Random search (reference optimisers/sequential/random_optimiser.py): max_iter random arms, each evaluated at
`n_resources`; on simulation problems one device launch, otherwise the user's evaluators are driven on the host."""
from autotune.core import Optimiser, optimisation_metric_user
from autotune.optimisers.sequential.tpe_optimiser import TpeOptimiser


class RandomOptimiser(TpeOptimiser):
    _tpe_enabled = False

    def __init__(self, n_resources, max_iter=None, max_time=None, min_or_max=min,
                 optimisation_func=Optimiser.default_optimisation_func, is_simulation=False, scheduler=None,
                 plot_simulation=False):
        super().__init__(n_resources, max_iter, max_time, min_or_max, optimisation_func, is_simulation, scheduler,
                         None, plot_simulation)

    @optimisation_metric_user
    def run_optimisation(self, problem, verbosity=False):
        if self.is_simulation:
            return super().run_optimisation(problem, verbosity)
        while not self._needs_to_stop():
            evaluator = problem.get_evaluator()
            opt_goals = evaluator.evaluate(self.n_resources)
            self._update_evaluation_history(evaluator, opt_goals)
            self._update_optimiser_metrics()
            if verbosity:
                self._print_evaluation(self.optimisation_func(opt_goals))
        return self._get_best_evaluation()
