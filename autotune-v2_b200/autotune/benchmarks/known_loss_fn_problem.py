"""Closest-known-loss-function approximation (reference benchmarks/known_loss_fn_problem.py).

A proposed arm is "evaluated" by looking up the recorded arm nearest to it (squared error of the hyperparameters
normalised by the domain) and reading that arm's recorded loss curve.  The scan runs on the GPU
(`torch.ops.at2.nn_lookup`, float64, first minimum wins - the nearest index is bit-identical to the reference's).
"""
from autotune.core import Arm, HyperparameterOptimisationProblem, OptimisationGoals, SimulationEvaluator, SimulationProblem


class KnownFnEvaluator(SimulationEvaluator):

    def __init__(self, known_fs, proposed_arm, domain, should_plot=False, _device_db=None):
        super().__init__(0, 0, 0, 0)
        self.should_plot = should_plot
        self.known_fs = known_fs
        self.proposed_arm = self.arm = proposed_arm
        self.domain = domain
        self._device_db = _device_db

    def _db(self):
        if self._device_db is None:
            from autotune.b200.engine import KnownFnSpec
            self._device_db = KnownFnSpec(self.known_fs, self.domain, tuple(self.domain.hyperparams_names()))
        return self._device_db

    def evaluate(self, n_resources):
        """Loss of the nearest recorded arm's curve at int(n_resources) (known_loss_fn_problem.py:22-62)."""
        time = int(n_resources)
        db = self._db()
        names = list(self.proposed_arm.__dict__.keys())
        if set(names) != set(db.names):
            raise KeyError(f'the proposed arm must define exactly the domain hyperparameters {db.names}')
        # the reference sums the squared-error terms in the attribute order of the proposed arm (:36-48), whatever it is
        idx, _ = db.lookup([[float(self.proposed_arm[n]) for n in db.names]], order=[db.names.index(n) for n in names])
        nearest = db.arms[int(idx[0])]
        curve = self.known_fs[nearest]
        return OptimisationGoals(fval=curve[time - 1], test_error=-1, validation_error=-1, fvals=curve)


class KnownFnProblem(HyperparameterOptimisationProblem, SimulationProblem):

    def __init__(self, known_fs, real_problem):
        HyperparameterOptimisationProblem.__init__(self, real_problem.domain, real_problem.hyperparams_to_opt)
        SimulationProblem.__init__(self)
        self.known_fs = known_fs
        self._device_db = None

    def device_db(self):
        """The recorded arms and curves as device tensors (uploaded once per problem)."""
        if self._device_db is None:
            from autotune.b200.engine import KnownFnSpec
            self._device_db = KnownFnSpec(self.known_fs, self.domain, self.hyperparams_to_opt)
        return self._device_db

    def get_evaluator(self, arm=None, should_plot=False):
        if arm is None:
            arm = Arm()
            arm.draw_hp_val(domain=self.domain, hyperparams_to_opt=self.hyperparams_to_opt)
        return KnownFnEvaluator(known_fs=self.known_fs, proposed_arm=arm, should_plot=should_plot, domain=self.domain,
                                _device_db=self.device_db())
