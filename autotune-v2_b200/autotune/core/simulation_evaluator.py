"""SimulationEvaluator: one arm's Gamma-process loss trajectory (reference core/simulation_evaluator.py:51-116).

The trajectory is produced on the GPU (`torch.ops.at2.sim_trajectories`, Philox + Marsaglia-Tsang + the recurrence of
simulation_evaluator.py:97-116); this class only keeps the reference's attribute surface (`non_smooth_fs`, `fs`,
`simulate`).  Randomness is the library's counter-based Philox keyed by a per-evaluator id, not numpy's global RNG.
"""
import itertools

_EVALUATOR_IDS = itertools.count()


class SimulationEvaluator:

    def __init__(self, ml_aggressiveness, necessary_aggressiveness, up_spikiness, max_resources, is_smooth=True):
        self.max_resources = max_resources
        self.non_smooth_fs = []
        self.is_smooth = is_smooth
        self.ml_aggressiveness = ml_aggressiveness
        self.necessary_aggressiveness = necessary_aggressiveness
        self.up_spikiness = up_spikiness
        self._stream_id = next(_EVALUATOR_IDS)   # Philox stream of this evaluator
        self._draws = 0                          # bumps every time the trajectory is (re)simulated
        self._smooth_fs = None

    @property
    def fs(self):
        """Raw trajectory, or its Savitzky-Golay smoothing (window int(0.17*R+6) made odd, cubic) when is_smooth."""
        if not self.is_smooth:
            return self.non_smooth_fs
        if self._smooth_fs is None:
            from autotune.b200 import trajectories
            self._smooth_fs = trajectories.smooth_one(self.non_smooth_fs, self.max_resources)
        return self._smooth_fs

    def simulate(self, time, n, f_n):
        """Extend `non_smooth_fs` (which must already hold f_1) up to `time` points of an n-point trajectory."""
        if len(self.non_smooth_fs) >= time:
            return
        if len(self.non_smooth_fs) != 1 or time != n:
            raise NotImplementedError('the device path simulates whole trajectories: call simulate(n, n, f_n) on [f_1]')
        from autotune.b200 import trajectories
        self.non_smooth_fs = trajectories.simulate_one(
            f_1=self.non_smooth_fs[0], f_n=f_n, n=n, ml_aggressiveness=self.ml_aggressiveness,
            necessary_aggressiveness=self.necessary_aggressiveness, up_spikiness=self.up_spikiness,
            stream_id=self._stream_id, draw=self._draws)
        self._draws += 1
        self._smooth_fs = None
