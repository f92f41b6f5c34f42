"""Loss-curve shape families and how they are dealt to arms (reference core/shape_family_scheduler.py:10-77)."""
import random
from abc import abstractmethod
from typing import Any, NamedTuple, Optional


class ShapeFamily(NamedTuple):
    arm: Optional[Any]
    ml_agg: float             # how hard a step bites into the remaining "function debt"
    necessary_agg: float      # how late the forced convergence to f_n sets in
    up_spikiness: float       # size of upward spikes
    is_smooth: bool = False   # Savitzky-Golay smoothing of the read-out
    start_shift: float = 0    # f(1) is shifted down by this (before noise)
    end_shift: float = 200    # f(n) is shifted down by this


class EvaluatorParams(NamedTuple):
    """ShapeFamily + (max_res, noise): the positional arguments of `SimulationProblem.get_evaluator`."""
    arm: Optional[Any]
    ml_agg: float
    necessary_agg: float
    up_spikiness: float
    is_smooth: bool = False
    start_shift: float = 0
    end_shift: float = 200
    max_res: int = 81
    noise: float = 0.3


class ShapeFamilyScheduler:

    def __init__(self, shape_families, max_resources, init_noise):
        self.shape_families, self.max_resources, self.init_noise = shape_families, max_resources, init_noise

    @abstractmethod
    def get_family(self, arm=None):
        pass

    def _params(self, family, arm):
        default_arm, *shape = family
        return EvaluatorParams(arm if arm is not None else default_arm, *shape, self.max_resources, self.init_noise)

    # id the batched device path keys its family draw on (autotune.b200.engine)
    device_kind = None


class RoundRobinShapeFamilyScheduler(ShapeFamilyScheduler):
    device_kind = 'round-robin'

    def __init__(self, shape_families, max_resources, init_noise):
        super().__init__(shape_families, max_resources, init_noise)
        self.index = 0

    def get_family(self, arm=None):
        family = self.shape_families[self.index % len(self.shape_families)]
        self.index += 1
        return self._params(family, arm)


class UniformShapeFamilyScheduler(ShapeFamilyScheduler):
    device_kind = 'uniform'

    def get_family(self, arm=None):
        return self._params(random.choice(self.shape_families), arm)
