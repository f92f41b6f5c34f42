from autotune.core.arm import Arm
from autotune.core.evaluation import Evaluation
from autotune.core.evaluator import Evaluator
from autotune.core.hyperparams_domain import Domain
from autotune.core.model_builder import ModelBuilder
from autotune.core.optimisation_goals import OptimisationGoals
from autotune.core.optimiser import Optimiser, optimisation_metric_user
from autotune.core.params import Param
from autotune.core.problem_def import HyperparameterOptimisationProblem, SimulationProblem
from autotune.core.shape_family_scheduler import (
    EvaluatorParams, RoundRobinShapeFamilyScheduler, ShapeFamily, ShapeFamilyScheduler, UniformShapeFamilyScheduler)
from autotune.core.simulation_evaluator import SimulationEvaluator

__all__ = [
    'HyperparameterOptimisationProblem', 'SimulationProblem', 'Param', 'Arm', 'Domain', 'ModelBuilder', 'Evaluator',
    'OptimisationGoals', 'Evaluation', 'Optimiser', 'optimisation_metric_user', 'ShapeFamilyScheduler',
    'RoundRobinShapeFamilyScheduler', 'ShapeFamily', 'EvaluatorParams', 'UniformShapeFamilyScheduler',
    'SimulationEvaluator']
