"""Evaluation: (evaluator, optimisation_goals) pair (reference core/evaluation.py:7-14)."""
from typing import Any, NamedTuple


class Evaluation(NamedTuple):
    evaluator: Any            # holds the arm it evaluated
    optimisation_goals: Any
