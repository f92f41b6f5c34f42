"""Evaluator base class (reference core/evaluator.py:15-89) - the object an optimiser calls `.evaluate(n_resources)` on."""
from abc import abstractmethod
from pathlib import Path

import pandas as pd


class Evaluator:

    def __init__(self, model_builder, output_dir=None, file_name='model.pth'):
        self.output_dir = output_dir
        if output_dir is not None:
            stamp = str(pd.Timestamp.utcnow())
            for ch in ': .+':
                stamp = stamp.replace(ch, '-')
            self.directory = Path(output_dir) / f'arm-{stamp}'
            self.directory.mkdir(parents=True, exist_ok=True)
            self.file_path = self.directory / file_name
            self.loss_progress_file = self.directory / f'loss_progress.{file_name}'
        self.loss_history = []
        self.arm = model_builder.arm
        self.n_resources = 0

    @abstractmethod
    def evaluate(self, n_resources):
        """Return OptimisationGoals after spending `n_resources` on this arm."""

    def _train(self, epoch, max_batches, batch_size):
        raise TypeError('Cannot call _train on this evaluator')

    def _test(self, is_validation):
        raise TypeError('Cannot call _test on this evaluator')

    def _save_checkpoint(self, epoch, val_error, test_error):
        raise TypeError('Cannot call _save_checkpoint on this evaluator')

    def __str__(self):
        return f"Evaluator of arm:\n{self.arm}"

    def __eq__(self, other):
        return isinstance(other, Evaluator) and self.arm == other.arm

    def __hash__(self):
        return hash(str(self))
