"""ModelBuilder: holds the arm a model would be built from (reference core/model_builder.py:10-28)."""


class ModelBuilder:
    __slots__ = ('arm', 'ml_model')

    def __init__(self, arm, ml_model=None):
        self.arm, self.ml_model = arm, ml_model

    def construct_model(self):
        """Simulation problems have no model to construct."""
        return None
