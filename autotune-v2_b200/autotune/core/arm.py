"""Arm: one assignment of hyperparameter values (reference core/arm.py:9-92)."""
from types import SimpleNamespace


class Arm(SimpleNamespace):
    """`Arm(x=1.0, y=2.0)`; values are plain attributes, `arm['x']` also works."""

    def set_default_values(self, *, domain, hyperparams_to_opt):
        """Hyperparameters of the domain that are not optimised take their `init_val` (arm.py:29-42)."""
        for name in domain.hyperparams_names():
            if hasattr(self, name) or name in hyperparams_to_opt:
                continue
            default = domain[name].init_val
            if default is None:
                raise ValueError(f"No default value for param {name} was supplied")
            setattr(self, name, default)

    def draw_hp_val(self, *, domain, hyperparams_to_opt):
        """Random value (numpy global RNG) for every optimised hyperparameter, in domain order (arm.py:44-55)."""
        self.set_default_values(domain=domain, hyperparams_to_opt=hyperparams_to_opt)
        for name in domain.hyperparams_names():
            if name in hyperparams_to_opt:
                setattr(self, name, domain[name].get_param_range(1, stochastic=True)[0])

    @staticmethod
    def normalize(arm, *, domain):
        """(v - min_val) / (max_val - min_val) per shared hyperparameter (arm.py:57-70).  For log-scale params the
        bounds are exponents while v is linear - the reference's quirk, preserved."""
        shared = set(domain.hyperparams_names()) & set(arm.__dict__.keys())
        return Arm(**{n: (arm[n] - domain[n].min_val) / (domain[n].max_val - domain[n].min_val) for n in shared})

    def __str__(self):
        width = max(len(n) for n in self.__dict__)
        return '\n'.join(f"   - {n}:{' ' * (width - len(n) + 1)}{v}" for n, v in self.__dict__.items())

    def __hash__(self):
        return hash(str(self))

    def __getitem__(self, name):
        return getattr(self, name)
