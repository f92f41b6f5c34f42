"""Profiles of a family of loss curves (reference experiments/simulation_evaluation/profiles.py): per-epoch average,
median and standard deviation, the "dynamic order" profile (how much the ranking of the curves is preserved from one
epoch to the next) and the "ends order" profile (ranking at the first vs the last epoch).  Statistics run on the
device; curves come from the full-trajectory kernel (`plot_simulated`)."""
import numpy as np
import torch

from autotune.benchmarks.opt_function_problem import DOMAINS
from autotune.core import RoundRobinShapeFamilyScheduler, ShapeFamily


def _dev(curves):
    t = curves if isinstance(curves, torch.Tensor) else torch.as_tensor(np.asarray(curves, dtype=np.float64))
    return t.to('cuda')


def get_avg_profile(loss_functions_per_fam):
    return _dev(loss_functions_per_fam).double().mean(dim=0).cpu().numpy()


def get_med_profile(loss_functions_per_fam):
    return torch.quantile(_dev(loss_functions_per_fam).double(), 0.5, dim=0, interpolation='linear').cpu().numpy()


def get_std_profile(loss_functions_per_fam):
    return _dev(loss_functions_per_fam).double().std(dim=0, unbiased=False).cpu().numpy()


def get_dynamic_order_profile(loss_functions_per_fam):
    """[T-1] rank-agreement between consecutive epochs (O(F^2 T) pairwise kernel, torch.ops.at2.order_profile)."""
    dyn, _ = torch.ops.at2.order_profile(_dev(loss_functions_per_fam).float().contiguous())
    return dyn.double().cpu().tolist()


def get_ends_order_profile(loss_functions_per_fam):
    _, ends = torch.ops.at2.order_profile(_dev(loss_functions_per_fam).float().contiguous())
    return float(ends[0])


def plot_simulated(func_name, n_simulations, max_resources=81, n_resources=None,
                   shape_families=(ShapeFamily(None, 0.9, 10, 0.1),), init_noise=0,
                   scheduler=RoundRobinShapeFamilyScheduler, seed=0):
    """`n_simulations` simulated loss curves (lists of max_resources values, smoothed where the family says so) with
    the families dealt by `scheduler` - one batched launch of the trajectory kernel instead of one evaluator per curve."""
    from autotune.b200 import trajectories
    from autotune.b200.engine import HyperbandSimSpec
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    if n_resources is not None:
        assert n_resources <= max_resources
    kind = getattr(scheduler, 'device_kind', None)
    if kind is None:
        raise NotImplementedError('only the uniform and round-robin schedulers are supported')
    fams = [tuple(f[1:]) for f in shape_families]
    spec = HyperbandSimSpec(func_name, fams, max_resources, 3, init_noise, kind)
    rng = np.random.default_rng(seed)
    dom = DOMAINS[func_name]
    x = rng.uniform(dom.x.min_val, dom.x.max_val, n_simulations)
    y = rng.uniform(dom.y.min_val, dom.y.max_val, n_simulations)
    fam = np.arange(n_simulations) % len(fams) if kind == 'round-robin' else rng.integers(0, len(fams), n_simulations)
    f = NOISE_FREE_FUNCTIONS[func_name](x, y)
    shifts = np.asarray([[fm[4], fm[5]] for fm in fams], dtype=np.float64)
    f1 = f + init_noise * rng.standard_normal(n_simulations) - shifts[fam, 0]
    fn = f - shifts[fam, 1]
    raw = trajectories.simulate_arms(spec, f1, fn, fam, np.arange(n_simulations) + (int(seed) << 32), seed=seed)
    out = raw.double()
    smooth_mask = torch.as_tensor(np.asarray([bool(fams[k][3]) for k in fam]), device=out.device)
    if bool(smooth_mask.any()):
        out = torch.where(smooth_mask[:, None], trajectories.smooth(raw), out)
    return out.cpu().tolist()
