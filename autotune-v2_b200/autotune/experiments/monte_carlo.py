"""Batched replacement of the reference's repetition loops (experiments/run_simulation.py:144-182 and
experiments/run_closest_loss_fn_approximation.py:175-217): N independent optimiser runs as one kernel pipeline, sharded
over the GPUs of the box when torch.distributed is initialised, with the reference's summary statistics and result
pickle schema."""
import pickle

import numpy as np

from autotune.optimisers import (
    HybridHyperbandTpeOptimiser, HyperbandOptimiser, RandomOptimiser, TpeOptimiser)
from autotune.optimisers import _device


def _runner_for(optimiser, problem):
    """(spec, runner(spec, n, seed, rep_offset, device) -> ofe tensor) for an optimiser/problem pair."""
    import torch
    from autotune.b200 import engine
    from autotune.benchmarks.known_loss_fn_problem import KnownFnProblem
    _device.require_fval_objective(optimiser.optimisation_func)
    if isinstance(problem, KnownFnProblem):
        if not isinstance(optimiser, HyperbandOptimiser) or optimiser._tpe_transfer is not None:
            raise NotImplementedError('closest-known-function Monte Carlo is implemented for HyperbandOptimiser')
        db = problem.device_db()

        def runner(_spec, n, seed, off, dev):
            # float64: the OFEs are recorded losses, returned exactly as the reference returns them
            return engine.run_knownfn_hyperband(db, optimiser.max_iter, optimiser.eta, n, seed=seed, rep_offset=off,
                                                min_or_max=optimiser.min_or_max)['ofe']
        runner.dtype = torch.float64
        return None, runner
    spec = _device.simulation_spec(optimiser, problem)
    if isinstance(optimiser, HyperbandOptimiser):
        tpe = optimiser._tpe_config()
        if not tpe.enabled:
            def runner(spec_, n, seed, off, dev):      # plain Hyperband: the leanest kernel
                return engine.run_hyperband_repetitions(spec_, n, seed=seed, rep_offset=off, device=dev)['ofe']
        else:
            def runner(spec_, n, seed, off, dev):
                return engine.run_optimiser_repetitions(spec_, n, tpe, seed=seed, rep_offset=off, device=dev)['ofe']
        return spec, runner
    if isinstance(optimiser, TpeOptimiser):            # TpeOptimiser and RandomOptimiser
        plan = engine.single_round_plan(spec.R, int(optimiser.max_iter), int(optimiser.n_resources))
        tpe = engine.tpe_config(optimiser._tpe_enabled)

        def runner(spec_, n, seed, off, dev):
            return engine.run_optimiser_repetitions(spec_, n, tpe, seed=seed, rep_offset=off, device=dev, plan=plan)['ofe']
        return spec, runner
    raise NotImplementedError(f'{type(optimiser).__name__} has no batched device path')


def run_repetitions(optimiser, problem, n_repetitions, seed=0, device=None, group=None):
    """OFEs of `n_repetitions` independent runs of `optimiser` on `problem` (all ranks' results, in repetition order,
    as a CUDA tensor: float32 for the simulation problems, the float64 recorded losses for the closest-known-function
    problem).  Equivalent to looping `optimiser.run_optimisation(problem)` with fresh objects as the
    reference drivers do - the same kernels, one launch per rank."""
    from autotune.b200.engine import _cuda_device
    from autotune.b200.epdf import gather_slices, shard_range
    import torch
    import torch.distributed as dist
    device = _cuda_device(device)
    spec, runner = _runner_for(optimiser, problem)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(n_repetitions, rank, world)
    local = runner(spec, hi - lo, seed, lo, device) if hi > lo else \
        torch.empty(0, device=device, dtype=getattr(runner, 'dtype', torch.float32))
    return gather_slices(local, n_repetitions, group)


def summarise(method, optimums):
    """The dictionary the reference pickles (run_simulation.py:171-182), from a device tensor or a sequence."""
    vals = np.asarray(optimums.detach().cpu() if hasattr(optimums, 'detach') else optimums, dtype=np.float64)
    n = len(vals)
    top = sorted(vals.tolist())[:int(n / 4)]
    return {'method': method, 'n_simulations': n, 'all_optimums': vals.tolist(), 'avg_optimum': np.mean(vals),
            'sample_std_dev': np.std(vals, ddof=1), 'avg_top_25%': np.mean(top) if top else float('nan'),
            'top_quartile': top}


def save_results(path, method, optimums):
    with open(path, 'wb') as f:
        pickle.dump(summarise(method, optimums), f)
