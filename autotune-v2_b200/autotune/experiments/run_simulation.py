"""Monte-Carlo evaluation of optimisers on simulated loss functions - CLI compatible with the reference's
experiments/run_simulation.py (same flags, method names, result pickle), plus -n / --seed / --output-dir, which are
module constants there.  The N repetitions run as one device pipeline instead of a python loop.

    python -m autotune.experiments.run_simulation -p sim-rastrigin -m "sim(hb)" -n 100000
"""
import argparse
import time
from os.path import join as join_path

import numpy as np

from autotune.benchmarks import OptFunctionSimulationProblem
from autotune.core import RoundRobinShapeFamilyScheduler, ShapeFamily, UniformShapeFamilyScheduler
from autotune.experiments.monte_carlo import run_repetitions, save_results, summarise
from autotune.optimisers import (
    HybridHyperbandTpeNoTransferOptimiser, HybridHyperbandTpeOptimiser, HybridHyperbandTpeTransferAllOptimiser,
    HybridHyperbandTpeTransferLongestOptimiser, HybridHyperbandTpeTransferSameOptimiser, HyperbandOptimiser,
    RandomOptimiser, TpeOptimiser)

OUTPUT_DIR = "."
N_RESOURCES, MAX_TIME, MAX_ITER, ETA = 24, None, 81, 3
PROBLEM, METHOD, MIN_OR_MAX = "sim-rastrigin", "sim(hb+tpe+transfer+same)", "min"
N_SIMULATIONS, INIT_NOISE = 7000, 10
SCHEDULING = "uniform"

families_of_shapes_egg = (
    ShapeFamily(None, 1.5, 10, 15, False, 0, 1000),   # aggressive start
    ShapeFamily(None, 0.5, 7, 10, False, 0, 1000),    # average aggressiveness
    ShapeFamily(None, 0.2, 4, 7, True, 0, 1000),      # slow start, aggressive end (smoothed)
)
families_of_shapes_general = (
    ShapeFamily(None, 1.5, 10, 15, False), ShapeFamily(None, 0.5, 7, 10, False), ShapeFamily(None, 0.2, 4, 7, True),
    ShapeFamily(None, 1.5, 10, 15, False, 200, 400), ShapeFamily(None, 0.5, 7, 10, False, 200, 400),
    ShapeFamily(None, 0.2, 4, 7, True, 200, 400),
)
families_flat = (ShapeFamily(None, 0, 1, 0, False, 0, 0),)
FAMILY_SETS = {'general': families_of_shapes_general, 'general3': families_of_shapes_general[:3],
               'egg': families_of_shapes_egg, 'flat': families_flat}
PROBLEM_FUNCS = {'sim-branin': 'branin', 'sim-egg': 'egg', 'sim-camel': 'camel', 'sim-wave': 'wave',
                 'sim-rastrigin': 'rastrigin'}


def optimisation_func(opt_goals):
    """fval."""
    return opt_goals.fval


def _get_args(argv=None):
    parser = argparse.ArgumentParser(description='Running optimisations')
    parser.add_argument('-time', '--max-time', default=MAX_TIME, type=int, help='max time (stop if exceeded)')
    parser.add_argument('-iter', '--max-iter', default=MAX_ITER, type=int, help='max iterations (stop if exceeded')
    parser.add_argument('-p', '--problem', default=PROBLEM, type=str, help='problem (eg. sim-rastrigin, sim-branin)')
    parser.add_argument('-m', '--method', default=METHOD, type=str, help='method (eg. sim(hb), sim(tpe))')
    parser.add_argument('-opt', '--min-or-max', default=MIN_OR_MAX, type=str, help="min or max")
    parser.add_argument('-res', '--n-resources', default=N_RESOURCES, type=int, help='n_resources', required=False)
    parser.add_argument('-eta', default=ETA, type=int, help='halving rate for Hyperband', required=False)
    # constants of the reference script, exposed here
    parser.add_argument('-n', '--n-simulations', default=N_SIMULATIONS, type=int, help='number of repetitions')
    parser.add_argument('--init-noise', default=INIT_NOISE, type=float)
    parser.add_argument('--families', default='general', choices=sorted(FAMILY_SETS))
    parser.add_argument('--scheduling', default=SCHEDULING, choices=['uniform', 'round-robin'])
    parser.add_argument('--seed', default=0, type=int)
    parser.add_argument('--output-dir', default=OUTPUT_DIR, type=str)
    return parser.parse_args(argv)


def get_problem(arguments):
    return OptFunctionSimulationProblem(func_name=PROBLEM_FUNCS[arguments.problem.lower()])


def get_optimiser(args):
    method = args.method.lower()
    min_or_max = min if args.min_or_max == 'min' else max
    scheduler_cls = {'uniform': UniformShapeFamilyScheduler, 'round-robin': RoundRobinShapeFamilyScheduler}[args.scheduling]
    scheduler = scheduler_cls(shape_families=FAMILY_SETS[args.families], max_resources=args.max_iter,
                              init_noise=args.init_noise)
    common = dict(max_iter=args.max_iter, max_time=args.max_time, min_or_max=min_or_max,
                  optimisation_func=optimisation_func, is_simulation=True, scheduler=scheduler)
    if method == "sim(random)":
        return RandomOptimiser(n_resources=args.n_resources, **common)
    if method == "sim(hb)":
        return HyperbandOptimiser(eta=args.eta, **common)
    if method == "sim(tpe)":
        return TpeOptimiser(n_resources=args.n_resources, **common)
    if method == "sim(hb+tpe)":
        return HybridHyperbandTpeOptimiser(eta=args.eta, **common)
    hybrids = {"sim(hb+tpe+transfer+none)": HybridHyperbandTpeNoTransferOptimiser,
               "sim(hb+tpe+transfer+all)": HybridHyperbandTpeTransferAllOptimiser,
               "sim(hb+tpe+transfer+longest)": HybridHyperbandTpeTransferLongestOptimiser,
               "sim(hb+tpe+transfer+same)": HybridHyperbandTpeTransferSameOptimiser}
    if method in hybrids:
        return hybrids[method](eta=args.eta, **common)
    if method in ("sim(sigopt)", "sim(hb+sigopt)"):
        raise NotImplementedError('SigOpt is a remote proprietary service; not part of this package')
    raise ValueError(f"Supplied problem {method} does not exist")


def main(argv=None):
    args = _get_args(argv)
    print(f"\n    Problem:          {args.problem.upper()}\n    Method:           {args.method.upper()}\n"
          f"    Func {args.min_or_max.upper()}imizes:   {optimisation_func.__doc__}\n")
    problem, optimiser = get_problem(args), get_optimiser(args)
    t0 = time.time()
    optimums = run_repetitions(optimiser, problem, args.n_simulations, seed=args.seed)
    stats = summarise(args.method, optimums)
    print(f"""\n\n------------- OPTIMUM STATISTICS OVER {args.n_simulations} SIMULATIONS -------------
    average optimum: {stats['avg_optimum']}
    sample std dev:  {stats['sample_std_dev']}
    avg top 25%:     {stats['avg_top_25%']}
    wall time:       {time.time() - t0:.3f} s
    """)
    path = join_path(args.output_dir, f"hist-{args.problem}-{args.method}-{args.n_simulations}.pkl")
    save_results(path, args.method, optimums)
    print('results written to', path)
    return stats


if __name__ == "__main__":
    main()
