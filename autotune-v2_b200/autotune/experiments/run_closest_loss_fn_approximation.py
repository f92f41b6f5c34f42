"""Monte-Carlo evaluation of optimisers on the closest-known-loss-function approximation - CLI compatible with the
reference's experiments/run_closest_loss_fn_approximation.py (same flags, method names, result pickle with
`all_norm_optimums`), plus -n / --seed / --known-functions-dir / --synthetic-arms, which are module constants or absent
there.  The N repetitions run as one device pipeline (nearest-recorded-arm lookup fused with successive halving) instead
of the reference's python loop (:183-196).

    python -m autotune.experiments.run_closest_loss_fn_approximation -p mnist -m "sim(hb)" -iter 81 -n 10000

The reference's recorded-arm databases (`../../loss_functions/...pkl`, :93-129) are not distributed with it
(.gitignore:67); when they are absent a synthetic database of the same shape is used and the fact is printed.
"""
import argparse
import pickle
import time
from os.path import exists, join as join_path

import numpy as np

from autotune.benchmarks import (
    AVAILABLE_OPT_FUNCTIONS, CifarProblem, KnownFnProblem, MnistProblem, MrbiProblem, OptFunctionProblem, SvhnProblem)
from autotune.benchmarks.mnist_like import synthetic_database
from autotune.experiments.monte_carlo import run_repetitions
from autotune.optimisers import HyperbandOptimiser

KNOWN_FUNCTIONS_DIR = "../../loss_functions/"
OUTPUT_DIR = "."
N_RESOURCES, MAX_TIME, MAX_ITER, ETA = 18, None, 6, 3
PROBLEM, METHOD, MIN_OR_MAX = "mnist", "sim(hb)", "min"      # the reference's default method is a TPE hybrid, see below
N_SIMULATIONS = 7000
SYNTHETIC_ARMS = {'mnist': 425, 'cifar': 300}                 # sizes of the reference's databases


def optimisation_func(opt_goals):
    """fval."""
    return opt_goals.fval


def _get_args(argv=None):
    parser = argparse.ArgumentParser(description='Running optimisations')
    parser.add_argument('-i', '--input-dir', default=None, type=str, help='input dir (unused: nothing is trained)')
    parser.add_argument('-o', '--output-dir', default=OUTPUT_DIR, type=str, help='output dir')
    parser.add_argument('-time', '--max-time', default=MAX_TIME, type=int, help='max time (stop if exceeded)')
    parser.add_argument('-iter', '--max-iter', default=MAX_ITER, type=int, help='max iterations (stop if exceeded')
    parser.add_argument('-p', '--problem', default=PROBLEM, type=str, help='problem (eg. cifar, mnist, branin)')
    parser.add_argument('-m', '--method', default=METHOD, type=str, help='method (eg. sim(hb))')
    parser.add_argument('-opt', '--min-or-max', default=MIN_OR_MAX, type=str, help="min or max")
    parser.add_argument('-res', '--n-resources', default=N_RESOURCES, type=int, help='n_resources', required=False)
    parser.add_argument('-eta', default=ETA, type=int, help='halving rate for Hyperband', required=False)
    # constants of the reference script, exposed here
    parser.add_argument('-n', '--n-simulations', default=N_SIMULATIONS, type=int, help='number of repetitions')
    parser.add_argument('--seed', default=0, type=int)
    parser.add_argument('--known-functions-dir', default=KNOWN_FUNCTIONS_DIR, type=str)
    parser.add_argument('--synthetic-arms', default=None, type=int, help='size of the synthetic database')
    parser.add_argument('--curve-length', default=300, type=int, help='length of the synthetic loss curves')
    return parser.parse_args(argv)


def get_real_problem(arguments):
    """The problem whose domain normalises the arms (:75-90); real training problems are replaced by their domains."""
    name = arguments.problem.lower()
    if name == "cifar":
        return CifarProblem(arguments.input_dir, arguments.output_dir, dataset_loader=None)
    if name == "mnist":
        return MnistProblem(arguments.input_dir, arguments.output_dir)
    if name == "svhn":
        return SvhnProblem(arguments.input_dir, arguments.output_dir)
    if name == "mrbi":
        return MrbiProblem(arguments.input_dir, arguments.output_dir)
    if name in AVAILABLE_OPT_FUNCTIONS:
        return OptFunctionProblem(name)
    raise ValueError(f"Supplied problem {name} does not exist")


def _load_recorded(arguments, directory):
    """The reference's loaders (:95-123), for users who have its pickles.  Returns None when a file is missing."""
    name = arguments.problem.lower()
    if name == "cifar":
        files = [join_path(directory, "cifar", f"best_loss_functions-{n}.pkl") for n in range(1, 2)]
    else:
        files = [join_path(directory, "mnist", f"results-mnist-random-{n}.pkl") for n in (10, 20, 15, 19, 1, 100, 30)] + \
                [join_path(directory, "mnist", f"results-mnist-tpe-{n}.pkl") for n in (30, 200)]
    if not all(exists(f) for f in files):
        return None
    known = {}
    for file in files:
        with open(file, "rb") as f_:
            data = pickle.load(f_)
        if name == "cifar":
            for _, evals in data.items():
                for loss_func, arm in evals:
                    if len(loss_func[1:]) > 300:
                        known[arm] = loss_func[1:]
        else:
            _, eval_history, _ = data
            for evaluator_t, _ in eval_history:
                known[evaluator_t.arm] = evaluator_t.loss_history[1:]
    return known


def get_known_functions(arguments, known_functions_dir=KNOWN_FUNCTIONS_DIR):
    """{Arm: [loss per epoch]} (:93-129)."""
    name = arguments.problem.lower()
    if name in ("cifar", "mnist"):
        known = _load_recorded(arguments, known_functions_dir)
        if known is None:
            n_arms = arguments.synthetic_arms or SYNTHETIC_ARMS[name]
            print(f"> no recorded arms under {known_functions_dir!r}: using a synthetic {name}-shaped database "
                  f"({n_arms} arms, {arguments.curve_length}-epoch curves)")
            known = synthetic_database(get_real_problem(arguments), n_arms, arguments.curve_length, seed=arguments.seed)
        return known
    if name in ("svhn", "mrbi"):
        return {}
    if name in AVAILABLE_OPT_FUNCTIONS:              # 1000 random arms, constant curves of length max_iter (:118-123)
        fn_problem = OptFunctionProblem(name)
        known_fs = {}
        for _ in range(1000):
            evaluator = fn_problem.get_evaluator()
            known_fs[evaluator.arm] = [evaluator.evaluate(1).fval for _ in range(arguments.max_iter)]
        return known_fs
    raise ValueError(f"Supplied problem {name} does not exist")


def get_sequential_optimiser(args):
    method = args.method.lower()
    min_or_max = min if args.min_or_max == 'min' else max
    if method == "sim(hb)":
        return HyperbandOptimiser(eta=args.eta, max_iter=args.max_iter, max_time=args.max_time, min_or_max=min_or_max,
                                  optimisation_func=optimisation_func, is_simulation=True)
    if method in ("sim(random)", "sim(tpe)", "sim(hb+tpe)") or "sim(hb+tpe+transfer" in method:
        raise NotImplementedError(f'{method} on recorded arms: the device TPE sampler covers the 2-D uniform domains of the '
                                  'simulation problems; mixed log/quantised domains would need hyperopt (not a dependency)')
    if method in ("sim(sigopt)", "sim(hb+sigopt)"):
        raise NotImplementedError('SigOpt is a remote proprietary service; not part of this package')
    raise ValueError(f"Supplied problem {method} does not exist in SEQUENTIAL mode")


def summarise(method, optimums):
    """The dictionary the reference pickles (:207-217): `all_norm_optimums`, no `top_quartile`."""
    vals = np.asarray(optimums.detach().cpu() if hasattr(optimums, 'detach') else optimums, dtype=np.float64)
    n = len(vals)
    top = sorted(vals.tolist())[:int(n / 4)]
    return {"method": method, "n_simulations": n, "all_norm_optimums": vals.tolist(), "avg_optimum": np.mean(vals),
            "sample_std_dev": np.std(vals, ddof=1), "avg_top_25%": np.mean(top) if top else float('nan')}


def main(argv=None):
    args = _get_args(argv)
    print(f"\n    Problem:          {args.problem.upper()}\n    Method:           {args.method.upper()}\n"
          f"    Func {args.min_or_max.upper()}imizes:   {optimisation_func.__doc__}\n")
    known_fns = get_known_functions(args, args.known_functions_dir)
    real_problem = get_real_problem(args)
    real_problem.log_domain()
    problem = KnownFnProblem(known_fs=known_fns, real_problem=real_problem)
    optimiser = get_sequential_optimiser(args)
    t0 = time.time()
    optimums = run_repetitions(optimiser, problem, args.n_simulations, seed=args.seed)
    stats = summarise(args.method, optimums)
    print(f"""\n\n------------- OPTIMUM STATISTICS OVER {args.n_simulations} SIMULATIONS -------------
    average optimum: {stats['avg_optimum']}
    sample std dev:  {stats['sample_std_dev']}
    avg top 25%:     {stats['avg_top_25%']}
    distinct optima: {len(set(stats['all_norm_optimums']))}
    wall time:       {time.time() - t0:.3f} s
    """)
    path = join_path(args.output_dir, f"known-fns-hist-{args.method}-{args.n_simulations}.pkl")
    with open(path, "wb") as f:
        pickle.dump(stats, f)
    print('results written to', path)
    return stats


if __name__ == "__main__":
    main()
