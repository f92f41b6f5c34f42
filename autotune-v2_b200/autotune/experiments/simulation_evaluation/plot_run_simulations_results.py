"""Compare the stored OFE vectors of several methods on one simulated problem: two-sample KS tests, histograms and the
EPDF-OFE (reference experiments/simulation_evaluation/plot_run_simulations_results.py), with the statistics computed on
the device.  The per-problem plot ranges, histogram steps and KDE bandwidths are the reference's (:27-54).

    python -m autotune.experiments.simulation_evaluation.plot_run_simulations_results \\
        --simulation rastrigin-families-2 --dir epdf-ofes/simulation-rastrigin-families-2 -n 7000
"""
import argparse
import pickle
from os.path import exists, join as join_path

from autotune.experiments.simulation_evaluation.plot_results import plot_epdf_ofe, plot_histograms
from autotune.experiments.simulation_evaluation.stats import ks_2samp

SIMULATION = "rastrigin-families-2"
START = {"flat-branin": 0.37, "flat-drop-wave": -1, "flat-rastrigin": 0, "rastrigin-families-1": -200,
         "rastrigin-families-2": -400}
END = {"flat-branin": 1.5, "flat-drop-wave": -0.3, "flat-rastrigin": 18, "rastrigin-families-1": -170,
       "rastrigin-families-2": -370}
STEP = {"flat-branin": 0.05, "flat-drop-wave": 0.02, "flat-rastrigin": 1, "rastrigin-families-1": 1,
        "rastrigin-families-2": 1}
BANDWIDTH = {"flat-branin": 0.1, "flat-drop-wave": 0.02, "flat-rastrigin": 3, "rastrigin-families-1": 3,
             "rastrigin-families-2": 3}
METHODS = [("Hyperband", "sim(hb)"), ("NONE", "sim(hb+tpe+transfer+none)"), ("ALL", "sim(hb+tpe+transfer+all)"),
           ("SURV", "sim(hb+tpe+transfer+longest)"), ("SAME", "sim(hb+tpe+transfer+same)"), ("TPE", "sim(tpe)"),
           ("TPE 2xBUDGET", "sim(tpe2xbudget)")]
PAIRS = [("TPE 2xBUDGET", "Hyperband"), ("Hyperband", "ALL"), ("ALL", "SAME"), ("SAME", "NONE"), ("NONE", "SURV")]   # :68-73
FONT_SIZE = 50


def unpickle(output_dir, problem, method, n_simulations, til=1):
    """all_optimums of hist-<problem>-<method>-<n>[-<til>].pkl: the reference's stored files carry the `-til` suffix
    (:14-21), the drivers of this package write none."""
    for name in (f"hist-{problem}-{method}-{n_simulations}-{til}.pkl", f"hist-{problem}-{method}-{n_simulations}.pkl",
                 f"hist-{method}-{n_simulations}-{til}.pkl"):
        path = join_path(output_dir, name)
        if exists(path):
            with open(path, "rb") as f:
                return list(pickle.load(f)["all_optimums"])
    return None


def compare(datasets, simulation, show=True):
    """datasets: {label: OFE sequence or tensor}.  Returns {'ks': {(a, b): (statistic, pvalue)}, 'histograms', 'epdf'}."""
    start, end, step, bandwidth = START[simulation], END[simulation], STEP[simulation], BANDWIDTH[simulation]
    bins = [start + step * i for i in range(1 + int((end - start) / step))]
    labels = list(datasets)
    ks = {}
    for a, b in PAIRS:
        if a in datasets and b in datasets:
            res = ks_2samp(datasets[a], datasets[b])
            ks[(a, b)] = (float(res.statistic), float(res.pvalue))
            if show:
                print(f"{a} vs {b}", res)
    data = tuple(datasets[k] for k in labels)
    return {'ks': ks, 'bins': bins, 'labels': labels,
            'histograms': plot_histograms(data, labels, bins, font_size=FONT_SIZE),
            'epdf': plot_epdf_ofe(data, labels, start=start, end=end, bandwidth=bandwidth, order_profile=False,
                                  font_size=FONT_SIZE)}


def main(argv=None):
    parser = argparse.ArgumentParser(description='KS tests, histograms and EPDF-OFE of stored optimiser results')
    parser.add_argument('--simulation', default=SIMULATION, choices=sorted(START))
    parser.add_argument('--dir', default=None, help='directory of the result pickles')
    parser.add_argument('--problem', default=None, help='problem tag in the file names (default: from --simulation)')
    parser.add_argument('-n', '--n-simulations', default=7000, type=int)
    args = parser.parse_args(argv)
    output_dir = args.dir or f"../../../epdf-ofes/simulation-{args.simulation}/"
    problem = args.problem or {"flat-branin": "sim-branin", "flat-drop-wave": "sim-wave"}.get(args.simulation, "sim-rastrigin")
    datasets = {}
    for label, method in METHODS:
        vals = unpickle(output_dir, problem, method, args.n_simulations)
        if vals is not None:
            datasets[label] = vals
    if not datasets:
        raise FileNotFoundError(f'no result pickles for {problem} with {args.n_simulations} simulations under {output_dir}')
    print({k: len(v) for k, v in datasets.items()})
    return compare(datasets, args.simulation)


if __name__ == "__main__":
    main()
