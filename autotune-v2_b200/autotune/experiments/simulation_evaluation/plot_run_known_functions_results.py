"""Compare the stored OFE vectors of several methods on the closest-known-loss-function problem: KS tests, histograms and
three EPDF-OFE zoom levels (reference experiments/simulation_evaluation/plot_run_known_functions_results.py; bins and
bandwidths are the reference's, :27,65-70), with the statistics computed on the device.

    python -m autotune.experiments.simulation_evaluation.plot_run_known_functions_results --dir <pickles> -n 7000
"""
import argparse
import pickle
from os.path import exists, join as join_path

from autotune.experiments.simulation_evaluation.plot_results import plot_epdf_ofe, plot_histograms
from autotune.experiments.simulation_evaluation.stats import ks_2samp

OUTPUT_DIR = "../../../epdf-ofes/closest-known-loss-fn-mnist/"
METHODS = [("Hyperband", "sim(hb)"), ("NONE", "sim(hb+tpe+transfer+none)"), ("ALL", "sim(hb+tpe+transfer+all)"),
           ("SURV", "sim(hb+tpe+transfer+longest)"), ("SAME", "sim(hb+tpe+transfer+same)"), ("TPE", "sim(tpe)"),
           ("TPE 2xBUDGET", "sim(tpe2xbudget)")]
PAIRS = [("TPE", "TPE 2xBUDGET"), ("TPE 2xBUDGET", "Hyperband"), ("Hyperband", "ALL"), ("ALL", "SAME"), ("SAME", "NONE"),
         ("SURV", "NONE")]                                                             # :43-48
FONT_SIZE = 50


def unpickle(output_dir, method, n_simulations, til=1):
    """all_norm_optimums of known-fns-hist-<method>-<n>[-<til>].pkl (the reference's files carry the suffix, :13-19)."""
    for name in (f"known-fns-hist-{method}-{n_simulations}-{til}.pkl", f"known-fns-hist-{method}-{n_simulations}.pkl"):
        path = join_path(output_dir, name)
        if exists(path):
            with open(path, "rb") as f:
                return list(pickle.load(f)["all_norm_optimums"])
    return None


def compare(datasets, show=True):
    bins = [0.074 + i * 0.001 for i in range(20)]                                      # :27
    labels = list(datasets)
    ks = {}
    for a, b in PAIRS:
        if a in datasets and b in datasets:
            res = ks_2samp(datasets[a], datasets[b])
            ks[(a, b)] = (float(res.statistic), float(res.pvalue))
            if show:
                print(f"{a} vs {b}", res)
    data = tuple(datasets[k] for k in labels)
    zooms = [(bins[int(len(bins) / 4)], 0.003), (bins[int(len(bins) / 2)], 0.005), (bins[-1], 0.005)]   # :65-70
    return {'ks': ks, 'bins': bins, 'labels': labels,
            'histograms': plot_histograms(data, labels, bins, font_size=FONT_SIZE),
            'epdf': [plot_epdf_ofe(data, labels, start=bins[0], end=end, bandwidth=h, order_profile=False,
                                   font_size=FONT_SIZE) for end, h in zooms]}


def main(argv=None):
    parser = argparse.ArgumentParser(description='KS tests, histograms and EPDF-OFE of stored closest-known-function results')
    parser.add_argument('--dir', default=OUTPUT_DIR, help='directory of the result pickles')
    parser.add_argument('-n', '--n-simulations', default=7000, type=int)
    args = parser.parse_args(argv)
    datasets = {}
    for label, method in METHODS:
        vals = unpickle(args.dir, method, args.n_simulations)
        if vals is not None:
            datasets[label] = vals
    if not datasets:
        raise FileNotFoundError(f'no known-fns-hist-*-{args.n_simulations}*.pkl under {args.dir}')
    print({k: len(v) for k, v in datasets.items()})
    return compare(datasets)


if __name__ == "__main__":
    main()
