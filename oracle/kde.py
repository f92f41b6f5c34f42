"""Oracle: EPDF-OFE density (reference experiments/simulation_evaluation/plot_results.py:11-17).

TEST INFRASTRUCTURE - see oracle/__init__.py.  The reference calls sklearn's KernelDensity(kernel='epanechnikov') and
exponentiates score_samples on linspace(start, end, 1000); in one dimension that is the closed form below
(checked against sklearn itself in tests/test_oracle_kde.py).
"""
import numpy as np


def epdf(x, start, end, bandwidth, grid_points=1000, kernel='epanechnikov', block=256):
    x = np.asarray(x, np.float64)
    grid = np.linspace(start, end, grid_points)
    out = np.empty(grid_points)
    for i in range(0, grid_points, block):                      # blocked to bound memory at 1e6 samples
        u = (grid[i:i + block, None] - x[None, :]) / bandwidth
        if kernel == 'epanechnikov':
            out[i:i + block] = 0.75 * np.maximum(0.0, 1.0 - u * u).sum(axis=1)
        elif kernel == 'gaussian':
            out[i:i + block] = np.exp(-0.5 * u * u).sum(axis=1) / np.sqrt(2 * np.pi)
        else:
            raise ValueError(kernel)
    return out / (len(x) * bandwidth)


def summary(optimums):
    """Stats the reference prints and pickles (experiments/run_simulation.py:161-182)."""
    n = len(optimums)
    srt = sorted(optimums)
    return {'avg_optimum': np.mean(optimums), 'sample_std_dev': np.std(optimums, ddof=1),
            'avg_top_25%': np.mean(srt[:int(n / 4)]), 'top_quartile': srt[:int(n / 4)]}
