"""EPDF-OFE: estimated PDF of the optimal final errors (reference experiments/simulation_evaluation/plot_results.py).

`plot_smooth` evaluates the Epanechnikov kernel density on a 1000-point grid with the device KDE kernel
(torch.ops.at2.kde) - the reference calls sklearn's KernelDensity for this - and draws it when matplotlib is available.
"""
import numpy as np


def epdf(x, start, end, bandwidth=0.5, grid_points=1000, kernel='epanechnikov'):
    """Density of the sample `x` (sequence or tensor) on linspace(start, end, grid_points) as a numpy array."""
    import torch

    from autotune.b200.epdf import density
    t = x if hasattr(x, 'is_cuda') else torch.as_tensor(np.asarray(x, dtype=np.float32))
    t = t.to(device='cuda', dtype=torch.float32).contiguous()
    return density(t, start, end, bandwidth, grid_points, kernel).cpu().numpy().astype(np.float64)


def plot_smooth(x, start, end, bandwidth=0.5):
    dens = epdf(x, start, end, bandwidth)
    try:
        import matplotlib.pyplot as plt
        plt.plot(np.linspace(start, end, 1000), dens, '-', label="epanechnikov", linewidth=3)
    except ImportError:
        pass
    return dens


def plot_epdf_ofe(data_, labels_, *, start, end, bandwidth, order_profile=False, font_size=10):
    """One EPDF per dataset; returns the list of densities (and shows the figure when matplotlib is available)."""
    smoothed = [plot_smooth(dataset, start, end, bandwidth=bandwidth) for dataset in data_]
    try:
        import matplotlib.pyplot as plt
        plt.xlabel("Minimum final error found by one optimizer run (OFE)", fontsize=font_size)
        plt.ylabel("Estimated PDF", fontsize=font_size)
        plt.legend(labels_, prop={'size': font_size})
        plt.show()
    except ImportError:
        pass
    return smoothed


def histograms(data_, bins, density=False):
    """Counts (or matplotlib's `density=True` normalisation) of every dataset over the bin edges `bins`, on the device:
    the numbers behind `plot_histograms` (reference plot_results.py:53-61).  The last bin is closed on the right, as in
    numpy / matplotlib."""
    import torch
    edges = torch.as_tensor(np.asarray(bins, dtype=np.float64), device='cuda')
    out = []
    for dataset in data_:
        t = dataset if hasattr(dataset, 'is_cuda') else torch.as_tensor(np.asarray(dataset, dtype=np.float64))
        t = t.to(device='cuda', dtype=torch.float64).contiguous()
        idx = torch.bucketize(t, edges, right=True) - 1                   # bin i holds edges[i] <= x < edges[i+1]
        idx = torch.where(t == edges[-1], torch.full_like(idx, edges.numel() - 2), idx)
        inside = (idx >= 0) & (idx < edges.numel() - 1)
        counts = torch.bincount(idx[inside], minlength=edges.numel() - 1).double()
        if density:
            counts = counts / (counts.sum() * (edges[1:] - edges[:-1]))
        out.append(counts.cpu().numpy())
    return out


def plot_histograms(data_, labels_, bins, *, font_size=10):
    """Histograms of the OFE samples (reference plot_results.py:53-69); returns the density-normalised counts and draws the
    reference's two figures when matplotlib is available."""
    dens = histograms(data_, bins, density=True)
    try:
        import matplotlib.pyplot as plt
        host = [d.detach().cpu().numpy() if hasattr(d, 'detach') else np.asarray(d) for d in data_]
        for kw in (dict(histtype='bar', density=True), dict(histtype='step', fill=True, linewidth=2)):
            plt.hist(host, bins=bins, label=labels_, cumulative=False, stacked=False, **kw)
            plt.legend(prop={'size': font_size})
            plt.xlabel("Minimum final error found by one optimizer run (OFE)", fontsize=font_size)
            plt.ylabel("Count occurrences", fontsize=font_size)
            plt.show()
    except ImportError:
        pass
    return dens
