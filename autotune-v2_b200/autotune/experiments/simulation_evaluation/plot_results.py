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
