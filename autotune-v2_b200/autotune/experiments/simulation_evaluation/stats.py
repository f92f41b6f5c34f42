"""Two-sample Kolmogorov-Smirnov test between OFE vectors - what the reference prints for every pair of methods
(experiments/simulation_evaluation/plot_run_simulations_results.py:69-73 via scipy.stats.ks_2samp).  The statistic D is
computed on the device (torch.ops.at2.ks_statistic over the two sorted samples); the p-value is scipy's own
method='asymp' formula (kstwo.sf(D, round(n1 n2 / (n1 + n2)))), which is what scipy itself switches to above 10 000 samples."""
from collections import namedtuple

import numpy as np
import torch
from scipy.stats import distributions

KstestResult = namedtuple('KstestResult', ['statistic', 'pvalue'])


def _sorted_device(x):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    return torch.sort(t.to(device='cuda', dtype=torch.float32).contiguous()).values


def ks_2samp(data1, data2):
    a, b = _sorted_device(data1), _sorted_device(data2)
    d = float(torch.ops.at2.ks_statistic(a, b)[0])
    n1, n2 = a.numel(), b.numel()
    m, n = sorted([float(n1), float(n2)], reverse=True)
    en = m * n / (m + n)
    return KstestResult(d, float(np.clip(distributions.kstwo.sf(d, np.round(en)), 0, 1)))
