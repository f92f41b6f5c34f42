"""Oracle KDE vs the reference's actual call (sklearn KernelDensity, plot_results.py:11-17) on the stored OFE vectors."""
import numpy as np
import pytest
from sklearn.neighbors import KernelDensity

from oracle.kde import epdf
from tests.helpers import GOLDEN

CFG = {  # plot_run_simulations_results.py:27-54
    'simulation-flat-branin/hist-sim(hb)-7000-1': (0.37, 1.5, 0.1),
    'simulation-flat-drop-wave/hist-sim-wave-sim(hb)-7000-1': (-1, -0.3, 0.02),
    'simulation-rastrigin-families-2/hist-sim-rastrigin-sim(hb)-7000-1': (-400, -370, 3),
}


@pytest.mark.parametrize('key', sorted(CFG))
def test_closed_form_matches_sklearn(key):
    x = np.load(f'{GOLDEN}/epdf_ofes.npz')[key]
    start, end, h = CFG[key]
    kde = KernelDensity(kernel='epanechnikov', bandwidth=h).fit(x[:, None])
    want = np.exp(kde.score_samples(np.linspace(start, end, 1000)[:, None]))
    got = epdf(x, start, end, h)
    assert np.max(np.abs(got - want)) < 1e-12 * max(1.0, want.max())
    assert abs(np.trapezoid(epdf(x, x.min() - 2 * h, x.max() + 2 * h, h, 4000),
                            np.linspace(x.min() - 2 * h, x.max() + 2 * h, 4000)) - 1) < 1e-3


def test_gaussian_matches_sklearn():
    x = np.load(f'{GOLDEN}/epdf_ofes.npz')['simulation-flat-branin/hist-sim(hb)-7000-1']
    kde = KernelDensity(kernel='gaussian', bandwidth=0.1).fit(x[:, None])
    want = np.exp(kde.score_samples(np.linspace(0.37, 1.5, 1000)[:, None]))
    assert np.max(np.abs(epdf(x, 0.37, 1.5, 0.1, kernel='gaussian') - want)) < 1e-10
