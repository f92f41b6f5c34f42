"""Kernel (e) against the oracle closed form (which is pinned to sklearn on CPU)."""
import numpy as np
import pytest
import torch

from oracle.kde import epdf
from tests.helpers import GOLDEN

import autotune.b200  # noqa: F401,E402  (registers torch.ops.at2.*)

pytestmark = pytest.mark.gpu

CFG = {
    'simulation-flat-branin/hist-sim(hb)-7000-1': (0.37, 1.5, 0.1),
    'simulation-flat-drop-wave/hist-sim-wave-sim(hb)-7000-1': (-1, -0.3, 0.02),
    'simulation-flat-rastrigin/hist-sim-rastrigin-sim(hb)-7000-1': (0, 18, 3),
    'simulation-rastrigin-families-1/hist-sim-rastrigin-sim(hb)-7000-1': (-200, -170, 3),
    'simulation-rastrigin-families-2/hist-sim-rastrigin-sim(hb)-7000-1': (-400, -370, 3),
}


def _dev_kde(x, start, end, h, G=1000, kernel=0):
    t = torch.as_tensor(np.asarray(x, np.float32)).cuda()
    return torch.ops.at2.kde(t, float(start), float(end), G, float(h), kernel).cpu().numpy().astype(np.float64)


@pytest.mark.parametrize('key', sorted(CFG))
def test_epanechnikov_on_reference_vectors(key):
    x = np.load(f'{GOLDEN}/epdf_ofes.npz')[key]
    start, end, h = CFG[key]
    want = epdf(x.astype(np.float32).astype(np.float64), start, end, h)
    got = _dev_kde(x, start, end, h)
    # float32 pair terms, double reduction: 1e-5 of the density scale
    assert np.max(np.abs(got - want)) < 1e-5 * want.max()


@pytest.mark.parametrize('n', [1, 3, 511, 512, 513, 2049, 100000, 1000003])
@pytest.mark.parametrize('kernel', [0, 1])
def test_sizes_and_kernels(n, kernel):
    rng = np.random.default_rng(n)
    x = rng.normal(-385, 7.7, n).astype(np.float32)
    G = 1000 if n < 1000000 else 200
    want = epdf(x.astype(np.float64), -400, -370, 3, G, 'epanechnikov' if kernel == 0 else 'gaussian')
    got = _dev_kde(x, -400, -370, 3, G, kernel)
    assert np.max(np.abs(got - want)) < 2e-5 * max(want.max(), 1e-30)


def test_grid_shapes_and_errors():
    x = torch.randn(1000, device='cuda')
    for G in (2, 7, 1000, 1024, 1025, 3000):
        d = torch.ops.at2.kde(x, -4.0, 4.0, G, 0.5, 0)
        want = epdf(x.cpu().numpy().astype(np.float64), -4, 4, 0.5, G)
        assert d.shape == (G,) and np.max(np.abs(d.cpu().numpy() - want)) < 1e-5
    with pytest.raises(RuntimeError):
        torch.ops.at2.kde(x, -4.0, 4.0, 1000, 0.0, 0)        # bandwidth must be positive
    with pytest.raises(RuntimeError):
        torch.ops.at2.kde(x.cpu(), -4.0, 4.0, 1000, 0.5, 0)  # no CPU fallback
    # determinism: the partial sums are added as fixed-point integers (order-independent)
    assert torch.equal(torch.ops.at2.kde(x, -4.0, 4.0, 1000, 0.5, 0), torch.ops.at2.kde(x, -4.0, 4.0, 1000, 0.5, 0))
