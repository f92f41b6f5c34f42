"""Oracle post-processing vs the reference's own profiles.py / cross_validate_approximation.py outputs."""
import numpy as np

from oracle import known_fn as okf
from oracle import postprocessing as opp
from tests.helpers import GOLDEN


def test_order_profiles_match_reference():
    g = np.load(f'{GOLDEN}/postprocessing.npz')
    c = g['prof_curves']
    assert np.allclose(opp.dynamic_order_profile(c), g['prof_dynamic'], rtol=0, atol=1e-15)
    assert abs(opp.ends_order_profile(c) - g['prof_ends']) < 1e-15


def test_cross_validation_matches_reference():
    g = np.load(f'{GOLDEN}/postprocessing.npz')
    arms, curves = okf.synthetic_db(okf.MNIST_DOMAIN, 103, 60, seed=1)
    res, min_len = opp.cross_validate(10, arms, curves, okf.MNIST_DOMAIN)
    assert min_len == g['cv_min_len'] and res.shape == g['cv_res'].shape
    assert np.array_equal(res, g['cv_res'])
