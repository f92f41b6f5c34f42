"""Oracle closest-known-loss-function lookup vs outputs of the reference's KnownFnEvaluator (tests/golden/knownfn.npz)."""
import numpy as np
import pytest

from oracle import known_fn as okf
from tests.helpers import GOLDEN

DOMS = {'mnist': (okf.MNIST_DOMAIN, 425), 'cifar': (okf.CIFAR_DOMAIN, 300)}


def _db(tag):
    dom, n = DOMS[tag]
    arms, curves = okf.synthetic_db(dom, n, 300, seed=0)
    return dom, arms, curves


@pytest.mark.parametrize('tag', ['mnist', 'cifar'])
def test_nearest_index_and_loss_match_reference(tag):
    g = np.load(f'{GOLDEN}/knownfn.npz')
    dom, arms, curves = _db(tag)
    names = okf.attribute_order(dom)
    assert list(g[f'{tag}_names']) == names
    assert np.array_equal(g[f'{tag}_db'], np.asarray([[float(a[n]) for n in names] for a in arms]))
    for k, (q, want_i, want_l) in enumerate(zip(g[f'{tag}_query'], g[f'{tag}_query_idx'], g[f'{tag}_query_loss'])):
        arm = dict(zip(names, q))
        i, _ = okf.nearest(arm, arms, dom)
        assert i == want_i
        assert okf.evaluate(arm, arms, curves, dom, (1, 3, 9, 27, 81)[k % 5]) == want_l


@pytest.mark.parametrize('tag', ['mnist', 'cifar'])
@pytest.mark.parametrize('R', [81, 6])
def test_hyperband_on_known_functions_matches_reference(tag, R):
    g = np.load(f'{GOLDEN}/knownfn.npz')
    dom, arms, curves = _db(tag)
    names = okf.attribute_order(dom)
    for rep, (ofe, ofe_id) in enumerate(zip(g[f'{tag}_hb_R{R}_ofe'], g[f'{tag}_hb_R{R}_ofe_id'])):
        qs = [dict(zip(names, row)) for row in g[f'{tag}_hb_R{R}_arms'][rep]]
        out = okf.run_hyperband(arms, curves, dom, R, 3, arms=qs)
        assert out['ofe'] == ofe and out['ofe_id'] == ofe_id
