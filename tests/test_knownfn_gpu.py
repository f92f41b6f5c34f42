"""Kernel (d): nearest recorded arm + Hyperband on the closest-known-function problem, against reference goldens
(tests/golden/knownfn.npz, produced by the reference's KnownFnEvaluator / HyperbandOptimiser) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import known_fn as okf
from tests.helpers import GOLDEN

import autotune.b200  # noqa: F401,E402  (registers torch.ops.at2.*)

pytestmark = pytest.mark.gpu

DOMS = {'mnist': (okf.MNIST_DOMAIN, 425, ()), 'cifar': (okf.CIFAR_DOMAIN, 300,
                                                        ('learning_rate', 'n_units_1', 'n_units_2', 'n_units_3', 'batch_size'))}


def _product_domain(odom):
    from autotune.core import Domain, Param
    return Domain(**{p['name']: Param(p['name'], p['min_val'], p['max_val'], init_val=p['init_val'], distrib='uniform',
                                      scale=p['scale'], interval=p['interval']) for p in odom})


def _spec(tag):
    from autotune.b200.engine import KnownFnSpec
    from autotune.core import Arm
    odom, n, to_opt = DOMS[tag]
    arms, curves = okf.synthetic_db(odom, n, 300, seed=0)
    attr = okf.attribute_order(odom)
    known_fs = {Arm(**{k: a[k] for k in attr}): list(c) for a, c in zip(arms, curves)}
    dom = _product_domain(odom)
    to_opt = to_opt or tuple(dom.hyperparams_names())
    return KnownFnSpec(known_fs, dom, to_opt), arms, curves, odom, known_fs, dom, to_opt


def _domain_order(g, tag, rows, spec):
    """golden rows are in attribute order (g[tag_names]); the device wants domain order."""
    names = list(g[f'{tag}_names'])
    return rows[..., [names.index(n) for n in spec.names]]


@pytest.mark.parametrize('tag', ['mnist', 'cifar'])
def test_nearest_index_bit_exact_vs_reference(tag):
    g = np.load(f'{GOLDEN}/knownfn.npz')
    spec, *_ = _spec(tag)
    q = _domain_order(g, tag, g[f'{tag}_query'], spec)
    times = np.asarray([(1, 3, 9, 27, 81)[k % 5] for k in range(len(q))], np.int32)
    idx, dist, loss = spec.lookup(q, times)
    assert np.array_equal(idx.cpu().numpy(), g[f'{tag}_query_idx'])          # bit-exact index work
    assert np.array_equal(loss.cpu().numpy(), g[f'{tag}_query_loss'])


@pytest.mark.parametrize('tag', ['mnist', 'cifar'])
def test_lookup_against_oracle_with_ties_and_many_queries(tag):
    spec, arms, curves, odom, *_ = _spec(tag)
    rng = np.random.default_rng(3)
    db = spec.db.cpu().numpy()
    # queries: random points, exact recorded arms (distance 0), and midpoints of two recorded arms
    q = np.concatenate([db[rng.integers(0, len(db), 40)], 0.5 * (db[:20] + db[20:40]),
                        db[rng.integers(0, len(db), 40)] * (1 + 1e-9 * rng.normal(size=(40, db.shape[1])))])
    idx, dist = spec.lookup(q)
    for k, row in enumerate(q):
        arm = {n: row[spec.names.index(n)] for n in okf.attribute_order(odom)}
        i, d = okf.nearest(arm, arms, odom)
        # the index is bit-exact; the distance itself may differ in the last bit because python's `** 2` goes through
        # libm pow(), which is not correctly rounded for every input, while the device squares exactly
        assert int(idx[k]) == i and abs(float(dist[k]) - d) <= 4e-16 * max(d, 1e-300)
    # few queries -> the database is split over CTAs; many queries -> one chunk: same answer
    big = np.tile(q, (400, 1))
    idx_big, _ = spec.lookup(big)
    assert torch.equal(idx_big.view(400, -1), idx.expand(400, -1))


def test_duplicate_recorded_arms_first_wins():
    from autotune.b200.engine import KnownFnSpec
    from autotune.core import Arm
    dom = _product_domain(okf.MNIST_DOMAIN)
    names = list(dom.hyperparams_names())

    class A2(Arm):      # distinct dictionary keys with identical hyperparameters
        def __hash__(self):
            return id(self)

        def __eq__(self, other):
            return self is other
    vals = dict(learning_rate=0.01, weight_decay=0.001, momentum=0.5, batch_size=100)
    known_fs = {A2(**vals): [float(t)] * 10 for t in range(5)}
    assert len(known_fs) == 5
    spec = KnownFnSpec(known_fs, dom, tuple(names))
    idx, _ = spec.lookup([[vals[n] for n in names]] * 3)
    assert idx.tolist() == [0, 0, 0]


@pytest.mark.parametrize('tag', ['mnist', 'cifar'])
@pytest.mark.parametrize('R', [81, 6])
def test_hyperband_predrawn_matches_reference(tag, R):
    from autotune.b200.engine import run_knownfn_hyperband
    g = np.load(f'{GOLDEN}/knownfn.npz')
    spec, arms, curves, odom, *_ = _spec(tag)
    pre = _domain_order(g, tag, g[f'{tag}_hb_R{R}_arms'], spec)
    out = run_knownfn_hyperband(spec, R, 3, pre.shape[0], predrawn=pre, details=True)
    assert np.array_equal(out['ofe'].cpu().numpy(), g[f'{tag}_hb_R{R}_ofe'])
    assert np.array_equal(out['ofe_arm'].cpu().numpy(), g[f'{tag}_hb_R{R}_ofe_id'])
    attr = okf.attribute_order(odom)
    for rep in range(pre.shape[0]):
        qs = [dict(zip(attr, row)) for row in g[f'{tag}_hb_R{R}_arms'][rep]]
        want = okf.run_hyperband(arms, curves, odom, R, 3, arms=qs)
        assert np.array_equal(out['nn_idx'][rep].cpu().numpy(), want['nn_idx'])
        assert np.array_equal(out['survivors'][rep].cpu().numpy(), want['surv_ids'])
        assert np.array_equal(out['round_best_arm'][rep].cpu().numpy(), want['round_best_id'])
        assert np.array_equal(out['round_best'][rep].cpu().numpy(), want['round_best'])


def test_hyperband_philox_distribution_and_properties():
    """10^4 repetitions (BASELINE config 4) with in-kernel draws: the drawn arms lie in the domain, every OFE is a value
    of some recorded curve, and the OFE distribution matches fresh oracle runs (KS p > 0.01)."""
    import random

    from scipy.stats import ks_2samp

    from autotune.b200.engine import run_knownfn_hyperband
    spec, arms, curves, odom, *_ = _spec('mnist')
    out = run_knownfn_hyperband(spec, 81, 3, 10000, seed=2, details=False)
    ofe = out['ofe'].cpu().numpy()
    assert np.isin(ofe, curves).all()
    v = out['best_arm_vals'].cpu().numpy()
    for h, p in enumerate(odom):
        lo, hi = (np.exp(p['min_val']), np.exp(p['max_val'])) if p['scale'] == 'log' else (p['min_val'], p['max_val'])
        assert (v[:, h] >= lo - 1e-9).all() and (v[:, h] <= hi + 1e-9).all()
    assert np.array_equal(v[:, 3], np.floor(v[:, 3]))          # batch_size has interval 1
    np.random.seed(11)
    random.seed(11)
    want = np.asarray([okf.run_hyperband(arms, curves, odom, 81, 3)['ofe'] for _ in range(150)])
    res = ks_2samp(ofe, want)
    assert res.pvalue > 0.01, res
    # sharding invariance
    a = run_knownfn_hyperband(spec, 81, 3, 600, seed=2, rep_offset=100)['ofe']
    assert torch.equal(a, out['ofe'][100:700])


def test_problem_api_evaluate_matches_reference_semantics():
    from autotune.benchmarks import KnownFnProblem
    from autotune.core import Arm, HyperparameterOptimisationProblem
    spec, arms, curves, odom, known_fs, dom, to_opt = _spec('mnist')

    class Real(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise NotImplementedError
    problem = KnownFnProblem(known_fs=known_fs, real_problem=Real(dom, to_opt))
    g = np.load(f'{GOLDEN}/knownfn.npz')
    names = list(g['mnist_names'])
    for k in range(10):
        arm = Arm(**dict(zip(names, g['mnist_query'][k])))
        goals = problem.get_evaluator(arm=arm).evaluate((1, 3, 9, 27, 81)[k % 5])
        assert goals.fval == g['mnist_query_loss'][k]
        assert goals.fvals == list(curves[g['mnist_query_idx'][k]])
    ev = problem.get_evaluator()          # random arm from the domain
    assert 0 < ev.evaluate(27).fval < 1.1


def test_evaluator_accepts_any_attribute_order_of_the_proposed_arm():
    """The reference sums the squared-error terms in the attribute order of the PROPOSED arm (known_loss_fn_problem.py:36-48)
    and accepts whatever order a caller's Arm has.  A user-built Arm with the hyperparameters in another order must be
    looked up with that summation order, not refused: same nearest arm as the oracle run with the permuted order."""
    from autotune.benchmarks.known_loss_fn_problem import KnownFnEvaluator
    from autotune.core import Arm
    spec, arms, curves, odom, known_fs, dom, to_opt = _spec('mnist')
    rng = np.random.default_rng(11)
    db = spec.db.cpu().numpy()
    names = list(spec.names)
    for trial in range(12):
        row = db[rng.integers(0, len(db))] * (1 + 0.05 * rng.normal(size=db.shape[1]))
        row = np.clip(row, [np.exp(p['min_val']) if p['scale'] == 'log' else p['min_val'] for p in odom],
                      [np.exp(p['max_val']) if p['scale'] == 'log' else p['max_val'] for p in odom])
        perm = list(rng.permutation(len(names)))
        arm = Arm(**{names[i]: float(row[i]) for i in perm})
        assert list(arm.__dict__.keys()) == [names[i] for i in perm]
        ev = KnownFnEvaluator(known_fs, arm, dom, _device_db=spec)
        got = ev.evaluate(9)
        want_idx, _ = okf.nearest({names[i]: float(row[i]) for i in perm}, arms, odom)
        assert got.fvals == list(curves[want_idx]) and got.fval == curves[want_idx][8]
