"""Edge cases of the device path, each with what the reference does in the same situation: empty and single-repetition
jobs, ragged recorded curves, an empty database, degenerate KDE inputs, the smallest plans, every scheduler / objective."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import autotune.b200  # noqa: F401
from autotune.b200.engine import (HyperbandSimSpec, KnownFnSpec, run_hybrid_repetitions, run_hyperband_repetitions,
                                  run_knownfn_hyperband, run_random_repetitions, run_tpe_repetitions)
from autotune.b200.epdf import density, run_epdf_ofe, summary_stats
from autotune.benchmarks.mnist_like import mnist_domain_problem, synthetic_database
from autotune.core import Arm

pytestmark = pytest.mark.gpu

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]


def _mnist_db(n_arms=40, length=300):
    with redirect_stdout(io.StringIO()):
        prob = mnist_domain_problem()
    return prob, synthetic_database(prob, n_arms, length, seed=1)


def test_empty_jobs_give_empty_results():
    """`for _ in range(0)` in the reference's drivers leaves `optimums == []`."""
    spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
    for out in (run_hyperband_repetitions(spec, 0, details=True), run_hyperband_repetitions(spec, 0, early_stop=True),
                run_hybrid_repetitions(spec, 0, 'all', details=True), run_tpe_repetitions(spec, 0, 24, 81)):
        assert out['ofe'].shape == (0,) and out['best_xy'].shape == (0, 2)
    assert run_hyperband_repetitions(spec, 0, details=True)['survivors'].shape == (0, spec.plan.n_surv)
    prob, known = _mnist_db()
    ks = KnownFnSpec(known, prob.domain, prob.hyperparams_to_opt)
    assert run_knownfn_hyperband(ks, 81, 3, 0)['ofe'].shape == (0,)
    with pytest.raises(Exception):
        run_hyperband_repetitions(spec, -1)


def test_single_repetition_equals_the_first_of_many():
    spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
    many = run_hyperband_repetitions(spec, 100, seed=3, details=True)
    for runner, kw in ((run_hyperband_repetitions, {}), (run_hyperband_repetitions, {'early_stop': True})):
        one = runner(spec, 1, seed=3, details=True, **kw)
        for k in one:
            assert torch.equal(one[k], many[k][:1]), k
    hy = run_hybrid_repetitions(spec, 50, 'same', seed=3)
    assert torch.equal(run_hybrid_repetitions(spec, 1, 'same', seed=3, rep_offset=49)['ofe'], hy['ofe'][49:])


def test_ragged_and_empty_recorded_curves():
    """Recorded curves may have different lengths (the reference keeps python lists); reading beyond the shortest one
    raises IndexError there, on whichever arm happens to be nearest - here up front.  An empty database cannot be scanned."""
    prob, known = _mnist_db(40, 100)
    arms = list(known)
    known[arms[3]] = known[arms[3]][:30]                       # a short curve
    ks = KnownFnSpec(known, prob.domain, prob.hyperparams_to_opt)
    out = run_knownfn_hyperband(ks, 27, 3, 200, seed=2, details=True)      # reads positions <= 27 < 30: fine
    flat = np.concatenate([np.asarray(v, np.float64) for v in known.values()])
    assert np.isin(out['ofe'].cpu().numpy(), flat).all() and torch.isfinite(out['ckpt_loss']).all()
    with pytest.raises(IndexError):
        run_knownfn_hyperband(ks, 81, 3, 10)                   # 81 > 30
    with pytest.raises(ValueError):
        KnownFnSpec({}, prob.domain, prob.hyperparams_to_opt)
    # a one-arm database: every proposal maps to it
    ks1 = KnownFnSpec({arms[0]: known[arms[0]]}, prob.domain, prob.hyperparams_to_opt)
    out1 = run_knownfn_hyperband(ks1, 27, 3, 50, seed=2, details=True)
    assert (out1['nn_idx'] == 0).all() and torch.equal(out1['ofe'], torch.full_like(out1['ofe'], min(known[arms[0]][:27])))
    # an exact hit returns distance 0 and that arm
    idx, dist = ks.lookup([[float(arms[7][n]) for n in ks.names]])
    assert int(idx[0]) == 7 and float(dist[0]) == 0.0
    assert isinstance(arms[7], Arm)


def test_kde_degenerate_inputs():
    """sklearn's KernelDensity on one sample is the kernel itself; samples outside the grid contribute nothing."""
    one = density(torch.tensor([0.5], device='cuda'), 0.0, 1.0, 0.25, 1000)
    grid = np.linspace(0.0, 1.0, 1000)
    want = 0.75 * np.maximum(0.0, 1 - ((grid - 0.5) / 0.25) ** 2) / 0.25
    np.testing.assert_allclose(one.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    far = density(torch.full((1000,), 50.0, device='cuda'), 0.0, 1.0, 0.25, 1000)
    assert float(far.abs().max()) == 0.0
    tiny_grid = density(torch.randn(100, device='cuda'), -1.0, 1.0, 0.5, 2)
    assert tiny_grid.shape == (2,) and torch.isfinite(tiny_grid).all()
    with pytest.raises(Exception):
        density(torch.empty(0, device='cuda'), 0.0, 1.0, 0.25, 1000)       # sklearn: "Found array with 0 sample(s)"
    stats = summary_stats(torch.tensor([3.0, 1.0, 2.0, 4.0], device='cuda'))
    assert stats['avg_optimum'] == 2.5 and stats['avg_top_25%'] == 1.0 and stats['top_quartile'].tolist() == [1.0]


@pytest.mark.parametrize('R,eta', [(2, 2), (3, 3), (4, 2), (9, 3)])
def test_smallest_plans(R, eta):
    """s_max < 2 keeps every bracket (hyperband_optimiser.py:71); R = 9, eta = 3 is the first plan that drops brackets."""
    spec = HyperbandSimSpec('branin', [(0.9, 10, 0.1, False, 0, 200)], R, eta, 1)
    out = run_hyperband_repetitions(spec, 300, seed=4, details=True)
    assert torch.isfinite(out['ofe']).all() and out['round_best'].shape[1] == spec.plan.n_rounds
    assert torch.equal(out['ofe'], out['round_best'].min(dim=1).values)
    lazy = run_hyperband_repetitions(spec, 300, seed=4, details=True, early_stop=True)
    assert torch.equal(lazy['ofe'], out['ofe']) and torch.equal(lazy['survivors'], out['survivors'])
    rnd = run_random_repetitions(spec, 100, R, 5, seed=4)
    assert rnd['ofe'].shape == (100,)


def test_epdf_job_with_max_objective_and_round_robin():
    spec = HyperbandSimSpec('wave', FAM6, 27, 3, 5, scheduler='round-robin', min_or_max=max)
    res = run_epdf_ofe(spec, 3000, seed=6, start=-450.0, end=50.0, bandwidth=10.0)
    ofe = res['ofes']
    d = run_hyperband_repetitions(spec, 3000, seed=6, details=True)
    assert torch.equal(ofe, d['ofe']) and torch.equal(ofe, d['round_best'].max(dim=1).values)
    integral = float(res['density'].double().sum() * 500.0 / 999)
    assert 0.9 < integral < 1.01
