"""Post-processing layer on the device ("next" rows of SURVEY 8f) vs reference-generated goldens and scipy."""
import numpy as np
import pytest
import torch

import autotune.b200  # noqa: F401
from oracle import known_fn as okf
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_family_profiles_match_reference():
    from autotune.experiments.simulation_evaluation import profiles
    g = np.load(f'{GOLDEN}/postprocessing.npz')
    c = g['prof_curves']                       # values are float32-representable: the device compares the same numbers
    assert np.allclose(profiles.get_avg_profile(c), g['prof_avg'], rtol=1e-12)
    assert np.allclose(profiles.get_med_profile(c), g['prof_med'], rtol=1e-12)
    assert np.allclose(profiles.get_std_profile(c), g['prof_std'], rtol=1e-12)
    assert np.allclose(profiles.get_dynamic_order_profile(c), g['prof_dynamic'], atol=1e-6)   # exact counts / float32 ratio
    assert abs(profiles.get_ends_order_profile(c) - g['prof_ends']) < 1e-6


def test_plot_simulated_curves():
    from autotune.core import ShapeFamily
    from autotune.experiments.simulation_evaluation import profiles
    fams = (ShapeFamily(None, 1.5, 10, 15, False), ShapeFamily(None, 0.2, 4, 7, True))
    curves = np.asarray(profiles.plot_simulated('branin', 64, 81, shape_families=fams, init_noise=10))
    assert curves.shape == (64, 81) and np.isfinite(curves).all()
    from autotune.benchmarks.opt_function_problem import branin_value  # noqa: F401
    assert (curves[:, -1] < curves[:, 0] - 100).all()          # every curve descends by ~end_shift = 200
    dyn = profiles.get_dynamic_order_profile(curves)
    assert len(dyn) == 80 and 0.5 < dyn[-1] <= 1.0


def test_cross_validation_matches_reference():
    from autotune.core import Arm, Domain, HyperparameterOptimisationProblem, Param
    from autotune.experiments.known_functions_evaluation.cross_validate_approximation import cross_validate
    g = np.load(f'{GOLDEN}/postprocessing.npz')
    odom = okf.MNIST_DOMAIN
    arms, curves = okf.synthetic_db(odom, 103, 60, seed=1)
    dom = Domain(**{p['name']: Param(p['name'], p['min_val'], p['max_val'], init_val=p['init_val'], distrib='uniform',
                                     scale=p['scale'], interval=p['interval']) for p in odom})
    known_fs = {Arm(**a): list(c) for a, c in zip(arms, curves)}

    class Real(HyperparameterOptimisationProblem):
        def get_evaluator(self, arm=None):
            raise NotImplementedError
    res, min_len = cross_validate(10, known_fs, Real(dom))
    assert min_len == g['cv_min_len']
    assert np.array_equal(res, g['cv_res'])                   # same nearest arms, same float64 arithmetic


@pytest.mark.parametrize('n1,n2', [(7000, 7000), (100, 20000), (1, 5), (100000, 7000)])
def test_ks_statistic_matches_scipy(n1, n2):
    from scipy.stats import ks_2samp as ref_ks

    from autotune.experiments.simulation_evaluation.stats import ks_2samp
    rng = np.random.default_rng(n1 + n2)
    a = rng.normal(0, 1, n1).astype(np.float32)
    b = (rng.normal(0.03, 1.02, n2)).astype(np.float32)
    b[: min(n2, 50)] = a[: min(n1, 50)][: min(n2, 50)] if n1 >= 50 and n2 >= 50 else b[: min(n2, 50)]   # ties across samples
    got = ks_2samp(a, b)
    want = ref_ks(a.astype(np.float64), b.astype(np.float64), method='asymp')
    assert abs(got.statistic - want.statistic) < 1e-6
    assert abs(got.pvalue - want.pvalue) < 1e-5 + 1e-3 * want.pvalue


def test_ks_on_reference_vectors():
    """The KS comparisons the reference reports between its own methods: HB vs ALL D=0.068, ALL vs SAME 0.020."""
    from autotune.experiments.simulation_evaluation.stats import ks_2samp
    g = np.load(f'{GOLDEN}/epdf_ofes.npz')
    pre = 'simulation-rastrigin-families-2/hist-sim-rastrigin-sim'
    hb, all_, same = (g[f'{pre}({m})-7000-1'] for m in ('hb', 'hb+tpe+transfer+all', 'hb+tpe+transfer+same'))
    assert abs(ks_2samp(hb, all_).statistic - 0.068) < 2e-3
    assert abs(ks_2samp(all_, same).statistic - 0.020) < 2e-3


def test_results_comparison_driver_on_the_references_stored_vectors(tmp_path):
    """experiments/simulation_evaluation/plot_run_simulations_results.py on the reference's own OFE vectors (tests/golden/
    epdf_ofes.npz): device histograms == numpy's, KS statistics == scipy's, EPDF == the closed form; the reference's
    finding 'hybrids differ from Hyperband, hybrids resemble each other'  comes out."""
    import pickle

    from scipy.stats import ks_2samp as scipy_ks

    from autotune.experiments.simulation_evaluation import plot_run_simulations_results as drv
    from autotune.experiments.simulation_evaluation.plot_results import histograms
    from oracle import kde as okde
    g = np.load(f'{GOLDEN}/epdf_ofes.npz')
    prefix = 'simulation-rastrigin-families-2/hist-sim-rastrigin-'
    stored = {m: g[f'{prefix}{m}-7000-1'] for _, m in drv.METHODS if f'{prefix}{m}-7000-1' in g.files}
    assert 'sim(hb)' in stored and 'sim(hb+tpe+transfer+all)' in stored
    for method, vals in stored.items():                                     # the reference's file naming, -til suffix
        with open(tmp_path / f'hist-sim-rastrigin-{method}-7000-1.pkl', 'wb') as f:
            pickle.dump({'method': method, 'n_simulations': 7000, 'all_optimums': list(vals)}, f)
    out = drv.main(['--simulation', 'rastrigin-families-2', '--dir', str(tmp_path), '-n', '7000'])
    label_of = {m: lab for lab, m in drv.METHODS}
    assert set(out['labels']) == {label_of[m] for m in stored}
    bins = np.asarray(out['bins'])
    assert bins[0] == -400 and len(bins) == 31
    for lab, dens in zip(out['labels'], out['histograms']):
        vals = stored[{v: k for k, v in label_of.items()}[lab]]
        want, _ = np.histogram(vals, bins=bins, density=True)
        np.testing.assert_allclose(dens, want, rtol=1e-12)
    counts = histograms([stored['sim(hb)']], bins)[0]
    assert np.array_equal(counts, np.histogram(stored['sim(hb)'], bins=bins)[0])
    for (a, b), (stat, p) in out['ks'].items():
        ma, mb = ({v: k for k, v in label_of.items()}[x] for x in (a, b))
        ref = scipy_ks(stored[ma], stored[mb])
        assert stat == pytest.approx(ref.statistic, abs=1e-6) and p == pytest.approx(ref.pvalue, rel=0.05, abs=1e-12)
    assert out['ks'][('Hyperband', 'ALL')][1] < 1e-6                          # HB vs hybrids: distinguishable
    if ('ALL', 'SAME') in out['ks']:
        assert out['ks'][('ALL', 'SAME')][1] > 0.01                           # hybrids vs hybrids: not
    want = okde.epdf(stored['sim(hb)'], -400, -370, 3.0)
    np.testing.assert_allclose(out['epdf'][out['labels'].index('Hyperband')], want, rtol=2e-4, atol=1e-6)
