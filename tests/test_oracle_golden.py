"""The oracle (CPU restatement) against outputs of the reference itself (tests/golden/, made by make_golden.py)."""
import random

import numpy as np
import pytest

from oracle import gamma_process as gp
from oracle import hyperband as ohb
from tests.helpers import GOLDEN, hb_case_names, load_hb_case, load_seeded


def test_bracket_tables_match_reference():
    d = np.load(f'{GOLDEN}/bracket_tables.npz')
    keys = sorted({k.rsplit('_', 2)[0] if k.endswith(('round_n', 'round_keep', 'round_res')) else k.rsplit('_', 1)[0]
                   for k in d.files})
    pairs = sorted({tuple(int(v[1:] if v[0] == 'R' else v[3:]) for v in k.split('_')[:2]) for k in d.files})
    assert (81, 3) in pairs and (243, 3) in pairs and len(keys) > 0
    for R, eta in pairs:
        pre = f'R{R}_eta{eta}_'
        plan = ohb.bracket_plan(R, eta)
        assert [b['n'] for b in plan] == d[pre + 'n'].tolist()
        rounds = [rd for b in plan for rd in b['rounds']]
        assert [rd['n_eval'] for rd in rounds] == d[pre + 'round_n'].tolist()
        assert [rd['keep'] for rd in rounds] == d[pre + 'round_keep'].tolist()
        assert [rd['resource'] for rd in rounds] == d[pre + 'round_res'].tolist()   # bit-exact floats
    # R=81 -> 117 arms, 12 rounds; R=243 -> 369 arms, 18 rounds
    assert ohb.plan_arms(ohb.bracket_plan(81, 3)) == 117
    assert ohb.plan_arms(ohb.bracket_plan(243, 3)) == 369
    assert sum(len(b['rounds']) for b in ohb.bracket_plan(243, 3)) == 18


@pytest.mark.parametrize('name', hb_case_names())
def test_predrawn_repetitions_bit_exact(name):
    cfg, reps = load_hb_case(name)
    for rep in reps:
        out = ohb.run_predrawn(cfg['func'], cfg['families'], cfg['R'], cfg['eta'], cfg['noise'], rep['xy'],
                               rep['fam'], rep['z'], rep['gamma'], cfg['min_or_max'])
        assert np.array_equal(out['traj_raw'], rep['traj_raw'])          # bit-exact float64 trajectories
        assert np.array_equal(out['ckpts'], rep['ckpts'])
        assert np.array_equal(out['ckpt_loss'], rep['ckpt_loss'])
        if 'traj' in rep:
            assert np.array_equal(out['traj'], rep['traj'])
        for k in ('eval_ids', 'round_off', 'surv_ids', 'surv_off', 'keep', 'round_best_id', 'ofe_id'):
            assert np.array_equal(out[k], rep[k]), k                     # index work: bit-exact
        for k in ('eval_fvals', 'round_best', 'ofe'):
            assert np.array_equal(out[k], rep[k]), k


def test_constant_loss_first_arm_wins():
    """The tie-break rule of the reference's tests/hyperband_test.py:62: with every loss equal, the best arm is the
    first evaluation in history (= arm 0)."""
    cfg, reps = load_hb_case('branin_flat_R81')
    rep = reps[1]
    assert np.ptp(rep['ckpt_loss']) == 0
    out = ohb.run_predrawn(cfg['func'], cfg['families'], cfg['R'], cfg['eta'], cfg['noise'], rep['xy'], rep['fam'],
                           rep['z'], rep['gamma'])
    assert out['ofe_id'] == 0 == rep['ofe_id']
    assert out['surv_ids'][:27].tolist() == list(range(27))


@pytest.mark.parametrize('name', sorted(load_seeded()))
def test_seeded_global_rng_runs_bit_exact(name):
    cfg, ofes = load_seeded()[name]
    np.random.seed(cfg['seed'])
    random.seed(cfg['seed'])
    got = [ohb.run_global_rng(cfg['func'], cfg['families'], cfg['R'], cfg['eta'], cfg['noise'], cfg['min_or_max'],
                              cfg['scheduler']) for _ in range(cfg['n_reps'])]
    assert np.array_equal(np.asarray(got), ofes)


def test_savgol_operator_rows_match_filter():
    rng = np.random.default_rng(0)
    for R in (27, 81, 243):
        W = gp.savgol_operator(R)
        x = rng.normal(size=R) * 100
        ref = np.asarray(gp.readout(list(x), True, R))
        assert np.max(np.abs(W @ x - ref)) < 1e-9 * 100
    with pytest.raises(ValueError):
        gp.savgol_operator(6)
    with pytest.raises(ValueError):
        gp.readout([0.0] * 6, True, 6)


def test_gamma_parameters():
    # alpha in (1, 3.74], mode (alpha-1)*scale == 2 for every t
    for R in (27, 81, 243):
        for t in range(1, R):
            a, s = gp.gamma_shape_scale(t, R)
            assert 1 < a <= 3.7321
            assert abs((a - 1) * s - 2) < 1e-12


def test_vectorised_numpy_port_matches_reference_ofe_distributions():
    """oracle/vectorised.py (bench.py's `cpu_baseline.vectorised_numpy`) against the reference's stored OFE vectors."""
    from scipy.stats import ks_2samp

    from oracle import vectorised as ov
    fam6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
            (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
    g = np.load(f'{GOLDEN}/epdf_ofes.npz')
    for key, func, fams, noise in (('simulation-rastrigin-families-2/hist-sim-rastrigin-sim(hb)-7000-1', 'rastrigin', fam6, 10),
                                   ('simulation-flat-branin/hist-sim(hb)-7000-1', 'branin', [(0, 1, 0, False, 0, 0)], 0),
                                   ('simulation-flat-drop-wave/hist-sim-wave-sim(hb)-7000-1', 'wave', [(0, 1, 0, False, 0, 0)], 0)):
        got = ov.run_batch(func, fams, 81, 3, noise, 1500, np.random.default_rng(5))
        assert ks_2samp(got, g[key]).pvalue > 0.01, key
