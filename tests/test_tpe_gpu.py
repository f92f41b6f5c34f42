"""Kernel (c): Parzen-estimator arithmetic against the oracle restatement of hyperopt 0.2.4 (oracle/tpe.py - parity
unpinned at the arithmetic level, see its header), and the Hyperband-TPE hybrids against the reference's stored OFE
vectors (the only pins that exist: distributional)."""
import numpy as np
import pytest
import torch

from oracle import tpe as otpe
from tests.helpers import GOLDEN

import autotune.b200  # noqa: F401,E402  (registers torch.ops.at2.*)

pytestmark = pytest.mark.gpu

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]


def _tpe_bytes(**kw):
    from autotune.b200 import ops
    from autotune.b200.engine import tpe_config
    return ops.struct_tensor(tpe_config(**kw))


@pytest.mark.parametrize('n_obs', [10, 16, 17, 26, 40, 65, 100, 200])
def test_parzen_models_and_scores_match_oracle(n_obs):
    rng = np.random.default_rng(n_obs)
    n_prob, n_cand, lo, hi = 12, 48, -5.12, 5.12
    obs = rng.uniform(lo, hi, (n_prob, n_obs)).astype(np.float32)
    obs[1, :] = np.sort(obs[1, :])                      # already sorted
    obs[2, :] = obs[2, 0]                               # all observations identical (bandwidth floor)
    obs[3, : n_obs // 2] = lo + 1e-3                    # clustered at the lower bound
    loss = rng.normal(size=(n_prob, n_obs)).astype(np.float32)
    loss[4, :] = 1.0                                    # all losses tie: the oldest trials are "good"
    cand = rng.uniform(lo, hi, (n_prob, n_cand)).astype(np.float32)
    scores, models, lean, _ = torch.ops.at2.tpe_score(torch.from_numpy(obs).cuda(), torch.from_numpy(loss).cuda(), lo, hi,
                                                      _tpe_bytes(), torch.from_numpy(cand).cuda())
    scores, models, lean = scores.cpu().numpy(), models.cpu().numpy(), lean.cpu().numpy()
    # inside a suggestion the simulation kernel skips the weight normalisation and the truncation mass: a per-problem
    # constant shift of the scores, the same arg-max
    shift = lean - scores
    assert np.all(np.abs(shift - shift.mean(axis=1, keepdims=True)) < 2e-5), np.abs(shift - shift.mean(axis=1, keepdims=True)).max()
    top2 = np.sort(scores, axis=1)[:, -2:]
    clear = top2[:, 1] - top2[:, 0] > 5e-5
    assert np.array_equal(np.argmax(lean, axis=1)[clear], np.argmax(scores, axis=1)[clear]) and clear.sum() >= 4
    for k in range(n_prob):
        below, above = otpe.split_trials(obs[k].astype(np.float64), loss[k].astype(np.float64))
        for which, vals in enumerate((below, above)):
            w, mu, sg = otpe.adaptive_parzen_normal(vals, 1.0, 0.5 * (hi + lo), hi - lo)
            got = models[k, which, :len(w)]
            np.testing.assert_allclose(got[:, 0], w, rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(got[:, 1], mu, rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(got[:, 2], sg, rtol=2e-5, atol=1e-6)
        wb, mb, sb = otpe.adaptive_parzen_normal(below, 1.0, 0.5 * (hi + lo), hi - lo)
        wa, ma, sa = otpe.adaptive_parzen_normal(above, 1.0, 0.5 * (hi + lo), hi - lo)
        want = otpe.gmm1_lpdf(cand[k].astype(np.float64), wb, mb, sb, lo, hi) - \
            otpe.gmm1_lpdf(cand[k].astype(np.float64), wa, ma, sa, lo, hi)
        assert np.max(np.abs(scores[k] - want)) < 2e-5, (k, np.max(np.abs(scores[k] - want)))
        assert np.argmax(scores[k]) == np.argmax(want) or np.sort(want)[-1] - np.sort(want)[-2] < 5e-5


def _spec(func, fams, R=81, noise=10, mom='min'):
    from autotune.b200.engine import HyperbandSimSpec
    return HyperbandSimSpec(func, fams, R, 3, noise, 'uniform', mom)


@pytest.mark.parametrize('func,fams,R,noise,mom', [('rastrigin', FAM6, 81, 10, 'min'), ('branin', FLAT, 81, 0, 'min'),
                                                   ('wave', FAM6, 27, 10, 'max'), ('egg', FAM6[:3], 243, 10, 'min')])
def test_tpe_disabled_is_bit_identical_to_hyperband_kernel(func, fams, R, noise, mom):
    from autotune.b200.engine import run_hyperband_repetitions, run_optimiser_repetitions, tpe_config
    spec = _spec(func, fams, R, noise, mom)
    n = 700 if R < 243 else 150
    a = run_hyperband_repetitions(spec, n, seed=5, rep_offset=11, details=True)
    b = run_optimiser_repetitions(spec, n, tpe_config(False), seed=5, rep_offset=11, details=True)
    for k in ('ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors', 'ckpt_loss'):
        assert torch.equal(a[k], b[k]), k


def test_hybrid_structure_and_invariants():
    from autotune.b200.engine import FUNC_DOMAINS, run_hybrid_repetitions, run_hyperband_repetitions
    spec = _spec('rastrigin', FAM6)
    hb = run_hyperband_repetitions(spec, 2000, seed=9, details=True)
    for transfer in ('none', 'all', 'longest', 'same', 'threshold'):
        out = run_hybrid_repetitions(spec, 2000, transfer, threshold=9, seed=9, details=True)
        xy = out['arm_xy']
        x0, x1, y0, y1 = FUNC_DOMAINS['rastrigin']
        assert ((xy[..., 0] >= x0) & (xy[..., 0] < x1) & (xy[..., 1] >= y0) & (xy[..., 1] < y1)).all()
        assert torch.isfinite(out['ckpt_loss']).all()
        # the first 10 arms of the first bracket are hyperopt's random start-up trials: the kernel draws them from the
        # same streams as the Hyperband kernel, so their trajectories coincide with plain Hyperband's
        assert torch.equal(out['ckpt_loss'][:, :10], hb['ckpt_loss'][:, :10])
        # the 9-arm bracket never leaves start-up (9 < 10) unless history was transferred into it
        if transfer == 'none':
            assert torch.equal(out['ckpt_loss'][:, 108:], hb['ckpt_loss'][:, 108:])
        # suggested arms differ from the random ones
        assert not torch.equal(out['ckpt_loss'][:, 10:81], hb['ckpt_loss'][:, 10:81])
        assert torch.equal(out['ofe'], out['round_best'].min(dim=1).values)
        # determinism + sharding invariance
        again = run_hybrid_repetitions(spec, 500, transfer, threshold=9, seed=9, rep_offset=300)
        assert torch.equal(again['ofe'], out['ofe'][300:800])


# every hybrid OFE vector the reference stores (epdf_ofes/simulation-*/...sim(hb+tpe+transfer+*)...): five problem
# directories x four transfer modes
_SIM_DIRS = [('simulation-rastrigin-families-2/hist-sim-rastrigin-', 'rastrigin', FAM6, 10),
             ('simulation-rastrigin-families-1/hist-sim-rastrigin-', 'rastrigin', FAM6[:3], 10),
             ('simulation-flat-rastrigin/hist-sim-rastrigin-', 'rastrigin', FLAT, 0),
             ('simulation-flat-branin/hist-', 'branin', FLAT, 0),
             ('simulation-flat-drop-wave/hist-sim-wave-', 'wave', FLAT, 0)]
HYBRID_PICKLES = [(f'{pre}sim(hb+tpe+transfer+{mode})-7000-1', func, fams, noise, mode)
                  for pre, func, fams, noise in _SIM_DIRS for mode in ('none', 'all', 'same', 'longest')]
# Stand-alone TPE vectors of the flat problems (a flat trajectory is constant, so n_resources does not matter and the OFE
# is the best f(x, y) among the evaluations).  The pickles do NOT record max_iter; scanning it (scripts/tpe_budget_scan.py,
# device and oracle alike) leaves 13 evaluations for sim(tpe) and 25 for sim(tpe2xbudget) as the closest counts on all three
# problems (12 / 14 and 24 / 26 are each rejected on at least two of them) - i.e. ten random start-up trials plus 3 / 15
# suggestions.  An inferred configuration: a consistency check of the early TPE phase, not a pin.
TPE_PICKLES = [(f'{pre}sim({m})-7000-1', func, n_eval) for pre, func, fams, _ in _SIM_DIRS if fams is FLAT
               for m, n_eval in (('tpe', 13), ('tpe2xbudget', 25))]


@pytest.mark.parametrize('key,func,fams,noise,transfer', HYBRID_PICKLES)
def test_hybrid_ofe_distribution_matches_reference_pickles(key, func, fams, noise, transfer):
    """The only pins on TPE behaviour the reference offers: its stored 7000-repetition OFE vectors of the hybrids
    (two-sample KS at a family-wise 0.01 level over the twenty vectors and D < 0.03, fixed seed).  Plain Hyperband is distinguishable from these vectors (the reference's own
    HB-vs-hybrid KS is D = 0.07, p = 2e-14), so passing is not vacuous - see test_hybrid_differs_from_hyperband."""
    from scipy.stats import ks_2samp

    from autotune.b200.engine import run_hybrid_repetitions
    want = np.load(f'{GOLDEN}/epdf_ofes.npz')[key]
    got = run_hybrid_repetitions(_spec(func, fams, 81, noise), 20000, transfer, seed=21)['ofe'].cpu().numpy().astype(np.float64)
    res = ks_2samp(got, want)
    # twenty vectors are tested: 0.01 family-wise (Bonferroni) is 5e-4 per vector; measured: 19 of 20 above 0.01, one at 0.003
    assert res.pvalue > 0.01 / len(HYBRID_PICKLES), (res, got.mean(), want.mean(), got.std(), want.std())
    assert res.statistic < 0.03


@pytest.mark.parametrize('key,func,n_eval', TPE_PICKLES)
def test_standalone_tpe_ofe_distribution_close_to_reference_pickles(key, func, n_eval):
    """Stand-alone TpeOptimiser against the reference's stored sim(tpe) / sim(tpe2xbudget) OFE vectors of the flat problems at
    the inferred evaluation counts (see TPE_PICKLES): the two CDFs stay within 0.05 of each other (one evaluation more or
    fewer moves them apart by 0.03-0.05, so this bounds the count and the behaviour of the first suggestions together)."""
    from scipy.stats import ks_2samp

    from autotune.b200.engine import run_tpe_repetitions
    want = np.load(f'{GOLDEN}/epdf_ofes.npz')[key]
    got = run_tpe_repetitions(_spec(func, FLAT, 81, 0), 20000, 24, n_eval, seed=23)['ofe'].cpu().numpy().astype(np.float64)
    res = ks_2samp(got, want)
    assert res.statistic < 0.05, (res, got.mean(), want.mean())
    assert abs(got.mean() - want.mean()) < 0.04 * abs(want.mean()), (got.mean(), want.mean())


def test_hybrid_differs_from_hyperband():
    from scipy.stats import ks_2samp

    from autotune.b200.engine import run_hybrid_repetitions, run_hyperband_repetitions
    spec = _spec('rastrigin', FAM6)
    hb = run_hyperband_repetitions(spec, 20000, seed=1)['ofe'].cpu().numpy()
    hy = run_hybrid_repetitions(spec, 20000, 'all', seed=1)['ofe'].cpu().numpy()
    res = ks_2samp(hb, hy)
    assert res.statistic > 0.04 and hy.mean() < hb.mean()          # TPE finds better arms, as in the reference's pickles


def test_standalone_tpe_and_random_search_vs_oracle():
    import random

    from scipy.stats import ks_2samp

    from autotune.b200.engine import run_random_repetitions, run_tpe_repetitions
    from oracle import hybrid as ohy
    spec = _spec('rastrigin', FAM6)
    np.random.seed(3)
    random.seed(3)
    rs = np.random.RandomState(3)
    want_tpe = np.asarray([ohy.run_tpe('rastrigin', FAM6, 81, 10, 24, 81, rstate=rs) for _ in range(250)])
    want_rnd = np.asarray([ohy.run_random('rastrigin', FAM6, 81, 10, 24, 81) for _ in range(400)])
    got_tpe = run_tpe_repetitions(spec, 20000, 24, 81, seed=2)['ofe'].cpu().numpy().astype(np.float64)
    got_rnd = run_random_repetitions(spec, 20000, 24, 81, seed=2)['ofe'].cpu().numpy().astype(np.float64)
    assert ks_2samp(got_tpe, want_tpe).pvalue > 0.01
    assert ks_2samp(got_rnd, want_rnd).pvalue > 0.01
    assert got_tpe.mean() < got_rnd.mean()
    with pytest.raises(IndexError):
        run_tpe_repetitions(spec, 10, 82, 81)                      # fs[81] on an 81-point curve: python raises IndexError


@pytest.mark.parametrize('func,fams,R,noise', [('rastrigin', FAM6, 81, 10), ('egg', FAM6[:3], 243, 10), ('branin', FLAT, 81, 0),
                                               ('wave', FAM6, 27, 5)])
def test_tpe_observed_losses_equal_the_trajectory_readouts(func, fams, R, noise):
    """The TPE phase observes an arm's loss at the bracket's first resource level as f(x, y) + offset, the offset coming
    from an affine walk of the arm's Gamma stream made before the arm's hyperparameters were suggested
    (sim_core.cuh philox_affine_offset).  It must be the read-out the full trajectory - simulated afterwards with the
    suggested hyperparameters, the one successive halving ranks - shows at that checkpoint (FP32 rounding apart)."""
    from autotune.b200.engine import run_hybrid_repetitions
    spec = _spec(func, fams, R, noise)
    out = run_hybrid_repetitions(spec, 300, 'all', seed=31, details=True)
    obs, ck = out['obs_loss'], out['ckpt_loss']
    plan, n_checked = spec.plan, 0
    for b in range(plan.n_brackets):
        first, n_b = plan.bracket_first_arm[b], plan.bracket_n[b]
        c0 = plan.round_ckpt[plan.bracket_first_round[b]]
        o = obs[:, first:first + n_b]
        seen = ~torch.isnan(o)
        if n_b > 10:                                         # a TPE bracket: every arm but the last was observed
            assert seen[:, :-1].all() and not seen[:, -1].any()
        else:                                                # 9 arms < 10 start-up trials: no TPE phase, nothing observed
            assert not seen.any()
        want = ck[:, first:first + n_b, c0]
        scale = want[seen].abs().max().item() if seen.any() else 1.0
        assert torch.allclose(o[seen], want[seen], rtol=2e-6, atol=2e-6 * scale + 1e-5)
        n_checked += int(seen.sum())
    assert n_checked >= 300 * 26


@pytest.mark.parametrize('case', ['interior', 'edges', 'few'])
def test_truncated_mixture_sampler_matches_hyperopts_rejection_sampler(case):
    """A suggestion's candidates: tpe.GMM1 draws (component, normal) pairs until the value is inside [low, high); the device
    samples the same distribution directly (component with probability ~ weight x mass inside the bounds, then the inverse
    CDF of the truncated normal).  Two-sample KS of 32 x 2000 device draws against the oracle's rejection sampler, on good
    models whose components sit in the interior, on the bounds, and on a single observation."""
    from scipy.stats import ks_2samp
    lo, hi = -5.12, 5.12
    n_obs = {'interior': 40, 'edges': 40, 'few': 10}[case]
    rng = np.random.default_rng(7)
    obs1 = rng.uniform(lo, hi, n_obs).astype(np.float32)
    loss1 = rng.normal(size=n_obs).astype(np.float32)
    order = np.argsort(loss1, kind='stable')
    if case == 'edges':                                  # the good members hug the bounds: half of their mass is outside
        obs1[order[0]], obs1[order[1]] = np.float32(lo + 1e-3), np.float32(hi - 1e-3)
    n_prob = 2000
    obs = np.tile(obs1, (n_prob, 1))
    loss = np.tile(loss1, (n_prob, 1))
    cand = np.zeros((n_prob, 8), np.float32)
    samples = torch.ops.at2.tpe_score(torch.from_numpy(obs).cuda(), torch.from_numpy(loss).cuda(), lo, hi, _tpe_bytes(),
                                      torch.from_numpy(cand).cuda())[3].cpu().numpy().astype(np.float64).ravel()
    assert samples.min() >= lo and samples.max() < hi and len(np.unique(samples)) > 0.99 * samples.size
    below, _ = otpe.split_trials(obs1.astype(np.float64), loss1.astype(np.float64))
    w, mu, sg = otpe.adaptive_parzen_normal(below, 1.0, 0.5 * (hi + lo), hi - lo)
    want = otpe.gmm1_sample(w, mu, sg, lo, hi, 20000, np.random.RandomState(11))
    res = ks_2samp(samples, want)
    assert res.pvalue > 0.01, (case, res)


@pytest.mark.parametrize('R,eta,n_reps', [(81, 3, 200), (729, 9, 12)])
def test_transfer_bookkeeping_follows_the_reference_predicates(R, eta, n_reps):
    """How many earlier evaluations each bracket's TPE run receives (details output n_injected) equals the count the
    reference's predicates give (hybrid_hyperband_tpe_with_transfer.py:96-98,128-130,161-209) on the resource count it
    records per evaluator: int(r_i) of the arm's latest round - which is 0, not R, in the round whose read wraps to the last
    point (R = 729 with eta = 3 or 9: r = 0.9999999999999999), so those arms are NOT transferable in any mode.  (R = 729, eta = 9:
    two brackets of 729 + 81 arms, the largest such plan whose tables fit the TPE kernel's shared memory.)"""
    from oracle import hyperband as ohb
    from oracle.hybrid import is_transferable

    from autotune.b200.engine import run_hybrid_repetitions
    from autotune.b200.engine import HyperbandSimSpec
    spec = HyperbandSimSpec('rastrigin', FAM6[:2], R, eta, 10, 'uniform', 'min')
    plan = ohb.bracket_plan(R, eta)
    if R == 729:
        assert plan[0]['rounds'][0]['time'] == 0
    for mode in ('all', 'longest', 'same', 'threshold'):
        out = run_hybrid_repetitions(spec, n_reps, mode, threshold=9, seed=4, details=True)
        surv, inj = out['survivors'].cpu().numpy(), out['n_injected'].cpu().numpy()
        assert inj.shape == (n_reps, len(plan))
        for rep in range(n_reps):
            latest, first, pos = {}, 0, 0
            for bi, b in enumerate(plan):
                r0 = b['rounds'][0]['resource']
                if bi > 0 and b['n'] > 10:      # a bracket of <= 10 arms never leaves hyperopt's start-up: nothing to inject
                    want = sum(is_transferable(mode, r, r0, R, 9) for r in latest.values())
                    assert inj[rep, bi] == want, (mode, rep, bi, inj[rep], want)
                alive = list(range(first, first + b['n']))
                for rd in b['rounds']:
                    for a in alive:
                        latest[a] = rd['time']
                    k = min(rd['keep'], len(alive))
                    alive = surv[rep, pos:pos + k].tolist()
                    pos += k
                first += b['n']


@pytest.mark.parametrize('n_obs', [12, 40, 117])
def test_suggestion_argmax_never_differs_from_float64_over_many_problems(n_obs):
    """A suggestion is the arg-max of the score over 24 candidates.  Over 2048 random suggestion problems per history length
    the device scores - the full form and the lean one used inside tpe_sim - stay within 2e-5 of the float64 restatement
    (measured 1e-6, scripts/tpe_score_stats.py) and pick the same candidate every time."""
    n_prob, n_cand, lo, hi = 2048, 24, -5.12, 5.12
    rng = np.random.default_rng(1000 + n_obs)
    obs = rng.uniform(lo, hi, (n_prob, n_obs)).astype(np.float32)
    loss = rng.normal(size=(n_prob, n_obs)).astype(np.float32)
    cand = rng.uniform(lo, hi, (n_prob, n_cand)).astype(np.float32)
    scores, _, lean, _ = torch.ops.at2.tpe_score(torch.from_numpy(obs).cuda(), torch.from_numpy(loss).cuda(), lo, hi,
                                                 _tpe_bytes(), torch.from_numpy(cand).cuda())
    scores, lean = scores.cpu().numpy().astype(np.float64), lean.cpu().numpy().astype(np.float64)
    flips = 0
    for k in range(n_prob):
        if len(np.unique(obs[k])) < n_obs:
            # two equal observations: hyperopt orders the components with np.argsort(mus), which is not stable, so which of
            # the two gets which age weight is implementation-defined there (seen once in 6144 problems, score off by 1e-3)
            continue
        below, above = otpe.split_trials(obs[k].astype(np.float64), loss[k].astype(np.float64))
        wb, mb, sb = otpe.adaptive_parzen_normal(below, 1.0, 0.5 * (hi + lo), hi - lo)
        wa, ma, sa = otpe.adaptive_parzen_normal(above, 1.0, 0.5 * (hi + lo), hi - lo)
        want = otpe.gmm1_lpdf(cand[k].astype(np.float64), wb, mb, sb, lo, hi) - \
            otpe.gmm1_lpdf(cand[k].astype(np.float64), wa, ma, sa, lo, hi)
        assert np.abs(scores[k] - want).max() < 2e-5
        d = lean[k] - want
        assert np.abs(d - d.mean()).max() < 2e-5
        flips += int(np.argmax(lean[k]) != np.argmax(want)) + int(np.argmax(scores[k]) != np.argmax(want))
    assert flips == 0
