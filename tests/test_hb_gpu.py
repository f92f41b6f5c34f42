"""Parity tests proper for kernels (a)+(b): the CUDA path (through the torch ops -> C ABI) against reference-generated
golden fixtures, the oracle, and size-independent properties at full size.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, hb_case_names, load_hb_case

import autotune.b200  # noqa: F401,E402  (registers torch.ops.at2.*)

pytestmark = pytest.mark.gpu

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]


def _spec(cfg):
    from autotune.b200.engine import HyperbandSimSpec
    return HyperbandSimSpec(cfg['func'], cfg['families'], cfg['R'], cfg['eta'], cfg['noise'], 'uniform',
                            cfg['min_or_max'])


def _host_f1_fn(cfg, rep):
    """f_1 / f_n per arm in float64 with the product's own host functions (no oracle involved)."""
    from autotune.b200.engine import host_start_and_target
    return host_start_and_target(_spec(cfg), rep['xy'], rep['fam'], rep['z'])


def _run_predrawn(cfg, reps, dtype):
    from autotune.b200.engine import run_hyperband_predrawn
    f1fn = [_host_f1_fn(cfg, r) for r in reps]
    out = run_hyperband_predrawn(_spec(cfg), np.stack([a for a, _ in f1fn]), np.stack([b for _, b in f1fn]),
                                 np.stack([r['fam'] for r in reps]), np.stack([r['gamma'] for r in reps]),
                                 np.stack([r['xy'] for r in reps]), dtype=dtype)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize('name', hb_case_names())
def test_predrawn_f64_bit_exact_vs_reference(name):
    cfg, reps = load_hb_case(name)
    out = _run_predrawn(cfg, reps, torch.float64)
    smooth_fam = np.asarray([bool(f[3]) for f in cfg['families']])
    for i, rep in enumerate(reps):
        is_s = smooth_fam[rep['fam']]
        # trajectories read at the checkpoints: bit-exact for raw families, ~1e-13 through the Savitzky-Golay rows
        got, want = out['ckpt_loss'][i], rep['ckpt_loss']
        # golden ckpts may hold python's time-0 quirk; the plan maps it to position R
        assert np.array_equal(got[~is_s], want[~is_s])
        if is_s.any():
            assert np.max(np.abs(got[is_s] - want[is_s]) / np.maximum(1, np.abs(want[is_s]))) < 1e-10
        # index work: bit-exact
        assert np.array_equal(out['survivors'][i], rep['surv_ids'])
        assert np.array_equal(out['round_best_arm'][i], rep['round_best_id'])
        assert out['ofe_arm'][i] == rep['ofe_id']
        np.testing.assert_allclose(out['round_best'][i], rep['round_best'], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out['ofe'][i], rep['ofe'], rtol=1e-10, atol=1e-10)
        raw_best = ~is_s[rep['round_best_id']]
        assert np.array_equal(out['round_best'][i][raw_best], rep['round_best'][raw_best])
        assert np.array_equal(out['best_xy'][i], rep['xy'][rep['ofe_id']])


@pytest.mark.parametrize('name', ['rastrigin6_R81', 'egg3_R243', 'wave_flat_R81', 'rastrigin6_R27_max'])
def test_predrawn_f32_within_tolerance(name):
    """Production arithmetic (FP32, FMA-contracted, approximate reciprocal) on the same inputs: trajectories within
    1e-4 relative of the float64 reference (SURVEY 7 hard part 1 measured 1.6e-5 drift for plain FP32)."""
    cfg, reps = load_hb_case(name)
    out = _run_predrawn(cfg, reps, torch.float32)
    for i, rep in enumerate(reps):
        want = rep['ckpt_loss']
        err = np.abs(out['ckpt_loss'][i] - want) / np.maximum(1, np.abs(want))
        assert err.max() < 1e-4, err.max()
        assert abs(out['ofe'][i] - rep['ofe']) <= 1e-4 * max(1, abs(rep['ofe']))
        same = np.mean(out['survivors'][i] == rep['surv_ids'])
        assert same > 0.9          # near-ties may legitimately flip in FP32


def test_constant_loss_first_arm_wins_on_device():
    cfg, reps = load_hb_case('branin_flat_R81')
    out = _run_predrawn(cfg, reps[1:2], torch.float64)
    assert out['ofe_arm'][0] == 0
    assert out['survivors'][0][:27].tolist() == list(range(27))
    assert out['round_best_arm'][0].tolist() == reps[1]['round_best_id'].tolist()


def _torch_halving(spec, ckpt_loss, descending=False):
    """Independent restatement of successive halving with torch stable sorts on the device (size-independent check
    of kernel (b)): returns survivors [n, n_surv], round_best [n, rounds], ofe [n]."""
    n = ckpt_loss.shape[0]
    surv, rbest = [], []
    p = spec.plan
    for b in range(p.n_brackets):
        ids = torch.arange(p.bracket_first_arm[b], p.bracket_first_arm[b] + p.bracket_n[b],
                           device=ckpt_loss.device).expand(n, -1)
        for r in range(p.bracket_first_round[b], p.bracket_first_round[b + 1]):
            loss = torch.gather(ckpt_loss[:, :, p.round_ckpt[r]], 1, ids)
            order = torch.sort(loss, dim=1, stable=True, descending=descending).indices
            keep = min(p.round_keep[r], ids.shape[1])
            rbest.append(loss.max(dim=1).values if descending else loss.min(dim=1).values)
            ids = torch.gather(ids, 1, order[:, :keep])
            surv.append(ids)
    rb = torch.stack(rbest, 1)
    return torch.cat(surv, 1).int(), rb, (rb.max(1).values if descending else rb.min(1).values)


@pytest.mark.parametrize('func,families,R,noise,mom', [('rastrigin', FAM6, 81, 10, 'min'), ('egg', FAM6[:3], 243, 10, 'min'),
                                                      ('wave', FAM6, 27, 10, 'max'), ('branin', FLAT, 81, 0, 'min')])
def test_philox_halving_consistent_at_scale(func, families, R, noise, mom):
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec(func, families, R, 3, noise, 'uniform', mom)
    n = 3000 if R < 243 else 600
    out = run_hyperband_repetitions(spec, n, seed=7, details=True)
    surv, rb, ofe = _torch_halving(spec, out['ckpt_loss'], mom == 'max')
    assert torch.equal(surv, out['survivors'])
    assert torch.equal(rb, out['round_best'])
    assert torch.equal(ofe, out['ofe'])
    assert torch.isfinite(out['ckpt_loss']).all()
    # the winning arm really carries the OFE, and best_xy lies in the domain
    from autotune.b200.engine import FUNC_DOMAINS
    x0, x1, y0, y1 = FUNC_DOMAINS[func]
    xy = out['best_xy']
    assert ((xy[:, 0] >= x0) & (xy[:, 0] <= x1) & (xy[:, 1] >= y0) & (xy[:, 1] <= y1)).all()
    # last checkpoint of a raw arm is its target f_n: never above its start by construction of the flat family
    if families is FLAT:
        assert torch.equal(out['ckpt_loss'][:, :, 0], out['ckpt_loss'][:, :, -1])


def test_philox_sharding_and_seed_invariance():
    """Counter-based RNG keyed by the global repetition id: any split of the repetition range over calls (= over GPUs)
    gives bit-identical OFEs; a different seed gives a different sample."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
    whole = run_hyperband_repetitions(spec, 5000, seed=3)['ofe']
    parts = torch.cat([run_hyperband_repetitions(spec, 1234, seed=3, rep_offset=0)['ofe'],
                       run_hyperband_repetitions(spec, 3000, seed=3, rep_offset=1234)['ofe'],
                       run_hyperband_repetitions(spec, 766, seed=3, rep_offset=4234)['ofe']])
    assert torch.equal(whole, parts)
    assert torch.equal(whole, run_hyperband_repetitions(spec, 5000, seed=3)['ofe'])
    assert not torch.equal(whole, run_hyperband_repetitions(spec, 5000, seed=4)['ofe'])


EPDF_CASES = [
    ('simulation-rastrigin-families-2/hist-sim-rastrigin-sim(hb)-7000-1', 'rastrigin', FAM6, 10),
    ('simulation-rastrigin-families-1/hist-sim-rastrigin-sim(hb)-7000-1', 'rastrigin', FAM6[:3], 10),
    ('simulation-flat-rastrigin/hist-sim-rastrigin-sim(hb)-7000-1', 'rastrigin', FLAT, 0),
    ('simulation-flat-branin/hist-sim(hb)-7000-1', 'branin', FLAT, 0),
    ('simulation-flat-drop-wave/hist-sim-wave-sim(hb)-7000-1', 'wave', FLAT, 0),
]


@pytest.mark.parametrize('key,func,families,noise', EPDF_CASES)
def test_philox_ofe_distribution_matches_reference_pickles(key, func, families, noise):
    """Distributional known-answer test: OFEs from the Philox path vs the reference's stored 7000-repetition EPDF-OFE
    vectors (configs recovered in SURVEY 4), two-sample KS p > 0.01 - the statistic the reference itself uses
    (plot_run_simulations_results.py:69-73).  Fixed seed, so the verdict is reproducible."""
    from scipy.stats import ks_2samp

    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    want = np.load(f'{GOLDEN}/epdf_ofes.npz')[key]
    spec = HyperbandSimSpec(func, families, 81, 3, noise)
    got = run_hyperband_repetitions(spec, 20000, seed=11)['ofe'].cpu().numpy().astype(np.float64)
    res = ks_2samp(got, want)
    assert res.pvalue > 0.01, (res, got.mean(), want.mean(), got.std(), want.std())


def test_philox_vs_oracle_global_rng_distribution():
    """Same, against fresh runs of the (reference-pinned) oracle with numpy's own RNG - egg-holder R=27 is not among the
    stored pickles."""
    import random

    from scipy.stats import ks_2samp

    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    from oracle import hyperband as ohb
    fams = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]
    np.random.seed(5)
    random.seed(5)
    want = np.asarray([ohb.run_global_rng('egg', fams, 27, 3, 10) for _ in range(1200)])
    spec = HyperbandSimSpec('egg', fams, 27, 3, 10)
    got = run_hyperband_repetitions(spec, 20000, seed=5)['ofe'].cpu().numpy().astype(np.float64)
    res = ks_2samp(got, want)
    assert res.pvalue > 0.01, (res, got.mean(), want.mean())


def test_full_size_job_properties():
    """BASELINE config 2 at full size (10^5 repetitions, flat Branin): the flat family makes every trajectory constant,
    so the OFE must equal the minimum Branin value over the repetition's 117 arms; and the OFE can never be below the
    function's global minimum."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec('branin', FLAT, 81, 3, 0)
    out = run_hyperband_repetitions(spec, 100000, seed=1)
    ofe = out['ofe']
    assert ofe.shape == (100000,) and torch.isfinite(ofe).all()
    assert ofe.min().item() >= 0.397887 - 1e-4
    xy = out['best_xy'].double()
    x, y = xy[:, 0], xy[:, 1]
    b, c, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 1 / (8 * np.pi)
    f = (y - b * x ** 2 + c * x - 6) ** 2 + 10 * (1 - t) * torch.cos(x) + 10
    assert torch.allclose(f, ofe.double(), rtol=1e-4, atol=1e-4)


def test_device_errors_are_loud():
    from autotune.b200.engine import HyperbandSimSpec
    with pytest.raises(ValueError, match='window_length'):
        HyperbandSimSpec('branin', FAM6, 6, 3, 10)          # the reference fails the same way inside scipy
    with pytest.raises(ValueError):
        HyperbandSimSpec('branin', FLAT, 81, 3, 0, min_or_max='median')


@pytest.mark.parametrize('func,families,R,eta,noise,mom', [('rastrigin', FAM6, 81, 3, 10, 'min'), ('branin', FLAT, 81, 3, 0, 'min'),
                                                          ('egg', FAM6[:3], 243, 3, 10, 'min'), ('wave', FAM6, 27, 3, 10, 'max'),
                                                          ('camel', FAM6, 16, 2, 1, 'min'), ('rastrigin', FAM6, 100, 3, 10, 'min')])
def test_early_stopping_is_bit_identical_to_full_simulation(func, families, R, eta, noise, mom):
    """Early stopping only skips epochs nobody reads: with per-arm counter-based streams every output (OFE, winning arm,
    round-bests, survivor sets) equals the full simulation's bit for bit, for any repetition range."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec(func, families, R, eta, noise, 'uniform', mom)
    n = 1500 if R < 243 else 300
    full = run_hyperband_repetitions(spec, n, seed=13, rep_offset=7, details=True)
    lazy = run_hyperband_repetitions(spec, n, seed=13, rep_offset=7, details=True, early_stop=True)
    for k in ('ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors'):
        assert torch.equal(full[k], lazy[k]), k
    assert 'ckpt_loss' not in lazy
    part = run_hyperband_repetitions(spec, 37, seed=13, rep_offset=7 + 100, early_stop=True)
    assert torch.equal(part['ofe'], full['ofe'][100:137])


@pytest.mark.parametrize('func,families,R,eta', [('branin', FAM6[:2], 6, 3), ('camel', FAM6, 16, 2), ('rastrigin', FAM6, 729, 3),
                                                 ('wave', FAM6, 1121, 3), ('egg', FAM6, 100, 3), ('branin', FAM6[:2], 2, 2),
                                                 ('rastrigin', FAM6, 64, 4)])
def test_unusual_plans_run_and_agree_with_independent_halving(func, families, R, eta):
    """Bracket tables far from the R=81 default: tiny (R=2, 6), eta=2/4, non-powers (R=100), wide first rounds (729 arms
    at R=1121) and the R=729 quirk where the first round reads the last trajectory point (python's fs[-1])."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec(func, families, R, eta, 10, 'round-robin', 'min')
    n = 64 if R > 300 else 400
    out = run_hyperband_repetitions(spec, n, seed=3, details=True)
    surv, rb, ofe = _torch_halving(spec, out['ckpt_loss'])
    assert torch.equal(surv, out['survivors']) and torch.equal(rb, out['round_best']) and torch.equal(ofe, out['ofe'])
    lazy = run_hyperband_repetitions(spec, n, seed=3, details=True, early_stop=True)
    assert torch.equal(lazy['ofe'], out['ofe']) and torch.equal(lazy['survivors'], out['survivors'])
    if R == 729:
        assert spec.rounds()[0]['time'] == 729     # int(0.9999999999999999) == 0 -> python index -1 -> the last point


@pytest.mark.parametrize('tile', [1, 3, 7, 16, 32])
def test_early_stopping_tile_widths_do_not_change_outputs(tile):
    """The early-stopping kernel cuts a job into tiles of G repetitions per warp (chosen from the job size; a small job
    gets narrow tiles), works through a round in chunks of repetitions, lists raw arms ahead of smoothed ones and ranks
    narrow rounds 32 / n repetitions at a time.  None of that may change a bit: every width (forced here) must reproduce
    the full simulation, ragged last tile and last chunk included."""
    from autotune.b200._native import debug_options
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    with debug_options(lazy_reps_per_tile=tile):
        for func, fams, R, eta, noise, mom, n in (('rastrigin', FAM6, 81, 3, 10, 'min', 1000 + 13), ('branin', FLAT, 81, 3, 0, 'min', 997),
                                                  ('egg', FAM6[:3], 243, 3, 10, 'min', 101), ('wave', FAM6[2:3], 27, 3, 10, 'max', 500),
                                                  ('camel', FAM6, 16, 2, 1, 'min', 333)):
            spec = HyperbandSimSpec(func, fams, R, eta, noise, 'uniform', mom)
            full = run_hyperband_repetitions(spec, n, seed=29, rep_offset=3, details=True, early_stop=False)
            lazy = run_hyperband_repetitions(spec, n, seed=29, rep_offset=3, details=True, early_stop=True)
            for k in ('ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors'):
                assert torch.equal(full[k], lazy[k]), (func, k)
    with debug_options(lazy_reps_per_tile=tile, hb_no_split=1):          # mixed passes as before the raw-first list
        spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
        a = run_hyperband_repetitions(spec, 700, seed=5, details=True, early_stop=False)
        b = run_hyperband_repetitions(spec, 700, seed=5, details=True, early_stop=True)
        for k in ('ofe', 'ofe_arm', 'round_best', 'survivors'):
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize('tile', [0, 5, 32])
def test_early_stopping_setup_cache_does_not_change_outputs(tile):
    """The early-stopping kernel draws an arm (hyperparameters, family, initial noise, base function) once per bracket and
    reads its start value, target and family back from a per-warp slab of scratch in the later rounds; without the slab
    (debug option lazy_no_cache, or a refused allocation) it redraws them every round.  Same streams, same bits - with the
    raw-first lists (mixed families), without them, and on the 4-D extension whose draw has two more coordinates."""
    from autotune.b200._native import debug_options
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    for func, fams, R, eta, noise, sched, n in (('rastrigin', FAM6, 81, 3, 10, 'uniform', 1500), ('branin', FLAT, 81, 3, 0, 'uniform', 700),
                                                ('egg', FAM6[:3], 243, 3, 10, 'round-robin', 90), ('egg4', FAM6[:2], 27, 3, 5, 'uniform', 400),
                                                ('rastrigin', FAM6, 729, 3, 10, 'uniform', 6)):
        spec = HyperbandSimSpec(func, fams, R, eta, noise, sched, 'min')
        with debug_options(lazy_reps_per_tile=tile):
            cached = run_hyperband_repetitions(spec, n, seed=31, rep_offset=7, details=True, early_stop=True)
        with debug_options(lazy_reps_per_tile=tile, lazy_no_cache=1):
            redrawn = run_hyperband_repetitions(spec, n, seed=31, rep_offset=7, details=True, early_stop=True)
        for k in ('ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors'):
            assert torch.equal(cached[k], redrawn[k]), (func, R, k)
    full = run_hyperband_repetitions(spec, n, seed=31, rep_offset=7, details=True, early_stop=False)
    assert torch.equal(full['ofe'], cached['ofe']) and torch.equal(full['survivors'], cached['survivors'])


def test_egg4_extension_flat_family_ofe_is_the_function_value():
    """4-D Egg-holder of BASELINE.json config 5 - an EXTENSION with no reference equivalent, defined in
    include/at2.h and oracle/opt_functions.egg4.  With the flat family (no noise, no shifts) every loss of an arm equals
    its function value, so the OFE must be the oracle's float64 egg4 at the reported hyperparameters; the 2-D 'egg' is
    untouched; early stopping and sharding keep their bit-identity."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    from oracle import opt_functions as of
    spec = HyperbandSimSpec('egg4', FLAT, 81, 3, 0)
    out = run_hyperband_repetitions(spec, 4000, seed=17, details=True)
    xy = out['best_xy'].double().cpu().numpy()
    assert xy.shape == (4000, 4) and (np.abs(xy) <= 512).all()
    want = of.egg4(xy[:, 0], xy[:, 1], xy[:, 2], xy[:, 3])
    np.testing.assert_allclose(out['ofe'].double().cpu().numpy(), want, rtol=2e-4, atol=0.25)   # FP32 sin(sqrt(.)) of |arg| < 1100
    assert len(np.unique(xy[:, 2])) > 3900 and abs(np.corrcoef(xy[:, 0], xy[:, 2])[0, 1]) < 0.1   # own coordinates
    # lower than the best of 117 random 2-D egg arms on average: three pair terms instead of one
    egg2 = run_hyperband_repetitions(HyperbandSimSpec('egg', FLAT, 81, 3, 0), 4000, seed=17)
    assert egg2['best_xy'].shape == (4000, 2) and out['ofe'].mean() < egg2['ofe'].mean()
    lazy = run_hyperband_repetitions(spec, 4000, seed=17, details=True, early_stop=True)
    for k in ('ofe', 'ofe_arm', 'best_xy', 'round_best', 'survivors'):
        assert torch.equal(out[k], lazy[k]), k
    part = run_hyperband_repetitions(spec, 100, seed=17, rep_offset=1234)
    assert torch.equal(part['ofe'], out['ofe'][1234:1334]) and torch.equal(part['best_xy'], out['best_xy'][1234:1334])
    from autotune.b200.engine import run_hybrid_repetitions
    with pytest.raises(Exception, match='2-D'):
        run_hybrid_repetitions(spec, 10, 'none')


def test_global_memory_tables_are_bit_identical_to_shared_memory_tables():
    """Plans too large for shared memory read their constant tables from global memory (GTAB kernel variants); forced
    here on ordinary plans, every output must equal the shared-memory variant's bit for bit - full and early-stopping."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    for func, fams, R in (('rastrigin', FAM6, 81), ('branin', FLAT, 81), ('egg', FAM6[:3], 243)):
        spec = HyperbandSimSpec(func, fams, R, 3, 10)
        n = 600 if R < 243 else 150
        from autotune.b200._native import debug_options
        a = run_hyperband_repetitions(spec, n, seed=23, details=True, early_stop=False)
        al = run_hyperband_repetitions(spec, n, seed=23, details=True, early_stop=True)
        with debug_options(force_gtab=1):
            b = run_hyperband_repetitions(spec, n, seed=23, details=True, early_stop=False)
            bl = run_hyperband_repetitions(spec, n, seed=23, details=True, early_stop=True)
        for k in a:
            assert torch.equal(a[k], b[k]), (func, k)
        for k in al:
            assert torch.equal(al[k], bl[k]) and torch.equal(al[k], a[k]), (func, k)


def test_tile_halving_is_bit_identical_to_per_repetition_halving():
    """Float path, warp-private tiles: the repetitions of a tile are halved together, narrow rounds side by side
    (halving.cuh tile_halving_f32).  Every output must equal the per-repetition path's, ragged last tiles, maximisation,
    eta = 2 / 4 plans and heavily tied losses included."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    for func, fams, R, eta, noise, mom, n in (('rastrigin', FAM6, 81, 3, 10, 'min', 2000 + 1), ('branin', FLAT, 81, 3, 0, 'min', 1999),
                                              ('egg', FAM6[:3], 243, 3, 10, 'min', 301), ('wave', FAM6, 27, 3, 10, 'max', 1000),
                                              ('camel', FAM6, 16, 2, 1, 'min', 777), ('rastrigin', FAM6, 64, 4, 10, 'min', 500),
                                              ('rastrigin', FAM6, 100, 3, 10, 'min', 400), ('branin', FAM6[:2], 2, 2, 10, 'min', 100)):
        spec = HyperbandSimSpec(func, fams, R, eta, noise, 'uniform', mom)
        from autotune.b200._native import debug_options
        a = run_hyperband_repetitions(spec, n, seed=31, rep_offset=5, details=True, early_stop=False)
        with debug_options(hb_no_tile_halving=1):
            b = run_hyperband_repetitions(spec, n, seed=31, rep_offset=5, details=True, early_stop=False)
        for k in a:
            assert torch.equal(a[k], b[k]), (func, R, k)


def test_raw_smoothed_split_passes_do_not_change_outputs():
    """Mixed families: raw and smoothed arms of a tile run in separate warp passes (hb_sim.cu).  Disabling the split must
    give the same bits."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    for func, fams, R, sched in (('rastrigin', FAM6, 81, 'uniform'), ('egg', FAM6[:3], 243, 'round-robin'), ('wave', FAM6, 27, 'uniform')):
        spec = HyperbandSimSpec(func, fams, R, 3, 10, sched)
        from autotune.b200._native import debug_options
        a = run_hyperband_repetitions(spec, 500, seed=29, rep_offset=3, details=True, early_stop=False)
        with debug_options(hb_no_split=1):
            b = run_hyperband_repetitions(spec, 500, seed=29, rep_offset=3, details=True, early_stop=False)
        for k in a:
            assert torch.equal(a[k], b[k]), (func, k)


@pytest.mark.parametrize('R,eta,mom', [(81, 3, 'min'), (81, 3, 'max'), (243, 3, 'min'), (27, 3, 'max'), (64, 2, 'min')])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_heavily_tied_losses_keep_pythons_stable_order(R, eta, mom, dtype):
    """Losses drawn from five values only: almost every comparison of a round is a tie, so the survivor lists are decided
    by python's stable sorted() / first-min rules (hyperband_optimiser.py:46-57,99-100) - in the register rounds (rank by
    counting + MATCH.ANY, 64-bit bitonic keys) of the float path and the shared-memory network of the float64 path alike."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_predrawn
    spec = HyperbandSimSpec('branin', FLAT, R, eta, 0, 'uniform', mom)
    n, A = 64, spec.n_arms
    g = torch.Generator().manual_seed(R * 7 + eta)
    level = torch.randint(0, 5, (n, A), generator=g).double() - 2.0          # includes -0.0 + 0.0 cases via 0
    level[0] = 1.0                                                            # one repetition with every loss equal
    level[1, ::2] = -0.0                                                      # signed zeros compare equal in python
    level[1, 1::2] = 0.0
    gamma = torch.full((n, A, R - 1), 2.5, dtype=torch.float64)
    out = run_hyperband_predrawn(spec, level, level.clone(), torch.zeros(n, A, dtype=torch.int32), gamma, dtype=dtype)
    ck = out['ckpt_loss']
    assert torch.equal(ck, level.to(dtype).cuda().unsqueeze(-1).expand_as(ck))  # flat family: the curve never moves
    surv, rb, ofe = _torch_halving(spec, ck, descending=(mom == 'max'))
    assert torch.equal(surv, out['survivors']) and torch.equal(rb, out['round_best']) and torch.equal(ofe, out['ofe'])
    first_round_keep = spec.plan.round_keep[0]
    assert out['survivors'][0][:first_round_keep].tolist() == list(range(first_round_keep))   # all equal: first arms win
    assert out['ofe_arm'][0] == 0 and out['round_best_arm'][0][0] == 0


def test_fused_gather_entry_point_on_one_gpu():
    """at2_hb_sim_gather_f32 with plain local buffers as gather targets (the multi-GPU run maps the peers' symmetric buffers
    or the NVSwitch multicast address there - scripts/check_fused_gather.py verifies that at 2 and 8 GPUs): every target
    receives the OFEs at gather_base + i, nothing else is touched, outputs equal at2_hb_sim_f32's."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
    dev = torch.device('cuda')
    n, base, total = 777, 1000, 2500
    want = run_hyperband_repetitions(spec, n, seed=8, rep_offset=base)
    a = torch.full((total,), -7.0, device=dev)
    b = torch.full((total,), -7.0, device=dev)
    outs = torch.ops.at2.hb_sim_gather(spec.tables(dev, torch.float32), spec.plan_bytes, spec.cfg_bytes, n, base, 8, a,
                                       [a.data_ptr(), b.data_ptr()], base)
    assert torch.equal(outs[0], want['ofe']) and torch.equal(outs[1], want['ofe_arm']) and torch.equal(outs[2], want['best_xy'])
    for buf in (a, b):
        assert torch.equal(buf[base:base + n], want['ofe'])
        assert (buf[:base] == -7.0).all() and (buf[base + n:] == -7.0).all()
    # the early-stopping kernel with the same fused gather (the product's default path at N > 1)
    c = torch.full((total,), -7.0, device=dev)
    outs = torch.ops.at2.hb_sim_gather(spec.tables(dev, torch.float32), spec.plan_bytes, spec.cfg_bytes, n, base, 8, c,
                                       [c.data_ptr()], base, False, True)
    assert torch.equal(outs[0], want['ofe']) and torch.equal(c[base:base + n], want['ofe'])
    assert (c[:base] == -7.0).all() and (c[base + n:] == -7.0).all()
    with pytest.raises(RuntimeError):
        torch.ops.at2.hb_sim_gather(spec.tables(dev, torch.float32), spec.plan_bytes, spec.cfg_bytes, n, base, 8, a,
                                    [a.data_ptr()], total - 5)            # would write past the buffer
    with pytest.raises(RuntimeError):
        torch.ops.at2.hb_sim_gather(spec.tables(dev, torch.float32), spec.plan_bytes, spec.cfg_bytes, n, base, 8, a, [], base)
