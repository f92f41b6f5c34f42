"""Parity tiers of `north_star`, measured directly on the device path (round-2 additions):

* the device Gamma process itself (Philox -> Box-Muller -> Marsaglia-Tsang -> recurrence -> read-out) against the oracle's
  restatement of core/simulation_evaluator.py:36-48,97-116 driven by numpy's own Gamma generator - two-sample KS on the
  loss an arm shows at the Hyperband checkpoints, per shape family, raw and smoothed, R = 81 and R = 243;
* the FP32 production arithmetic against the float64 reference order on identical pre-drawn increments: survivor-slot
  flip rate, winner flip rate and trajectory error over 10^4 repetitions, asserted against stated bounds;
* the Camel function on the device against its float64 definition;
* the evaluators handed back through the optimiser API carry the shape family their arm really drew on the device.
"""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import autotune.b200  # noqa: F401  (registers torch.ops.at2.*)

pytestmark = pytest.mark.gpu

RAW_A = (1.5, 10, 15, False, 0, 200)
RAW_B = (0.5, 7, 10, False, 200, 400)
SMOOTH = (0.2, 4, 7, True, 0, 200)


def _device_offsets(func, fam, R, n_reps, seed):
    """loss(arm, checkpoint) - f(x, y) for every arm of `n_reps` repetitions of a single-family, noise-free job.  The
    recurrence and the read-out are affine in (f_1, f_n) = (f - start_shift, f - end_shift), so this offset does not
    depend on the arm's hyperparameters: every arm is one sample of the same distribution."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions, run_optimiser_repetitions, tpe_config
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    spec = HyperbandSimSpec(func, [fam], R, 3, 0.0)
    full = run_hyperband_repetitions(spec, n_reps, seed=seed, details=True)
    # arm_xy comes from the TPE kernel run with TPE disabled - the same streams and arithmetic, bit for bit
    aux = run_optimiser_repetitions(spec, n_reps, tpe_config(False), seed=seed, details=True)
    assert torch.equal(full['ckpt_loss'], aux['ckpt_loss'])
    xy = aux['arm_xy'].double().cpu().numpy()
    f = np.vectorize(NOISE_FREE_FUNCTIONS[func])(xy[..., 0], xy[..., 1])
    off = full['ckpt_loss'].double().cpu().numpy() - f[..., None]
    return spec, off.reshape(-1, off.shape[-1])


def _oracle_offsets(fam, R, checkpoints, n, seed):
    """The same offsets from the oracle: simulate() from f_1 = -start_shift towards f_n = -end_shift (f = 0) with
    np.random.gamma(shape, scale) draws (simulation_evaluator.py:36-48), read out as `.fs` (:76-82)."""
    from oracle import gamma_process as gp
    ml, nec, up, smooth, s0, s1 = fam
    rng = np.random.RandomState(seed)
    shapes = [gp.gamma_shape_scale(t, R) for t in range(1, R)]
    out = np.empty((n, len(checkpoints)))
    for i in range(n):
        g = [rng.gamma(a, s) for a, s in shapes]
        raw = gp.simulate(-float(s0), -float(s1), R, ml, nec, up, lambda t: g[t - 1])
        fs = gp.readout(raw, smooth, R)
        out[i] = [fs[c - 1] for c in checkpoints]
    return out


@pytest.mark.parametrize('fam', [RAW_A, RAW_B, SMOOTH], ids=['raw-fast', 'raw-shifted', 'smoothed'])
@pytest.mark.parametrize('R', [81, 243])
def test_device_gamma_process_matches_reference_distribution(fam, R):
    from scipy.stats import ks_2samp
    n_reps = 200 if R == 81 else 60                      # x 117 / 364 arms: >= 2e4 trajectories
    spec, got = _device_offsets('rastrigin', fam, R, n_reps, seed=2024 + R)
    assert got.shape[0] >= 20000
    cks = spec.checkpoints
    want = _oracle_offsets(fam, R, cks, 20000 if R == 81 else 8000, seed=7 + R)
    for c, time in enumerate(cks):
        if time in (1, R) and not fam[3]:
            # a raw arm starts on f_1 and lands on f_n exactly (:113-115): deterministic, compare the values
            assert np.allclose(got[:, c], want[0, c], atol=2e-3), (time, got[:, c].min(), got[:, c].max(), want[0, c])
            continue
        res = ks_2samp(got[:, c], want[:, c])
        assert res.pvalue > 0.01, (fam, R, time, res, got[:, c].mean(), want[:, c].mean(), got[:, c].std(), want[:, c].std())
        # first two moments as a sanity check on top of the KS verdict
        se = want[:, c].std() / np.sqrt(len(want))
        assert abs(got[:, c].mean() - want[:, c].mean()) < 5 * se + 1e-3


# bounds asserted below; the measured values are reported in DESIGN.md section 4
FLIP_BOUND_SURVIVOR_SLOTS = 2e-3      # share of survivor slots that hold another arm id than the float64 reference order
FLIP_BOUND_WINNER = 1e-3              # share of repetitions whose OFE arm differs
REL_ERR_BOUND = 1e-4                  # max relative error of a checkpoint loss (|x| floored at 1)


def test_fp32_production_arithmetic_flip_rate_over_1e4_repetitions():
    """FP32 production arithmetic (folded step, approximate reciprocal) vs the reference's float64 operation order on the
    SAME pre-drawn inputs, 10^4 repetitions of the six-family Rastrigin job: how often a near-tie flips."""
    from autotune.b200.engine import HyperbandSimSpec, host_start_and_target, run_hyperband_predrawn
    fams = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
            (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
    R, noise = 81, 10.0
    spec = HyperbandSimSpec('rastrigin', fams, R, 3, noise)
    A = spec.n_arms
    t = np.arange(1, R)
    m = (R + 1) - t
    beta = (2 + np.sqrt(4 + 4 * m)) / (2 * m)
    rng = np.random.default_rng(123)
    slots = flips = winners = n_total = 0
    worst = 0.0
    ofe_worst = 0.0
    for chunk in range(4):
        n = 2500
        xy = rng.uniform(-5.12, 5.12, (n, A, 2))
        fam = rng.integers(0, len(fams), (n, A)).astype(np.int32)
        z = rng.standard_normal((n, A))
        gamma = rng.gamma(2 * beta + 1, 1 / beta, size=(n, A, R - 1))
        # f_1 / f_n vectorised here (the scalar host path of the product is exercised by the golden tests)
        f = 20 + (xy[..., 0] ** 2 - 10 * np.cos(2 * np.pi * xy[..., 0])) + (xy[..., 1] ** 2 - 10 * np.cos(2 * np.pi * xy[..., 1]))
        s0 = np.asarray([q[4] for q in fams], float)[fam]
        s1 = np.asarray([q[5] for q in fams], float)[fam]
        f1, fn = (f + noise * z) - s0, f - s1
        if chunk == 0:   # the vectorised f_1 / f_n agree with the product's host path on a sample
            a, b = host_start_and_target(spec, xy[:3], fam[:3], z[:3])
            assert np.allclose(a, f1[:3], rtol=1e-12, atol=1e-9) and np.allclose(b, fn[:3], rtol=1e-12, atol=1e-9)
        o64 = run_hyperband_predrawn(spec, f1, fn, fam, gamma, xy, dtype=torch.float64)
        o32 = run_hyperband_predrawn(spec, f1, fn, fam, gamma, xy, dtype=torch.float32)
        slots += o64['survivors'].numel()
        flips += int((o64['survivors'] != o32['survivors']).sum())
        winners += int((o64['ofe_arm'] != o32['ofe_arm']).sum())
        n_total += n
        err = (o32['ckpt_loss'].double() - o64['ckpt_loss']).abs() / o64['ckpt_loss'].abs().clamp(min=1)
        worst = max(worst, float(err.max()))
        ofe_worst = max(ofe_worst, float(((o32['ofe'].double() - o64['ofe']).abs() / o64['ofe'].abs().clamp(min=1)).max()))
        del o64, o32
    rate, wrate = flips / slots, winners / n_total
    print(f'FP32 vs float64 on identical increments, {n_total} repetitions: survivor-slot flip rate {rate:.3e} '
          f'({flips}/{slots}), winner flip rate {wrate:.3e} ({winners}/{n_total}), max rel. checkpoint error {worst:.3e}, '
          f'max rel. OFE error {ofe_worst:.3e}')
    assert n_total >= 10000
    assert rate <= FLIP_BOUND_SURVIVOR_SLOTS, rate
    assert wrate <= FLIP_BOUND_WINNER, wrate
    assert worst <= REL_ERR_BOUND, worst


def test_camel_on_device_matches_float64_definition():
    """benchmarks/opt_function_problem.py six-hump camel: on the flat family every trajectory is the constant f(x, y), so
    the OFE is the minimum Camel value over the repetition's arms - checked against the float64 definition at best_xy,
    against the known global minimum, and against an independent float64 evaluation of every arm."""
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions, run_optimiser_repetitions, tpe_config
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    spec = HyperbandSimSpec('camel', [(0, 1, 0, False, 0, 0)], 81, 3, 0)
    out = run_hyperband_repetitions(spec, 20000, seed=3)
    xy = out['best_xy'].double().cpu().numpy()
    assert (np.abs(xy[:, 0]) <= 3).all() and (np.abs(xy[:, 1]) <= 2).all()
    camel = np.vectorize(NOISE_FREE_FUNCTIONS['camel'])
    want = camel(xy[:, 0], xy[:, 1])
    got = out['ofe'].double().cpu().numpy()
    assert np.max(np.abs(got - want) / np.maximum(1, np.abs(want))) < 1e-5
    assert got.min() >= -1.0316285 - 1e-5
    # every arm, not only the winner: the loss the device ranks by is the float64 Camel value of the arm it drew
    aux = run_optimiser_repetitions(spec, 500, tpe_config(False), seed=3, details=True)
    axy = aux['arm_xy'].double().cpu().numpy()
    f = camel(axy[..., 0], axy[..., 1])
    loss = aux['ckpt_loss'][..., 0].double().cpu().numpy()
    assert np.max(np.abs(loss - f) / np.maximum(1, np.abs(f))) < 1e-5
    assert np.array_equal(aux['ofe'].cpu().numpy(), out['ofe'][:500].cpu().numpy())
    assert np.allclose(f.min(axis=1), got[:500], rtol=1e-5, atol=1e-5)


def test_history_evaluators_carry_the_arms_real_family():
    """hyperband_optimiser.py:99-100 keeps the evaluator that produced the round-best loss.  The device path rebuilds it;
    its shape family must be the one the arm drew on the device.  Six raw families that differ in start_shift identify
    themselves: with no initial noise the first point of an arm's curve is f(x, y) - start_shift."""
    from autotune.b200.engine import HyperbandSimSpec, arm_families, run_optimiser_repetitions, tpe_config
    from autotune.benchmarks import OptFunctionSimulationProblem
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    from autotune.core import ShapeFamily, UniformShapeFamilyScheduler
    from autotune.optimisers import HyperbandOptimiser, TpeOptimiser
    shifts = [0, 100, 200, 300, 400, 500]
    fams = [(1.0 + 0.1 * i, 5 + i, 3.0 * i, False, s, s + 200) for i, s in enumerate(shifts)]
    spec = HyperbandSimSpec('branin', fams, 81, 3, 0.0)
    seed = 99
    out = run_optimiser_repetitions(spec, 50, tpe_config(False), seed=seed, details=True)
    xy = out['arm_xy'].double().cpu().numpy()
    f = np.vectorize(NOISE_FREE_FUNCTIONS['branin'])(xy[..., 0], xy[..., 1])
    start = out['ckpt_loss'][..., 0].double().cpu().numpy()            # checkpoint 0 is position 0 (r = 1)
    seen = np.rint((f - start) / 100).astype(int)                       # index of the family the device used
    reps, arms = np.meshgrid(np.arange(50), np.arange(spec.n_arms), indexing='ij')
    assert np.array_equal(arm_families(spec, seed, reps, arms), seen)
    assert len(np.unique(seen)) == 6
    # through the optimiser API
    with redirect_stdout(io.StringIO()):
        problem = OptFunctionSimulationProblem('branin')
    sched = UniformShapeFamilyScheduler([ShapeFamily(None, *q) for q in fams], max_resources=81, init_noise=0)
    for opt in (HyperbandOptimiser(eta=3, max_iter=81, optimisation_func=lambda g: g.fval, is_simulation=True, scheduler=sched),
                TpeOptimiser(n_resources=1, max_iter=30, optimisation_func=lambda g: g.fval, is_simulation=True,
                             scheduler=sched)):
        np.random.seed(12)
        with redirect_stdout(io.StringIO()):
            opt.run_optimisation(problem)
        distinct = set()
        times = [r['time'] for r in spec.rounds()]
        for i, e in enumerate(opt.eval_history):
            ev = e.evaluator
            fxy = NOISE_FREE_FUNCTIONS['branin'](ev.arm.x, ev.arm.y)
            distinct.add(ev.start_shift)
            gap = fxy - e.optimisation_goals.fval
            if isinstance(opt, TpeOptimiser) or times[i] == 1:
                # read at r = 1: the recorded loss is the start value f - start_shift of the arm's real family
                assert abs(gap - ev.start_shift) < 1e-2, (i, gap, ev.start_shift)
            elif times[i] == 81:
                # read at r = R: a raw arm has landed on f_n = f - end_shift of its real family
                assert abs(gap - ev.end_shift) < 1e-2, (i, gap, ev.end_shift)
        if isinstance(opt, TpeOptimiser):
            assert len(distinct) > 1          # 30 random arms: several families


@pytest.mark.parametrize('R', [81, 9, 3])
def test_device_gamma_draws_match_the_gamma_law_at_a_million_samples(R):
    """High-power check of the Gamma sampler itself (Philox words -> Box-Muller -> Marsaglia-Tsang with the 16 + 16 bit accept
    uniform): the first step of 10^6 raw trajectories with f_1 = 0, f_n = -200, ml_agg = 1.5, up_spik = 15 and a steep time
    schedule (P = P11 = 0 at t = 1) is  -3 (g - 2)  for g > 2 and  15 / (1 + g)  for g < 2 (simulation_evaluator.py:102-115), so
    the draw g can be read back from the value.  One-sample KS against Gamma(shape(t=1), scale(t=1)) of :36-42: at n = 10^6 a
    deviation of 2e-3 in the CDF - e.g. a biased acceptance test - would be rejected."""
    from scipy import stats

    from autotune.b200.engine import HyperbandSimSpec
    from autotune.b200.trajectories import simulate_arms
    from oracle import gamma_process as gp
    n = 1000000
    spec = HyperbandSimSpec('branin', [(1.5, 40, 15, False, 0, 0)], R, 3, 0)
    raw = simulate_arms(spec, np.zeros(n, np.float32), np.full(n, -200.0, np.float32), np.zeros(n, np.int32),
                        np.arange(n, dtype=np.int64) + 12345, seed=99 + R)
    v = raw[:, 1].double().cpu().numpy()
    g = np.where(v < 0, 2.0 - v / 3.0, 15.0 / np.maximum(v, 1e-30) - 1.0)
    a, s = gp.gamma_shape_scale(1, R)
    res = stats.kstest(g, stats.gamma(a, scale=s).cdf)
    print(f'R={R}: shape {a:.4f} scale {s:.4f}: KS D = {res.statistic:.2e} (p = {res.pvalue:.3f}) over {n} device draws')
    assert res.statistic < 2.0e-3, (R, a, s, res, g.mean(), a * s)
    assert abs(g.mean() - a * s) < 5 * np.sqrt(a) * s / np.sqrt(n)          # 5 sigma on the mean
    assert abs((g > 2).mean() - stats.gamma(a, scale=s).sf(2.0)) < 2.5e-3       # share of down steps


def test_initial_noise_is_standard_normal_and_independent_of_the_hyperparameters():
    """An arm's initial noise z (benchmarks/opt_function_problem.py:138: f_1 = f + init_noise * N(0, 1)) is drawn from the same
    Philox block as its hyperparameters - its Box-Muller angle out of the low bytes the x / y / radius conversions never look
    at.  Read back from the first checkpoint of raw arms (f_1 - f + start_shift = init_noise * z): ~10^6 values, one-sample KS
    against N(0, 1), and no correlation with x, y or the radius' own uniform."""
    from scipy import stats

    from autotune.b200.engine import HyperbandSimSpec, run_optimiser_repetitions, tpe_config
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    noise, s0 = 10.0, 25.0
    spec = HyperbandSimSpec('branin', [(1.5, 10, 15, False, s0, 200)], 81, 3, noise)
    out = run_optimiser_repetitions(spec, 9000, tpe_config(False), seed=77, details=True)      # 9000 x 117 arms
    xy = out['arm_xy'].double().cpu().numpy().reshape(-1, 2)
    f1 = out['ckpt_loss'][:, :, 0].double().cpu().numpy().reshape(-1)
    f = np.vectorize(NOISE_FREE_FUNCTIONS['branin'])(xy[:, 0], xy[:, 1])
    z = (f1 - f + s0) / noise
    assert z.size >= 1000000
    # f_1 is a float32 of magnitude up to ~300: z is resolved to ~3e-6
    res = stats.kstest(z, 'norm')
    print(f'initial noise: n = {z.size}, mean {z.mean():.2e}, std {z.std():.5f}, KS D = {res.statistic:.2e} (p = {res.pvalue:.3f})')
    assert res.statistic < 2.0e-3, res
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    for other in (xy[:, 0], xy[:, 1], np.abs(z) * np.sign(xy[:, 0] - 2.5)):
        assert abs(np.corrcoef(z, other)[0, 1]) < 5e-3


def test_device_trajectories_match_the_reference_process_at_a_million_samples():
    """High-power check of the whole chain - word stream, Box-Muller, acceptance with the dithered 8-bit uniform, rejections
    that shift which words later steps consume, FP32 recurrence: 10^6 device trajectories (R = 27, one raw family, f_1 = 0,
    f_n = -200) against 10^6 trajectories of the reference's recurrence (core/simulation_evaluator.py:97-116) driven by
    numpy's own Gamma generator, two-sample KS at an early, a middle and the last-but-one position (D = 2.3e-3 would be
    p = 0.01 at these sizes)."""
    from scipy.stats import ks_2samp

    from autotune.b200.engine import HyperbandSimSpec
    from autotune.b200.trajectories import simulate_arms
    from oracle import gamma_process as gp
    R, n = 27, 1000000
    ml, nec, up = 1.5, 3.0, 15.0
    spec = HyperbandSimSpec('branin', [(ml, nec, up, False, 0, 0)], R, 3, 0)
    dev = simulate_arms(spec, np.zeros(n, np.float32), np.full(n, -200.0, np.float32), np.zeros(n, np.int32),
                        np.arange(n, dtype=np.int64) + 777, seed=4242).double().cpu().numpy()
    rng = np.random.default_rng(20261020)
    cur, f_n = np.zeros(n), -200.0
    ref = {}
    for t in range(1, R):                                   # the vectorised form of gamma_process.step, same expressions
        shape, scale = gp.gamma_shape_scale(t, R)
        g = rng.gamma(shape, scale, n)
        frac = t / (R - 1)
        m = cur + (g - 2) * ml * (f_n - cur) / 100
        down = m + (f_n - m) * frac ** nec
        u = cur + up / (1 + g)
        upv = np.full(n, f_n) if R - t == 1 else u + (f_n - u) * frac ** (1.1 * nec)
        cur = np.where(g > 2, down, np.where(g < 2, upv, cur))
        ref[t] = cur
    for t in (2, 9, R - 2):
        res = ks_2samp(dev[:, t], ref[t])
        print(f'position {t}: KS D = {res.statistic:.2e} (p = {res.pvalue:.3f}), means {dev[:, t].mean():.4f} / {ref[t].mean():.4f}')
        assert res.statistic < 3.0e-3, (t, res)
    assert np.allclose(dev[:, R - 1], f_n, atol=1e-3)        # every trajectory lands on its target
