"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/at2.h declares, and its host-side
helpers (bracket plan, constant tables, Savitzky-Golay rows, Philox block) agree with the reference-pinned oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from autotune.b200 import _native as nat
from oracle import gamma_process as gp
from oracle import hyperband as ohb
from tests.helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'at2.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(at2_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 10
    for name in sorted(declared):
        assert hasattr(nat.lib, name), f'{name} declared in include/at2.h but not exported'
    assert b'sm_100a' in nat.lib.at2_version()


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        nat.lib.at2_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def _plan_as_oracle(plan):
    out = []
    for b in range(plan.n_brackets):
        rounds = []
        for r in range(plan.bracket_first_round[b], plan.bracket_first_round[b + 1]):
            rounds.append((plan.round_n_eval[r], plan.round_time[r], plan.round_keep[r]))
        out.append((plan.bracket_n[b], rounds))
    return out


@pytest.mark.parametrize('R,eta', [(6, 3), (9, 3), (16, 2), (27, 3), (64, 4), (81, 3), (100, 3), (125, 5), (243, 3),
                                   (256, 2), (729, 3), (1000, 10), (1121, 3), (2187, 3), (2, 2), (5, 7), (50, 4)])
def test_bracket_plan_matches_oracle(R, eta):
    plan = nat.bracket_plan(R, eta)
    want = [(b['n'], [(rd['n_eval'], (rd['time'] - 1) % R + 1, rd['keep']) for rd in b['rounds']])
            for b in ohb.bracket_plan(R, eta)]   # time 0 reads fs[-1], i.e. position R
    assert _plan_as_oracle(plan) == want
    assert plan.n_arms == ohb.plan_arms(ohb.bracket_plan(R, eta))
    assert list(plan.ckpt_time[:plan.n_ckpts]) == sorted({(t - 1) % R + 1 for t in ohb.plan_checkpoints(ohb.bracket_plan(R, eta))})
    for r in range(plan.n_rounds):
        assert plan.ckpt_time[plan.round_ckpt[r]] == plan.round_time[r]
    # the resource count the hybrids record per evaluator is int(r_i) itself - 0, not R, where the read wraps (R = 729)
    assert list(plan.round_res[:plan.n_rounds]) == [rd['time'] for b in ohb.bracket_plan(R, eta) for rd in b['rounds']]
    if (R, eta) == (729, 3):
        assert plan.round_res[0] == 0 and plan.round_time[0] == 729


def test_bracket_plan_sweep_against_mpmath_smax():
    # exact integer floor(log_eta R) must equal the reference's 64-digit mpmath evaluation, incl. exact powers
    for eta in (2, 3, 4, 5, 7, 10):
        for R in list(range(2, 400)) + [eta ** k for k in range(1, 12) if eta ** k < 3000]:
            if R > 3000:
                continue
            want = ohb.bracket_plan(R, eta)
            plan = nat.Plan()
            rc = nat.lib.at2_bracket_plan(R, eta, C.byref(plan))
            if rc != 0:
                assert rc == -3, nat.lib.at2_last_error()    # capacity exceeded is the only acceptable failure
                continue
            assert [plan.bracket_n[b] for b in range(plan.n_brackets)] == [b['n'] for b in want], (R, eta)


def test_bracket_plan_against_reference_tables():
    d = np.load(f'{GOLDEN}/bracket_tables.npz')
    for R, eta in [(81, 3), (243, 3), (1121, 3), (16, 2), (6, 3)]:
        plan = nat.bracket_plan(R, eta)
        pre = f'R{R}_eta{eta}_'
        assert [plan.round_n_eval[r] for r in range(plan.n_rounds)] == d[pre + 'round_n'].tolist()
        assert [plan.round_keep[r] for r in range(plan.n_rounds)] == d[pre + 'round_keep'].tolist()
        assert [plan.round_time[r] for r in range(plan.n_rounds)] == [(int(v) - 1) % R + 1 for v in d[pre + 'round_res']]


def test_bracket_plan_rejects_bad_arguments():
    with pytest.raises(ValueError):
        nat.bracket_plan(1, 3)
    with pytest.raises(ValueError):
        nat.bracket_plan(81, 1)


def _cfg(families, func='rastrigin', noise=10.0):
    cfg = nat.SimConfig()
    cfg.func, cfg.n_families, cfg.init_noise = nat.FUNC_IDS[func], len(families), noise
    for i, (ml, nec, up, sm, s0, s1) in enumerate(families):
        cfg.fam[i] = nat.Family(ml, nec, up, s0, s1, int(sm), 0)
    return cfg


def _slots(cfg):
    slots = (C.c_int32 * 8)()
    n = nat.lib.at2_family_slots(C.byref(cfg), slots)
    return n, list(slots[:cfg.n_families])


FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]


@pytest.mark.parametrize('R', [27, 81, 243])
def test_tables_f64_bit_exact_with_python_floats(R):
    plan = nat.bracket_plan(R, 3)
    cfg = _cfg(FAM6)
    nbytes = nat.lib.at2_hb_tables_bytes(C.byref(plan), C.byref(cfg), 1)
    buf = np.zeros(nbytes // 8, np.float64)
    nat.check(nat.lib.at2_hb_tables_build(C.byref(plan), C.byref(cfg), 1, buf.ctypes.data))
    F, slot = _slots(cfg)                                             # record blocks: equal shapes share one
    assert (F, slot) == (3, [0, 1, 2, 0, 1, 2])
    rows = (R + 1) | 1                                                # csrc/at2_common.cuh at2_rec_rows
    rec = buf[:12 * F * rows].reshape(F, rows, 12)                    # 12-real records (csrc/at2_common.cuh AT2_REC)
    for f, fam in zip(slot, FAM6):
        for t in range(1, R):
            assert rec[f, t, 4] == (t / (R - 1)) ** fam[1]            # python float pow, simulation_evaluator.py:105
            assert rec[f, t, 5] == (t / (R - 1)) ** (1.1 * fam[1])    # :112
            a, s = gp.gamma_shape_scale(t, R)
            d = a - 1.0 / 3.0
            assert rec[f, t, 0] == d and rec[f, t, 3] == d * s
        assert np.isnan(rec[f, R, 1])                                 # sentinel row: acceptance can never pass
    # Savitzky-Golay rows at the checkpoints agree with scipy's operator to rounding
    W = gp.savgol_operator(R)
    sgw = 12 if plan.n_ckpts <= 12 else 24
    sg = buf[12 * F * rows:12 * F * rows + sgw * (R + 1)].reshape(R + 1, sgw)
    for c in range(plan.n_ckpts):
        row = plan.ckpt_time[c] - 1
        assert np.max(np.abs(sg[:R, c] - W[row])) < 1e-12
    assert np.all(sg[:, plan.n_ckpts:] == 0) and np.all(sg[R] == 0)


@pytest.mark.parametrize('R', [27, 81])
def test_tables_f32_folded_records(R):
    """Float tables (Philox paths): 12-real records {d, c, d log2e, d log2e + 32 | +-w, K1w, K0, P11 | spare}, row =
    f * rows + t with rows = (R + 1) | 1; the sign of w flags the plan's checkpoint positions; the folded-step table of the
    pre-drawn path {K1, K0, P11, 0} follows the Savitzky-Golay rows (csrc/at2_common.cuh)."""
    plan = nat.bracket_plan(R, 3)
    cfg = _cfg(FAM6)
    nbytes = nat.lib.at2_hb_tables_bytes(C.byref(plan), C.byref(cfg), 0)
    buf = np.zeros(nbytes // 4, np.float32)
    nat.check(nat.lib.at2_hb_tables_build(C.byref(plan), C.byref(cfg), 0, buf.ctypes.data))
    F, slot = _slots(cfg)
    sgw = 12 if plan.n_ckpts <= 12 else 24
    rows = (R + 1) | 1
    rec = buf[:12 * F * rows].reshape(F, rows, 12)
    assert np.all(rec[:, :, 8:] == 0)
    pw = buf[12 * F * rows + sgw * (R + 1):].reshape(F, rows, 4)
    ckpos = {plan.ckpt_time[c] - 1 for c in range(plan.n_ckpts)}
    f32 = np.float32
    for f, fam in zip(slot, FAM6):
        for t in range(1, R):
            a, s = gp.gamma_shape_scale(t, R)
            d = a - 1.0 / 3.0
            P, P11, ml = (t / (R - 1)) ** fam[1], (t / (R - 1)) ** (1.1 * fam[1]), fam[0] / 100.0
            w = d * s
            assert rec[f, t, 0] == f32(d) and rec[f, t, 2] == f32(d * 1.4426950408889634)
            assert rec[f, t, 3] == f32(d * 1.4426950408889634 + 32.0)
            assert rec[f, t, 4] == f32(-w if t in ckpos else w)
            assert rec[f, t, 5] == f32(w * (ml * (1.0 - P))) and rec[f, t, 6] == f32(P - 2.0 * ml * (1.0 - P))
            assert rec[f, t, 7] == f32(P11)
            assert tuple(pw[f, t]) == (f32(ml * (1.0 - P)), f32(P - 2.0 * ml * (1.0 - P)), f32(P11), f32(0))
            # the folded form is the reference's down step: k = a (1 - P) + P with a = (g - 2) ml_agg / 100
            g = 3.7
            assert abs((g * pw[f, t, 0] + pw[f, t, 1]) - ((g - 2) * ml * (1 - P) + P)) < 1e-6
        assert np.isnan(rec[f, R, 1]) and rec[f, R, 4] > 0       # sentinel: never accepts, never captures


def test_savgol_rows_all_rows_and_window_error():
    for R in (16, 27, 81, 243):
        rows = np.arange(R, dtype=np.int32)
        out = np.zeros((R, R))
        nat.check(nat.lib.at2_savgol_rows(R, R, rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                          out.ctypes.data_as(C.POINTER(C.c_double))))
        assert np.max(np.abs(out - gp.savgol_operator(R))) < 1e-12
        assert nat.lib.at2_savgol_window(R) == gp.savgol_window(R)
    rows = np.zeros(1, np.int32)
    out = np.zeros((1, 6))
    rc = nat.lib.at2_savgol_rows(6, 1, rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -1 and b'window_length' in nat.lib.at2_last_error()   # same failure the reference hits via scipy


def test_func_dims():
    assert [nat.lib.at2_func_dims(i) for i in range(-1, 7)] == [-1, 2, 2, 2, 2, 2, 4, -1]
    assert nat.FUNC_IDS['egg4'] == 5
