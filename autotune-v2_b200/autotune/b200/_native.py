"""ctypes binding of the C ABI in include/at2.h (libat2_b200.so, built in-tree by __graft_entry__.build()).

There is deliberately no fallback: if the shared library is missing, importing this module raises, and every compute
entry point raises when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))            # .../autotune-v2_b200
LIB_PATH = os.path.join(PKG_ROOT, 'lib', 'libat2_b200.so')

MAX_BRACKETS, MAX_ROUNDS, MAX_CKPTS, MAX_FAMILIES = 16, 64, 16, 8
MAX_PARAMS = 16
MAX_PEERS = 16
FUNC_IDS = {'egg': 0, 'branin': 1, 'camel': 2, 'wave': 3, 'rastrigin': 4,
            'egg4': 5}   # egg4: 4-D extension (include/at2.h AT2_FUNC_EGG4), Hyperband entry points only
SCHED_UNIFORM, SCHED_ROUND_ROBIN = 0, 1


class Plan(C.Structure):
    _fields_ = [('R', C.c_int32), ('eta', C.c_int32), ('n_brackets', C.c_int32), ('n_rounds', C.c_int32),
                ('n_arms', C.c_int32), ('n_ckpts', C.c_int32), ('n_surv', C.c_int32),
                ('bracket_n', C.c_int32 * MAX_BRACKETS), ('bracket_first_arm', C.c_int32 * MAX_BRACKETS),
                ('bracket_first_round', C.c_int32 * (MAX_BRACKETS + 1)),
                ('round_n_eval', C.c_int32 * MAX_ROUNDS), ('round_keep', C.c_int32 * MAX_ROUNDS),
                ('round_time', C.c_int32 * MAX_ROUNDS), ('round_ckpt', C.c_int32 * MAX_ROUNDS),
                ('ckpt_time', C.c_int32 * MAX_CKPTS), ('bracket_r0', C.c_double * MAX_BRACKETS),
                ('round_res', C.c_int32 * MAX_ROUNDS)]


class Family(C.Structure):
    _fields_ = [('ml_agg', C.c_double), ('nec_agg', C.c_double), ('up_spik', C.c_double),
                ('start_shift', C.c_double), ('end_shift', C.c_double), ('is_smooth', C.c_int32), ('_pad', C.c_int32)]


class SimConfig(C.Structure):
    _fields_ = [('func', C.c_int32), ('n_families', C.c_int32), ('scheduler', C.c_int32), ('descending', C.c_int32),
                ('x_min', C.c_double), ('x_max', C.c_double), ('y_min', C.c_double), ('y_max', C.c_double),
                ('init_noise', C.c_double), ('fam', Family * MAX_FAMILIES)]


class ParamDesc(C.Structure):
    _fields_ = [('min_val', C.c_double), ('max_val', C.c_double), ('interval', C.c_double), ('default_val', C.c_double),
                ('is_log', C.c_int32), ('optimised', C.c_int32)]


class DomainDesc(C.Structure):
    _fields_ = [('n_params', C.c_int32), ('_pad', C.c_int32), ('p', ParamDesc * MAX_PARAMS),
                ('order', C.c_int32 * MAX_PARAMS)]


class TpeConfig(C.Structure):
    _fields_ = [('enabled', C.c_int32), ('n_startup', C.c_int32), ('n_candidates', C.c_int32), ('lf', C.c_int32),
                ('transfer', C.c_int32), ('threshold', C.c_int32), ('gamma', C.c_double), ('prior_weight', C.c_double)]


TRANSFER_IDS = {'none': 0, 'all': 1, 'longest': 2, 'same': 3, 'threshold': 4}


class HbOutputs(C.Structure):
    _fields_ = [('d_ofe', C.c_void_p), ('d_ofe_arm', C.c_void_p), ('d_best_xy', C.c_void_p),
                ('d_round_best', C.c_void_p), ('d_round_best_arm', C.c_void_p), ('d_survivors', C.c_void_p),
                ('d_ckpt_loss', C.c_void_p)]


class Gather(C.Structure):
    _fields_ = [('d_target', C.c_void_p * MAX_PEERS), ('n_targets', C.c_int32), ('is_multicast', C.c_int32),
                ('base', C.c_int64)]


class At2Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(there is no CPU fallback for the B200 path)')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        'at2_version': (C.c_char_p, []),
        'at2_last_error': (C.c_char_p, []),
        'at2_last_launch_count': (i32, []),
        'at2_func_dims': (i32, [i32]),
        'at2_bracket_plan': (C.c_int, [i32, i32, C.POINTER(Plan)]),
        'at2_savgol_window': (i32, [i32]),
        'at2_savgol_rows': (C.c_int, [i32, i32, C.POINTER(i32), C.POINTER(C.c_double)]),
        'at2_hb_tables_bytes': (i64, [C.POINTER(Plan), C.POINTER(SimConfig), i32]),
        'at2_hb_tables_build': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), i32, vp]),
        'at2_hb_sim_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, i64, u64,
                                     C.POINTER(HbOutputs), vp]),
        'at2_hb_sim_gather_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, i64, u64, C.POINTER(HbOutputs),
                                            C.POINTER(Gather), vp]),
        'at2_hb_sim_early_stop_gather_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, i64, u64,
                                                       C.POINTER(HbOutputs), C.POINTER(Gather), vp]),
        'at2_debug_option': (C.c_int, [C.c_char_p, i32]),
        'at2_hb_sim_early_stop_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, i64, u64,
                                                C.POINTER(HbOutputs), vp]),
        'at2_hb_sim_predrawn_f64': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, vp, vp, vp, vp, vp,
                                              C.POINTER(HbOutputs), vp]),
        'at2_hb_sim_predrawn_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, vp, vp, vp, vp, vp,
                                              C.POINTER(HbOutputs), vp]),
        'at2_philox4x32_10': (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        'at2_fma_peak_probe': (C.c_int, [i32, i32, i64, i32, vp, vp]),
        'at2_sim_trajectories_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), vp, i64, vp, vp, vp, vp, u64, vp, vp]),
        'at2_single_round_plan': (C.c_int, [i32, i32, i32, C.POINTER(Plan)]),
        'at2_tpe_sim_f32': (C.c_int, [C.POINTER(Plan), C.POINTER(SimConfig), C.POINTER(TpeConfig), vp, i64, i64, u64,
                                      C.POINTER(HbOutputs), vp, vp, vp, vp]),
        'at2_tpe_score_f32': (C.c_int, [vp, vp, i32, i32, C.c_double, C.c_double, C.POINTER(TpeConfig), vp, i32, vp, vp, vp,
                                        vp, vp]),
        'at2_nn_workspace_bytes': (i64, [i64, i32, i32]),
        'at2_nn_lookup_f64': (C.c_int, [C.POINTER(DomainDesc), vp, i32, vp, i64, vp, vp, vp, i32, vp, vp, vp, i64, vp]),
        'at2_hb_knownfn_f64': (C.c_int, [C.POINTER(Plan), C.POINTER(DomainDesc), vp, i32, vp, i32, i32, i64, i64, u64,
                                         vp, C.POINTER(HbOutputs), vp, vp, vp]),
        'at2_nn_lookup_excluding_f64': (C.c_int, [C.POINTER(DomainDesc), vp, i32, vp, i64, vp, vp, vp, vp, i64, vp]),
        'at2_order_profile_f32': (C.c_int, [vp, i32, i32, vp, vp, vp]),
        'at2_ks_statistic_f32': (C.c_int, [vp, i64, vp, i64, vp, vp]),
        'at2_kde_workspace_floats': (i64, [i64, i32]),
        'at2_kde_f32': (C.c_int, [vp, i64, C.c_double, C.c_double, i32, C.c_double, i32, vp, vp, i64, vp]),
        'at2_kde_partial_f32': (C.c_int, [vp, i64, C.c_double, C.c_double, i32, C.c_double, i32, C.POINTER(C.c_void_p), i32, i32,
                                          i64, vp, i64, vp]),
        'at2_segmented_topk_f32': (C.c_int, [vp, vp, vp, i64, i32, i32, vp, vp]),
        'at2_segmented_topk_f64': (C.c_int, [vp, vp, vp, i64, i32, i32, vp, vp]),
        'at2_family_slots': (i32, [C.POINTER(SimConfig), C.POINTER(i32)]),
        'at2_kde_combine_f64': (C.c_int, [vp, i32, i32, i64, C.c_double, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib, sig


lib, SIGNATURES = _load()


def check(rc):
    if rc != 0:
        raise At2Error(f'libat2_b200 error {rc}: {lib.at2_last_error().decode()}')


def bracket_plan(R, eta):
    plan = Plan()
    rc = lib.at2_bracket_plan(int(R), int(eta), C.byref(plan))
    if rc != 0:
        raise ValueError(lib.at2_last_error().decode())
    return plan


class debug_options:
    """Context manager over at2_debug_option (include/at2.h): tile / CTA shapes and code-path switches for tests and
    measurements - `with debug_options(hb_no_split=1): ...`.  Everything is reset on exit (the options are process-wide)."""

    def __init__(self, **opts):
        self.opts = opts

    def __enter__(self):
        for k, v in self.opts.items():
            check(lib.at2_debug_option(k.encode(), int(v)))
        return self

    def __exit__(self, *exc):
        check(lib.at2_debug_option(b'reset', 0))
        return False
