"""Batched entry points: whole Monte-Carlo jobs (many independent optimiser repetitions) as single kernel pipelines.

This replaces the reference's repetition loop (experiments/run_simulation.py:147-157) - the reference has no
equivalent call; everything it needs is described by the same objects the reference's loop builds each iteration
(problem, scheduler, optimiser).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import ops

FUNC_DOMAINS = {  # benchmarks/opt_function_problem.py:15-46
    'egg': (-512, 512, -512, 512), 'branin': (-5, 10, 1, 15), 'camel': (-3, 3, -2, 2),
    'wave': (-5.12, 5.12, -5.12, 5.12), 'rastrigin': (-5.12, 5.12, -5.12, 5.12),
    # extension (include/at2.h AT2_FUNC_EGG4): 4-D Egg-holder, all four coordinates on [-512, 512]; no reference oracle
    'egg4': (-512, 512, -512, 512)}


def _family_tuple(f):
    """Accept ShapeFamily / EvaluatorParams (arm first) or a bare 6-tuple (ml, nec, up, smooth, start, end)."""
    f = tuple(f)
    if len(f) >= 7:
        f = f[1:7]
    if len(f) == 3:
        f = f + (False, 0, 200)
    if len(f) != 6:
        raise ValueError(f'cannot interpret shape family {f!r}')
    return (float(f[0]), float(f[1]), float(f[2]), bool(f[3]), float(f[4]), float(f[5]))


class HyperbandSimSpec:
    """Everything that defines a Hyperband-on-simulation Monte-Carlo job except the repetition count and the seed."""

    def __init__(self, func_name, shape_families, max_iter=81, eta=3, init_noise=0.0, scheduler='uniform',
                 min_or_max='min'):
        if func_name not in nat.FUNC_IDS:
            raise KeyError(func_name)
        if scheduler not in ('uniform', 'round-robin'):
            raise ValueError(f'unknown scheduler {scheduler!r}')
        if min_or_max in (min, max):
            min_or_max = min_or_max.__name__
        if min_or_max not in ('min', 'max'):
            raise ValueError(f"optimization must be a built in function: min or max, instead {min_or_max} was supplied")
        self.func_name, self.R, self.eta = func_name, int(max_iter), int(eta)
        self.families = [_family_tuple(f) for f in shape_families]
        if not 1 <= len(self.families) <= nat.MAX_FAMILIES:
            raise ValueError(f'between 1 and {nat.MAX_FAMILIES} shape families are supported')
        self.init_noise, self.scheduler, self.min_or_max = float(init_noise), scheduler, min_or_max
        self.plan = nat.bracket_plan(self.R, self.eta)
        cfg = nat.SimConfig()
        cfg.func, cfg.n_families = nat.FUNC_IDS[func_name], len(self.families)
        cfg.scheduler = nat.SCHED_UNIFORM if scheduler == 'uniform' else nat.SCHED_ROUND_ROBIN
        cfg.descending = int(min_or_max == 'max')
        cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max = FUNC_DOMAINS[func_name]
        cfg.init_noise = self.init_noise
        for i, (ml, nec, up, sm, s0, s1) in enumerate(self.families):
            cfg.fam[i] = nat.Family(ml, nec, up, s0, s1, int(sm), 0)
        self.cfg = cfg
        if any(f[3] for f in self.families) and nat.lib.at2_savgol_window(self.R) > self.R:
            # the reference fails the same way, inside scipy, on the first smooth read-out
            raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
        self.plan_bytes, self.cfg_bytes = ops.struct_tensor(self.plan), ops.struct_tensor(self.cfg)
        self._tables = {}

    # ---- derived sizes ----
    @property
    def n_arms(self):
        return self.plan.n_arms

    @property
    def n_rounds(self):
        return self.plan.n_rounds

    @property
    def checkpoints(self):
        return [self.plan.ckpt_time[c] for c in range(self.plan.n_ckpts)]

    @property
    def arm_epochs_per_repetition(self):
        """Necessary arm-epochs: every arm's R-1 steps."""
        return self.plan.n_arms * (self.R - 1)

    def rounds(self):
        p = self.plan
        return [dict(bracket=b, n_eval=p.round_n_eval[r], keep=min(p.round_keep[r], p.round_n_eval[r]),
                     time=p.round_time[r], ckpt=p.round_ckpt[r])
                for b in range(p.n_brackets) for r in range(p.bracket_first_round[b], p.bracket_first_round[b + 1])]

    def tables(self, device, dtype=torch.float32, plan=None):
        """Constant tables on `device` (built once per (device, dtype, plan) by at2_hb_tables_build).  The tables belong
        to a plan: the Savitzky-Golay columns and the checkpoint flags of the step records follow its checkpoints, so a
        run under another plan (stand-alone TPE / random search: single_round_plan) gets its own."""
        device = torch.device(device)
        plan = self.plan if plan is None else plan
        key = (str(device), dtype, bytes(plan))
        if key not in self._tables:
            is64 = int(dtype == torch.float64)
            nbytes = nat.lib.at2_hb_tables_bytes(C.byref(plan), C.byref(self.cfg), is64)
            if nbytes < 0:
                raise nat.At2Error(nat.lib.at2_last_error().decode())
            host = torch.empty(nbytes // (8 if is64 else 4), dtype=dtype)
            nat.check(nat.lib.at2_hb_tables_build(C.byref(plan), C.byref(self.cfg), is64, host.data_ptr()))
            self._tables[key] = host.to(device)
        return self._tables[key]


def _cuda_device(device):
    if not torch.cuda.is_available():
        raise RuntimeError('no CUDA device: the B200 path has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def run_hyperband_repetitions(spec, n_reps, seed=0, rep_offset=0, device=None, details=False, early_stop=None):
    """`n_reps` independent Hyperband repetitions on the simulation problem of `spec`, one fused kernel launch.

    Returns a dict of device tensors: ofe[n], ofe_arm[n], best_xy[n,2] ([n,4] for the egg4 extension) and, with details=True, round_best[n,rounds],
    round_best_arm, survivors[n, n_surv], ckpt_loss[n, arms, ckpts].  Repetition i is keyed by (seed, rep_offset+i).

    early_stop: None (default) = the early-stopping kernel unless `details` asks for the checkpoint losses of every arm
    (which only the full simulation computes); True / False force one or the other.  Both give the same OFEs, winning arms,
    round-bests and survivor lists bit for bit (tests/test_hb_gpu.py) - early stopping just does not simulate what no round
    reads (benchmarks/opt_function_simulation_problem.py:67-76 simulates every arm to R on its first evaluate()).
    """
    device = _cuda_device(device)
    if early_stop is None:
        early_stop = not details
    # straight to the library (ops.hb_sim_direct: the body of torch.ops.at2.hb_sim_f32 / hb_sim_early_stop without the
    # dispatcher round trip); early stopping simulates arms only as far as the round that reads them needs - same streams,
    # bit-identical outputs
    outs = ops.hb_sim_direct(spec.tables(device, torch.float32), spec.plan, spec.cfg, int(n_reps), int(rep_offset), int(seed),
                             bool(details), bool(early_stop))
    names = [k for k in ops.HB_OUTPUT_NAMES if k != 'ckpt_loss'] if early_stop else ops.HB_OUTPUT_NAMES
    return dict(zip(names, outs))


def arm_families(spec, seed, rep_ids, arm_ids):
    """Index into spec.families of the shape family each (repetition, arm) pair drew on the device.

    Round-robin: arm % F (core/shape_family_scheduler.py:75).  Uniform (:61): floor(F * z / 2**32) with z the third word of
    block 0 of the arm's own Philox stream, keyed by (seed; global repetition id * n_arms + arm) - the draw of
    csrc/sim_core.cuh draw_arm / arm_family, recomputed here through the ABI's at2_philox4x32_10 so that callers can rebuild an
    arm's evaluator with the family it really had."""
    F, A = len(spec.families), spec.n_arms
    rep_ids, arm_ids = np.broadcast_arrays(np.asarray(rep_ids, np.int64), np.asarray(arm_ids, np.int64))
    out = np.empty(arm_ids.shape, np.int64)
    if spec.scheduler != 'uniform':
        out[...] = arm_ids % F
        return out
    key = (C.c_uint32 * 2)(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ctr, res = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    for idx in np.ndindex(*arm_ids.shape):
        gid = int(rep_ids[idx]) * A + int(arm_ids[idx])
        ctr[0], ctr[1], ctr[2], ctr[3] = 0, gid & 0xFFFFFFFF, (gid >> 32) & 0xFFFFFFFF, 0      # AT2_STREAM_ARM
        nat.lib.at2_philox4x32_10(ctr, key, res)
        out[idx] = (int(res[2]) * F) >> 32
    return out


def host_start_and_target(spec, xy, fam, z):
    """f_1 = f(x,y) + noise*z - start_shift and f_n = f(x,y) - end_shift per arm, in float64 on the host
    (opt_function_simulation_problem.py:61-64).  Evaluated arm by arm through numpy *scalar* calls, as the reference
    does: numpy's vectorised cos/sin loops are not bit-identical to its scalar path for every input."""
    from autotune.benchmarks.opt_function_problem import NOISE_FREE_FUNCTIONS
    func = NOISE_FREE_FUNCTIONS[spec.func_name]
    xy, fam, z = np.asarray(xy, np.float64), np.asarray(fam), np.asarray(z, np.float64)
    f1, fn = np.empty(fam.shape, np.float64), np.empty(fam.shape, np.float64)
    for idx in np.ndindex(*fam.shape):
        f = func(np.float64(xy[idx + (0,)]), np.float64(xy[idx + (1,)]))
        ml, nec, up, sm, s0, s1 = spec.families[int(fam[idx])]
        fn[idx] = f - s1
        f1[idx] = (f + spec.init_noise * z[idx] * 1) - s0
    return f1, fn


def run_hyperband_predrawn(spec, f1, fn, family, gamma, xy=None, device=None, dtype=torch.float64):
    """Parity entry: recurrence + read-out + halving on pre-drawn per-arm inputs (host arrays or tensors).

    f1, fn: [n, A]; family: [n, A] int; gamma: [n, A, R-1]; xy: [n, A, 2] (echoed into best_xy only)."""
    device = _cuda_device(device)

    def dev(a, dt):
        return torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(device=device, dtype=dt).contiguous()
    f1, fn, gamma = dev(f1, dtype), dev(fn, dtype), dev(gamma, dtype)
    family = dev(family, torch.int32)
    xy = dev(xy, dtype) if xy is not None else torch.zeros(*f1.shape, 2, device=device, dtype=dtype)
    outs = torch.ops.at2.hb_sim_predrawn(spec.tables(device, dtype), spec.plan_bytes, spec.cfg_bytes, f1, fn, family,
                                         gamma, xy)
    return dict(zip(ops.HB_OUTPUT_NAMES, outs))


# ---- closest-known-loss-function approximation -------------------------------------------------------------------------
def domain_desc(domain, hyperparams_to_opt):
    """at2_domain for a `Domain` of continuous `Param`s: declaration order + the attribute order of a drawn arm
    (core/arm.py:29-55: non-optimised defaults first, then the optimised hyperparameters)."""
    names = list(domain.hyperparams_names())
    if not 1 <= len(names) <= nat.MAX_PARAMS:
        raise ValueError(f'between 1 and {nat.MAX_PARAMS} hyperparameters are supported')
    desc = nat.DomainDesc()
    desc.n_params = len(names)
    for i, n in enumerate(names):
        prm = domain[n]
        if getattr(prm, 'distrib', 'uniform') != 'uniform':
            raise NotImplementedError(f"hyperparameter {n}: only distrib='uniform' is on the device path")
        is_log = prm.scale == 'log'
        if is_log and prm.logbase != np.e:
            raise NotImplementedError(f'hyperparameter {n}: only logbase e is on the device path')
        optimised = n in hyperparams_to_opt
        if not optimised and prm.init_val is None:
            raise ValueError(f"No default value for param {n} was supplied")
        desc.p[i] = nat.ParamDesc(float(prm.min_val), float(prm.max_val), float(prm.interval or 0.0),
                                  float(prm.init_val) if prm.init_val is not None else 0.0, int(is_log), int(optimised))
    order = [i for i, n in enumerate(names) if n not in hyperparams_to_opt] + \
            [i for i, n in enumerate(names) if n in hyperparams_to_opt]
    for k, i in enumerate(order):
        desc.order[k] = i
    return desc, names


class KnownFnSpec:
    """A closest-known-function problem on the device: the recorded arms and their loss curves.

    known_fs: {Arm: [loss_1 .. loss_L]} as the reference's KnownFnProblem takes it (ragged curves are padded with NaN up
    to the longest; reading a padded position raises like the reference's IndexError would)."""

    def __init__(self, known_fs, domain, hyperparams_to_opt, device=None):
        self.desc, self.names = domain_desc(domain, hyperparams_to_opt)
        self.domain_bytes = ops.struct_tensor(self.desc)
        arms = list(known_fs.keys())
        if not arms:
            raise ValueError('known_fs is empty')
        self.arms = arms
        db = np.asarray([[float(a[n]) for n in self.names] for a in arms], np.float64)
        self.lengths = np.asarray([len(known_fs[a]) for a in arms], np.int64)
        L = int(self.lengths.max())
        curves = np.full((len(arms), L), np.nan)
        for i, a in enumerate(arms):
            curves[i, :self.lengths[i]] = np.asarray(known_fs[a], np.float64)
        device = _cuda_device(device)
        self.device = device
        self.db = torch.from_numpy(db).to(device)
        self.curves = torch.from_numpy(curves).to(device)

    def _domain_bytes_for(self, order):
        """Domain descriptor with another summation order (the squared error is summed in the attribute order of the
        PROPOSED arm, known_loss_fn_problem.py:36-48: any order a caller's Arm happens to have)."""
        if order is None or list(order) == [self.desc.order[k] for k in range(len(self.names))]:
            return self.domain_bytes
        if sorted(order) != list(range(len(self.names))):
            raise ValueError(f'order must be a permutation of range({len(self.names)})')
        desc = nat.DomainDesc.from_buffer_copy(bytes(self.desc))
        for k, i in enumerate(order):
            desc.order[k] = int(i)
        return ops.struct_tensor(desc)

    def lookup(self, queries, times=None, order=None):
        """Nearest recorded arm of each query row (raw values in domain order); with `times` also the loss read-out.
        `order`: summation order of the squared-error terms as indices into the domain's hyperparameters (default: the
        attribute order of a drawn arm)."""
        domain_bytes = self._domain_bytes_for(order)
        if torch.is_tensor(queries):
            q = queries.to(self.device, torch.float64).contiguous()
        else:
            q = torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float64)).to(self.device)
        if times is None:
            idx, dist, _ = torch.ops.at2.nn_lookup(domain_bytes, self.db, q, self.db.new_empty((0, 0)),
                                                   torch.empty(0, device=self.device, dtype=torch.int32))
            return idx, dist
        t = torch.as_tensor(np.asarray(times, np.int32)).to(self.device)
        return tuple(torch.ops.at2.nn_lookup(domain_bytes, self.db, q, self.curves, t))


def run_knownfn_hyperband(kspec, max_iter, eta, n_reps, seed=0, rep_offset=0, min_or_max='min', predrawn=None,
                          details=False):
    """`n_reps` Hyperband repetitions on the closest-known-function problem, one fused launch."""
    if min_or_max in (min, max):
        min_or_max = min_or_max.__name__
    plan = nat.bracket_plan(max_iter, eta)
    L_min = int(kspec.lengths.min())
    for c in range(plan.n_ckpts):
        if plan.ckpt_time[c] > L_min:
            raise IndexError('list index out of range')   # what the reference raises on a too-short recorded curve
    pre = kspec.db.new_empty(0) if predrawn is None else \
        torch.as_tensor(np.ascontiguousarray(predrawn, dtype=np.float64)).to(kspec.device).contiguous()
    outs = torch.ops.at2.hb_knownfn(ops.struct_tensor(plan), kspec.domain_bytes, kspec.db, kspec.curves,
                                    min_or_max == 'max', int(n_reps), int(rep_offset), int(seed), pre, bool(details))
    return dict(zip([k for k in ops.KF_OUTPUT_NAMES if details or k in ('ofe', 'ofe_arm', 'best_arm_vals')], outs))


# ---- TPE / hybrids ---------------------------------------------------------------------------------------------------------
def tpe_config(enabled=True, transfer='none', threshold=0, n_startup=10, n_candidates=24, gamma=0.25, prior_weight=1.0,
               lf=25):
    """at2_tpe_config with hyperopt 0.2.4's tpe.suggest defaults and the reference's n_startup_jobs base of 10
    (tpe_optimiser.py:52)."""
    if transfer not in nat.TRANSFER_IDS:
        raise ValueError(f'unknown transfer mode {transfer!r}')
    return nat.TpeConfig(int(enabled), int(n_startup), int(n_candidates), int(lf), nat.TRANSFER_IDS[transfer],
                         int(threshold or 0), float(gamma), float(prior_weight))


def single_round_plan(R, n_evals, n_resources):
    plan = nat.Plan()
    rc = nat.lib.at2_single_round_plan(int(R), int(n_evals), int(n_resources), C.byref(plan))
    if rc != 0:
        msg = nat.lib.at2_last_error().decode()
        raise (IndexError if 'index out of range' in msg else ValueError)(msg)
    return plan


def run_optimiser_repetitions(spec, n_reps, tpe, seed=0, rep_offset=0, device=None, details=False, plan=None):
    """Monte-Carlo repetitions of any optimiser of the family {Hyperband, Hyperband+TPE (+transfer), TPE, random search}
    on the simulation problem of `spec`: one fused launch (at2_tpe_sim_f32).  `plan` overrides the Hyperband plan of
    `spec` (stand-alone TPE / random search use single_round_plan)."""
    device = _cuda_device(device)
    plan = spec.plan if plan is None else plan
    outs = ops.tpe_sim_direct(spec.tables(device, torch.float32, plan), plan, spec.cfg, tpe, int(n_reps), int(rep_offset),
                              int(seed), bool(details))
    return dict(zip(ops.HB_OUTPUT_NAMES + ['arm_xy', 'obs_loss', 'n_injected'] if details else ops.HB_OUTPUT_NAMES, outs))


def run_hybrid_repetitions(spec, n_reps, transfer='none', threshold=None, seed=0, rep_offset=0, device=None,
                           details=False):
    """HybridHyperbandTpe{,NoTransfer,TransferAll,TransferLongest,TransferSame,TransferThreshold}Optimiser."""
    return run_optimiser_repetitions(spec, n_reps, tpe_config(True, transfer, threshold), seed, rep_offset, device, details)


def run_tpe_repetitions(spec, n_reps, n_resources, max_iter, seed=0, rep_offset=0, device=None, details=False):
    """Stand-alone TpeOptimiser(n_resources, max_iter): max_iter sequential evaluations at n_resources."""
    plan = single_round_plan(spec.R, max_iter, n_resources)
    return run_optimiser_repetitions(spec, n_reps, tpe_config(True), seed, rep_offset, device, details, plan)


def run_random_repetitions(spec, n_reps, n_resources, max_iter, seed=0, rep_offset=0, device=None, details=False):
    """RandomOptimiser(n_resources, max_iter): max_iter random arms evaluated at n_resources."""
    plan = single_round_plan(spec.R, max_iter, n_resources)
    return run_optimiser_repetitions(spec, n_reps, tpe_config(False), seed, rep_offset, device, details, plan)
