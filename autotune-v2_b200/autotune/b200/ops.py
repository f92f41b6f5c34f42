"""PyTorch custom ops (`torch.ops.at2.*`) over the C ABI of libat2_b200.

Each op allocates its outputs as torch tensors on the inputs' CUDA device and enqueues the library call on torch's
current stream of that device, so the ops compose with torch streams/events like any other CUDA op.  Struct arguments
(`at2_plan`, `at2_sim_config`) travel as CPU uint8 tensors holding the struct bytes.
"""
import ctypes as C
from typing import List

import torch

from . import _native as nat


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'{what} must live on a CUDA device (the B200 path has no CPU fallback); got {t.device}')


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _struct(cls, t: torch.Tensor):
    if t.device.type != 'cpu' or t.dtype != torch.uint8 or t.numel() != C.sizeof(cls):
        raise RuntimeError(f'expected a CPU uint8 tensor of {C.sizeof(cls)} bytes for {cls.__name__}')
    return cls.from_buffer_copy(t.numpy().tobytes())


def struct_tensor(obj) -> torch.Tensor:
    return torch.frombuffer(bytearray(bytes(obj)), dtype=torch.uint8).clone()


def _outputs(plan, n_reps, device, dtype, details, ckpt_loss=True, dims=2):
    o = {'ofe': torch.empty(n_reps, device=device, dtype=dtype),
         'ofe_arm': torch.empty(n_reps, device=device, dtype=torch.int32),
         'best_xy': torch.empty(n_reps, dims, device=device, dtype=dtype)}
    if details:
        o['round_best'] = torch.empty(n_reps, plan.n_rounds, device=device, dtype=dtype)
        o['round_best_arm'] = torch.empty(n_reps, plan.n_rounds, device=device, dtype=torch.int32)
        o['survivors'] = torch.empty(n_reps, plan.n_surv, device=device, dtype=torch.int32)
        if ckpt_loss:
            o['ckpt_loss'] = torch.empty(n_reps, plan.n_arms, plan.n_ckpts, device=device, dtype=dtype)
    ptr = lambda k: o[k].data_ptr() if k in o else None  # noqa: E731
    outs = nat.HbOutputs(ptr('ofe'), ptr('ofe_arm'), ptr('best_xy'), ptr('round_best'), ptr('round_best_arm'),
                         ptr('survivors'), ptr('ckpt_loss'))
    order = ['ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors', 'ckpt_loss']
    return outs, [o[k] for k in order if k in o]


HB_OUTPUT_NAMES = ['ofe', 'ofe_arm', 'best_xy', 'round_best', 'round_best_arm', 'survivors', 'ckpt_loss']


def hb_sim_direct(tables: torch.Tensor, plan, cfg, n_reps: int, rep_offset: int, seed: int, details: bool,
                  early_stop: bool) -> List[torch.Tensor]:
    """The body of at2::hb_sim_f32 / at2::hb_sim_early_stop on the ctypes structs themselves: what the batched entry points
    (autotune.b200.engine) call - no dispatcher round trip and no struct (de)serialisation, which is most of the host time
    of a reference-sized job (7 000 repetitions are ~0.1 ms of kernel)."""
    _require_cuda(tables, 'tables')
    with torch.cuda.device(tables.device):
        outs, tensors = _outputs(plan, n_reps, tables.device, torch.float32, details, ckpt_loss=not early_stop,
                                 dims=nat.lib.at2_func_dims(cfg.func))
        if n_reps > 0:   # an empty job is an empty result, as the reference's loop over range(0) would give
            fn = nat.lib.at2_hb_sim_early_stop_f32 if early_stop else nat.lib.at2_hb_sim_f32
            nat.check(fn(C.byref(plan), C.byref(cfg), tables.data_ptr(), n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF,
                         C.byref(outs), _stream_ptr(tables.device)))
    return tensors


@torch.library.custom_op('at2::hb_sim_f32', mutates_args=(), device_types='cuda')
def hb_sim_f32(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, n_reps: int, rep_offset: int,
               seed: int, details: bool) -> List[torch.Tensor]:
    """Fused Philox/Gamma simulation + successive halving for `n_reps` Hyperband repetitions (at2_hb_sim_f32)."""
    return hb_sim_direct(tables, _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes), n_reps, rep_offset, seed,
                         details, False)


def hb_sim_gather_direct(tables: torch.Tensor, plan, cfg, n_reps: int, rep_offset: int,
                  seed: int, gathered: torch.Tensor, gather_ptrs: List[int], gather_base: int, multicast: bool = False,
                  early_stop: bool = False) -> List[torch.Tensor]:
    """at2::hb_sim_f32 (or, with early_stop, at2::hb_sim_early_stop) fused with the all-gather of the OFEs
    (at2_hb_sim_gather_f32 / at2_hb_sim_early_stop_gather_f32): the OFE of repetition i is also stored to element
    gather_base + i of every buffer in `gather_ptrs` (device addresses of the peers' symmetric-memory buffers), or - with
    multicast=True - through the one NVSwitch multicast address in `gather_ptrs` (multimem.st).  `gathered` is this rank's
    own buffer (declared mutated; it must be among the targets)."""
    _require_cuda(tables, 'tables')
    if not 1 <= len(gather_ptrs) <= nat.MAX_PEERS or (multicast and len(gather_ptrs) != 1):
        raise RuntimeError(f'between 1 and {nat.MAX_PEERS} gather buffers (exactly one multicast address)')
    if gathered.dtype != torch.float32 or gather_base < 0 or gather_base + n_reps > gathered.numel():
        raise RuntimeError('gathered must be a float32 buffer holding elements gather_base .. gather_base + n_reps')
    with torch.cuda.device(tables.device):
        outs, tensors = _outputs(plan, n_reps, tables.device, torch.float32, False, dims=nat.lib.at2_func_dims(cfg.func))
        if n_reps > 0:
            g = nat.Gather()
            for i, ptr in enumerate(gather_ptrs):
                g.d_target[i] = ptr
            g.n_targets, g.is_multicast, g.base = len(gather_ptrs), int(multicast), gather_base
            fn = nat.lib.at2_hb_sim_early_stop_gather_f32 if early_stop else nat.lib.at2_hb_sim_gather_f32
            nat.check(fn(C.byref(plan), C.byref(cfg), tables.data_ptr(), n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF,
                         C.byref(outs), C.byref(g), _stream_ptr(tables.device)))
    return tensors


@torch.library.custom_op('at2::hb_sim_gather', mutates_args=('gathered',), device_types='cuda')
def hb_sim_gather(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, n_reps: int, rep_offset: int,
                  seed: int, gathered: torch.Tensor, gather_ptrs: List[int], gather_base: int, multicast: bool = False,
                  early_stop: bool = False) -> List[torch.Tensor]:
    """hb_sim_gather_direct as a custom op (struct arguments as CPU uint8 tensors)."""
    return hb_sim_gather_direct(tables, _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes), n_reps, rep_offset,
                                seed, gathered, gather_ptrs, gather_base, multicast, early_stop)


@torch.library.custom_op('at2::hb_sim_early_stop', mutates_args=(), device_types='cuda')
def hb_sim_early_stop(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, n_reps: int, rep_offset: int,
                      seed: int, details: bool) -> List[torch.Tensor]:
    """at2::hb_sim_f32 with early stopping (at2_hb_sim_early_stop_f32): identical outputs (minus ckpt_loss) at a fraction
    of the simulated epochs."""
    return hb_sim_direct(tables, _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes), n_reps, rep_offset, seed,
                         details, True)


@hb_sim_f32.register_fake
def _(tables, plan_bytes, cfg_bytes, n_reps, rep_offset, seed, details):
    raise RuntimeError('at2::hb_sim_f32 has no meta/CPU implementation')


@torch.library.custom_op('at2::hb_sim_predrawn', mutates_args=(), device_types='cuda')
def hb_sim_predrawn(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, f1: torch.Tensor,
                    fn: torch.Tensor, family: torch.Tensor, gamma: torch.Tensor, xy: torch.Tensor) -> List[torch.Tensor]:
    """Recurrence + read-out + successive halving on pre-drawn inputs; dtype of `f1` picks the f64 (bit-exact
    reference operation order) or f32 variant (at2_hb_sim_predrawn_f64 / _f32)."""
    for name, t in (('tables', tables), ('f1', f1), ('fn', fn), ('family', family), ('gamma', gamma), ('xy', xy)):
        _require_cuda(t, name)
    plan, cfg = _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes)
    dtype = f1.dtype
    if dtype not in (torch.float32, torch.float64):
        raise RuntimeError('f1 must be float32 or float64')
    n_reps = f1.shape[0]
    want = {'f1': (n_reps, plan.n_arms), 'fn': (n_reps, plan.n_arms), 'family': (n_reps, plan.n_arms),
            'gamma': (n_reps, plan.n_arms, plan.R - 1), 'xy': (n_reps, plan.n_arms, 2)}
    for name, t in (('f1', f1), ('fn', fn), ('family', family), ('gamma', gamma), ('xy', xy)):
        if tuple(t.shape) != want[name] or not t.is_contiguous():
            raise RuntimeError(f'{name} must be contiguous with shape {want[name]}, got {tuple(t.shape)}')
        if name != 'family' and t.dtype != dtype:
            raise RuntimeError(f'{name} must have dtype {dtype}')
    if family.dtype != torch.int32 or tables.dtype != dtype:
        raise RuntimeError('family must be int32 and tables must match the dtype of f1')
    fn_c = nat.lib.at2_hb_sim_predrawn_f64 if dtype == torch.float64 else nat.lib.at2_hb_sim_predrawn_f32
    with torch.cuda.device(f1.device):
        outs, tensors = _outputs(plan, n_reps, f1.device, dtype, True)
        nat.check(fn_c(C.byref(plan), C.byref(cfg), tables.data_ptr(), n_reps, f1.data_ptr(), fn.data_ptr(),
                       family.data_ptr(), gamma.data_ptr(), xy.data_ptr(), C.byref(outs), _stream_ptr(f1.device)))
    return tensors


@torch.library.custom_op('at2::fma_peak_probe', mutates_args=('sink',), device_types='cuda')
def fma_peak_probe(sink: torch.Tensor, blocks: int, threads: int, iters: int, variant: int) -> None:
    _require_cuda(sink, 'sink')
    with torch.cuda.device(sink.device):
        nat.check(nat.lib.at2_fma_peak_probe(blocks, threads, iters, variant, sink.data_ptr(), _stream_ptr(sink.device)))


def kde_direct(x: torch.Tensor, start: float, end: float, grid_points: int, bandwidth: float, kernel: int) -> torch.Tensor:
    """The body of at2::kde, called directly by autotune.b200.epdf.density (see hb_sim_direct)."""
    _require_cuda(x, 'x')
    if x.dtype != torch.float32 or x.dim() != 1 or not x.is_contiguous():
        raise RuntimeError('x must be a contiguous 1-D float32 tensor')
    n = x.numel()
    with torch.cuda.device(x.device):
        ws_n = nat.lib.at2_kde_workspace_floats(n, grid_points)     # 0 unless the grid is wider than the accumulator array
        ws = torch.empty(ws_n, device=x.device, dtype=torch.float32) if ws_n > 0 else None
        dens = torch.empty(grid_points, device=x.device, dtype=torch.float32)
        nat.check(nat.lib.at2_kde_f32(x.data_ptr(), n, start, end, grid_points, bandwidth, kernel, dens.data_ptr(),
                                      ws.data_ptr() if ws is not None else None, ws_n, _stream_ptr(x.device)))
    return dens


@torch.library.custom_op('at2::kde', mutates_args=(), device_types='cuda')
def kde(x: torch.Tensor, start: float, end: float, grid_points: int, bandwidth: float, kernel: int) -> torch.Tensor:
    """Kernel density of the float32 samples `x` on linspace(start, end, grid_points) (at2_kde_f32);
    kernel 0 = Epanechnikov (the reference's), 1 = Gaussian."""
    return kde_direct(x, start, end, grid_points, bandwidth, kernel)


def kde_partial_direct(x: torch.Tensor, start: float, end: float, grid_points: int, bandwidth: float, kernel: int,
                parts: torch.Tensor, target_ptrs: List[int], multicast: bool, slot_offset: int) -> None:
    """The sharded form of at2::kde (at2_kde_partial_f32): the UNNORMALISED float64 kernel sums of the samples `x` go to
    elements slot_offset .. slot_offset + grid_points - 1 of every buffer in `target_ptrs` (device addresses of the ranks'
    symmetric buffers, or one NVSwitch multicast address with multicast=True).  `parts` is this rank's own buffer (declared
    mutated; it must be among the targets)."""
    _require_cuda(x, 'x')
    if x.dtype != torch.float32 or x.dim() != 1 or not x.is_contiguous():
        raise RuntimeError('x must be a contiguous 1-D float32 tensor')
    if parts.dtype != torch.float64 or slot_offset < 0 or slot_offset + grid_points > parts.numel():
        raise RuntimeError('parts must be a float64 buffer holding elements slot_offset .. slot_offset + grid_points')
    if not 1 <= len(target_ptrs) <= nat.MAX_PEERS or (multicast and len(target_ptrs) != 1):
        raise RuntimeError(f'between 1 and {nat.MAX_PEERS} targets (exactly one multicast address)')
    n = x.numel()
    with torch.cuda.device(x.device):
        ws_n = nat.lib.at2_kde_workspace_floats(n, grid_points)
        ws = torch.empty(ws_n, device=x.device, dtype=torch.float32) if ws_n > 0 else None
        ptrs = (C.c_void_p * len(target_ptrs))(*target_ptrs)
        nat.check(nat.lib.at2_kde_partial_f32(x.data_ptr(), n, start, end, grid_points, bandwidth, kernel, ptrs, len(target_ptrs),
                                              int(multicast), slot_offset, ws.data_ptr() if ws is not None else None, ws_n,
                                              _stream_ptr(x.device)))


@torch.library.custom_op('at2::kde_partial', mutates_args=('parts',), device_types='cuda')
def kde_partial(x: torch.Tensor, start: float, end: float, grid_points: int, bandwidth: float, kernel: int,
                parts: torch.Tensor, target_ptrs: List[int], multicast: bool, slot_offset: int) -> None:
    """kde_partial_direct as a custom op."""
    kde_partial_direct(x, start, end, grid_points, bandwidth, kernel, parts, target_ptrs, multicast, slot_offset)


def kde_combine_direct(parts: torch.Tensor, n_parts: int, grid_points: int, n_total: int, bandwidth: float, kernel: int) -> torch.Tensor:
    """density = norm * (sum of the first n_parts slots of `parts`, in slot order) (at2_kde_combine_f64)."""
    _require_cuda(parts, 'parts')
    if parts.dtype != torch.float64 or not parts.is_contiguous() or parts.numel() < n_parts * grid_points:
        raise RuntimeError('parts must be a contiguous float64 buffer of at least n_parts * grid_points elements')
    with torch.cuda.device(parts.device):
        dens = torch.empty(grid_points, device=parts.device, dtype=torch.float32)
        nat.check(nat.lib.at2_kde_combine_f64(parts.data_ptr(), n_parts, grid_points, n_total, bandwidth, kernel,
                                              dens.data_ptr(), _stream_ptr(parts.device)))
    return dens


@torch.library.custom_op('at2::kde_combine', mutates_args=(), device_types='cuda')
def kde_combine(parts: torch.Tensor, n_parts: int, grid_points: int, n_total: int, bandwidth: float, kernel: int) -> torch.Tensor:
    """kde_combine_direct as a custom op."""
    return kde_combine_direct(parts, n_parts, grid_points, n_total, bandwidth, kernel)


def _opt_ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else None


@torch.library.custom_op('at2::nn_lookup', mutates_args=(), device_types='cuda')
def nn_lookup(domain_bytes: torch.Tensor, db: torch.Tensor, queries: torch.Tensor, curves: torch.Tensor,
              time: torch.Tensor) -> List[torch.Tensor]:
    """Nearest recorded arm of every query (raw hyperparameter values, domain order) under the reference's normalised
    squared error (at2_nn_lookup_f64).  Returns [idx int32[nq], dist f64[nq], loss f64[nq]]; `loss` is
    curves[idx][time-1] when `curves`/`time` are non-empty, else empty."""
    for name, t in (('db', db), ('queries', queries)):
        _require_cuda(t, name)
        if t.dtype != torch.float64 or t.dim() != 2 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 2-D float64 tensor')
    dom = _struct(nat.DomainDesc, domain_bytes)
    A, D = db.shape
    nq = queries.shape[0]
    if D != dom.n_params or queries.shape[1] != D:
        raise RuntimeError(f'db/queries must have {dom.n_params} columns')
    gather = curves.numel() > 0
    if gather and (curves.dtype != torch.float64 or curves.shape[0] != A or time.dtype != torch.int32 or time.numel() != nq):
        raise RuntimeError('curves must be float64 [A, L] and time int32 [nq]')
    with torch.cuda.device(db.device):
        idx = torch.empty(nq, device=db.device, dtype=torch.int32)
        dist = torch.empty(nq, device=db.device, dtype=torch.float64)
        loss = torch.empty(nq if gather else 0, device=db.device, dtype=torch.float64)
        ws_bytes = nat.lib.at2_nn_workspace_bytes(nq, A, D)
        ws = torch.empty(max(ws_bytes, 16), device=db.device, dtype=torch.uint8)
        nat.check(nat.lib.at2_nn_lookup_f64(C.byref(dom), db.data_ptr(), A, queries.data_ptr(), nq, idx.data_ptr(),
                                            dist.data_ptr(), _opt_ptr(curves) if gather else None,
                                            curves.shape[1] if gather else 0, _opt_ptr(time) if gather else None,
                                            _opt_ptr(loss), ws.data_ptr(), ws_bytes, _stream_ptr(db.device)))
    return [idx, dist, loss]


KF_OUTPUT_NAMES = ['ofe', 'ofe_arm', 'best_arm_vals', 'round_best', 'round_best_arm', 'survivors', 'ckpt_loss', 'nn_idx']


@torch.library.custom_op('at2::hb_knownfn', mutates_args=(), device_types='cuda')
def hb_knownfn(plan_bytes: torch.Tensor, domain_bytes: torch.Tensor, db: torch.Tensor, curves: torch.Tensor,
               descending: bool, n_reps: int, rep_offset: int, seed: int, predrawn: torch.Tensor,
               details: bool) -> List[torch.Tensor]:
    """Hyperband repetitions on the closest-known-function problem (at2_hb_knownfn_f64).  `predrawn` is either empty
    (arms are drawn from Philox streams) or float64 [n_reps, n_arms, D] raw proposed arms."""
    for name, t in (('db', db), ('curves', curves)):
        _require_cuda(t, name)
        if t.dtype != torch.float64 or t.dim() != 2 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 2-D float64 tensor')
    plan, dom = _struct(nat.Plan, plan_bytes), _struct(nat.DomainDesc, domain_bytes)
    A, D = db.shape
    if D != dom.n_params or curves.shape[0] != A:
        raise RuntimeError('db must be [A, D] with D = domain size and curves [A, L]')
    if predrawn.numel() > 0:
        _require_cuda(predrawn, 'predrawn')
        if predrawn.dtype != torch.float64 or tuple(predrawn.shape) != (n_reps, plan.n_arms, D) or not predrawn.is_contiguous():
            raise RuntimeError(f'predrawn must be contiguous float64 [{n_reps}, {plan.n_arms}, {D}]')
    dev = db.device
    with torch.cuda.device(dev):
        o = {'ofe': torch.empty(n_reps, device=dev, dtype=torch.float64),
             'ofe_arm': torch.empty(n_reps, device=dev, dtype=torch.int32),
             'best_arm_vals': torch.empty(n_reps, D, device=dev, dtype=torch.float64)}
        if details:
            o['round_best'] = torch.empty(n_reps, plan.n_rounds, device=dev, dtype=torch.float64)
            o['round_best_arm'] = torch.empty(n_reps, plan.n_rounds, device=dev, dtype=torch.int32)
            o['survivors'] = torch.empty(n_reps, plan.n_surv, device=dev, dtype=torch.int32)
            o['ckpt_loss'] = torch.empty(n_reps, plan.n_arms, plan.n_ckpts, device=dev, dtype=torch.float64)
            o['nn_idx'] = torch.empty(n_reps, plan.n_arms, device=dev, dtype=torch.int32)
        ptr = lambda k: o[k].data_ptr() if k in o else None  # noqa: E731
        outs = nat.HbOutputs(ptr('ofe'), ptr('ofe_arm'), None, ptr('round_best'), ptr('round_best_arm'),
                             ptr('survivors'), ptr('ckpt_loss'))
        if n_reps > 0:   # an empty job is an empty result, as the reference's loop over range(0) would give
            nat.check(nat.lib.at2_hb_knownfn_f64(C.byref(plan), C.byref(dom), db.data_ptr(), A, curves.data_ptr(),
                                                 curves.shape[1], int(descending), n_reps, rep_offset,
                                                 seed & 0xFFFFFFFFFFFFFFFF, _opt_ptr(predrawn), C.byref(outs),
                                                 ptr('nn_idx'), ptr('best_arm_vals'), _stream_ptr(dev)))
    return [o[k] for k in KF_OUTPUT_NAMES if k in o]


def tpe_sim_direct(tables: torch.Tensor, plan, cfg, tpe, n_reps: int, rep_offset: int, seed: int,
                   details: bool) -> List[torch.Tensor]:
    """Hyperband-TPE hybrids / stand-alone TPE / random search on the simulation problem (at2_tpe_sim_f32).
    Returns the at2::hb_sim_f32 outputs followed, when `details`, by arm_xy [n_reps, n_arms, 2], obs_loss [n_reps, n_arms]
    (the loss the TPE phase observed for an arm; NaN for arms that were not observed) and n_injected [n_reps, n_brackets]
    (evaluations of earlier brackets transferred into each bracket's TPE run)."""
    _require_cuda(tables, 'tables')
    with torch.cuda.device(tables.device):
        outs, tensors = _outputs(plan, n_reps, tables.device, torch.float32, details)
        arm_xy = torch.empty(n_reps, plan.n_arms, 2, device=tables.device, dtype=torch.float32) if details else None
        obs = torch.full((n_reps, plan.n_arms), float('nan'), device=tables.device, dtype=torch.float32) if details else None
        inj = torch.zeros(n_reps, plan.n_brackets, device=tables.device, dtype=torch.int32) if details else None
        if n_reps > 0:   # an empty job is an empty result, as the reference's loop over range(0) would give
            nat.check(nat.lib.at2_tpe_sim_f32(C.byref(plan), C.byref(cfg), C.byref(tpe), tables.data_ptr(), n_reps, rep_offset,
                                              seed & 0xFFFFFFFFFFFFFFFF, C.byref(outs),
                                              arm_xy.data_ptr() if details else None, obs.data_ptr() if details else None,
                                              inj.data_ptr() if details else None, _stream_ptr(tables.device)))
    return tensors + ([arm_xy, obs, inj] if details else [])


@torch.library.custom_op('at2::tpe_sim', mutates_args=(), device_types='cuda')
def tpe_sim(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, tpe_bytes: torch.Tensor, n_reps: int,
            rep_offset: int, seed: int, details: bool) -> List[torch.Tensor]:
    """tpe_sim_direct as a custom op (struct arguments as CPU uint8 tensors)."""
    return tpe_sim_direct(tables, _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes),
                          _struct(nat.TpeConfig, tpe_bytes), n_reps, rep_offset, seed, details)


@torch.library.custom_op('at2::tpe_score', mutates_args=(), device_types='cuda')
def tpe_score(obs: torch.Tensor, loss: torch.Tensor, low: float, high: float, tpe_bytes: torch.Tensor,
              candidates: torch.Tensor) -> List[torch.Tensor]:
    """log l(x) - log g(x) of the adaptive Parzen models built from (obs, loss) [n_problems, n_obs] for the given
    candidates [n_problems, n_cand] (at2_tpe_score_f32).  Returns [scores, models[n_problems, 2, cap+1, 3], lean_scores,
    samples[n_problems, 32]] (lean_scores: the candidates scored as a suggestion inside at2::tpe_sim scores them; samples:
    draws from the truncated good model as a suggestion draws its candidates - see include/at2.h)."""
    for name, t in (('obs', obs), ('loss', loss), ('candidates', candidates)):
        _require_cuda(t, name)
        if t.dtype != torch.float32 or t.dim() != 2 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 2-D float32 tensor')
    tpe = _struct(nat.TpeConfig, tpe_bytes)
    n_prob, n_obs = obs.shape
    cap = (n_obs + 31) // 32 * 32
    with torch.cuda.device(obs.device):
        scores = torch.empty_like(candidates)
        models = torch.zeros(n_prob, 2, cap + 1, 3, device=obs.device, dtype=torch.float32)
        lean = torch.empty_like(candidates)
        samples = torch.empty(n_prob, 32, device=obs.device, dtype=torch.float32)
        nat.check(nat.lib.at2_tpe_score_f32(obs.data_ptr(), loss.data_ptr(), n_obs, n_prob, low, high, C.byref(tpe),
                                            candidates.data_ptr(), candidates.shape[1], scores.data_ptr(),
                                            models.data_ptr(), lean.data_ptr(), samples.data_ptr(), _stream_ptr(obs.device)))
    return [scores, models, lean, samples]


@torch.library.custom_op('at2::sim_trajectories', mutates_args=(), device_types='cuda')
def sim_trajectories(tables: torch.Tensor, plan_bytes: torch.Tensor, cfg_bytes: torch.Tensor, f1: torch.Tensor,
                     fn: torch.Tensor, family: torch.Tensor, stream_id: torch.Tensor, seed: int) -> torch.Tensor:
    """Raw trajectories [n_arms, R] of the given arms (at2_sim_trajectories_f32)."""
    for name, t, dt in (('tables', tables, torch.float32), ('f1', f1, torch.float32), ('fn', fn, torch.float32),
                        ('family', family, torch.int32), ('stream_id', stream_id, torch.int64)):
        _require_cuda(t, name)
        if t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f'{name} must be contiguous {dt}')
    plan, cfg = _struct(nat.Plan, plan_bytes), _struct(nat.SimConfig, cfg_bytes)
    n = f1.numel()
    if fn.numel() != n or family.numel() != n or stream_id.numel() != n:
        raise RuntimeError('f1, fn, family and stream_id must have the same length')
    with torch.cuda.device(f1.device):
        raw = torch.empty(n, plan.R, device=f1.device, dtype=torch.float32)
        nat.check(nat.lib.at2_sim_trajectories_f32(C.byref(plan), C.byref(cfg), tables.data_ptr(), n, f1.data_ptr(),
                                                   fn.data_ptr(), family.data_ptr(), stream_id.data_ptr(),
                                                   seed & 0xFFFFFFFFFFFFFFFF, raw.data_ptr(), _stream_ptr(f1.device)))
    return raw


@torch.library.custom_op('at2::nn_lookup_excluding', mutates_args=(), device_types='cuda')
def nn_lookup_excluding(domain_bytes: torch.Tensor, db: torch.Tensor, queries: torch.Tensor,
                        exclude: torch.Tensor) -> List[torch.Tensor]:
    """Nearest recorded arm with a per-query held-out index range exclude[i] = (lo, hi) (at2_nn_lookup_excluding_f64)."""
    for name, t in (('db', db), ('queries', queries)):
        _require_cuda(t, name)
        if t.dtype != torch.float64 or t.dim() != 2 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 2-D float64 tensor')
    _require_cuda(exclude, 'exclude')
    dom = _struct(nat.DomainDesc, domain_bytes)
    A, D = db.shape
    nq = queries.shape[0]
    if exclude.dtype != torch.int32 or tuple(exclude.shape) != (nq, 2) or not exclude.is_contiguous():
        raise RuntimeError('exclude must be contiguous int32 [nq, 2]')
    with torch.cuda.device(db.device):
        idx = torch.empty(nq, device=db.device, dtype=torch.int32)
        dist = torch.empty(nq, device=db.device, dtype=torch.float64)
        ws = torch.empty(nq * 12 + 16, device=db.device, dtype=torch.uint8)
        nat.check(nat.lib.at2_nn_lookup_excluding_f64(C.byref(dom), db.data_ptr(), A, queries.data_ptr(), nq,
                                                      exclude.data_ptr(), idx.data_ptr(), dist.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), _stream_ptr(db.device)))
    return [idx, dist]


@torch.library.custom_op('at2::order_profile', mutates_args=(), device_types='cuda')
def order_profile(curves: torch.Tensor) -> List[torch.Tensor]:
    """[dynamic order profile [T-1], ends order profile [1]] of loss curves [n, T] (at2_order_profile_f32)."""
    _require_cuda(curves, 'curves')
    if curves.dtype != torch.float32 or curves.dim() != 2 or not curves.is_contiguous():
        raise RuntimeError('curves must be a contiguous 2-D float32 tensor')
    n, T = curves.shape
    with torch.cuda.device(curves.device):
        dyn = torch.empty(T - 1, device=curves.device, dtype=torch.float32)
        ends = torch.empty(1, device=curves.device, dtype=torch.float32)
        nat.check(nat.lib.at2_order_profile_f32(curves.data_ptr(), n, T, dyn.data_ptr(), ends.data_ptr(),
                                                _stream_ptr(curves.device)))
    return [dyn, ends]


@torch.library.custom_op('at2::ks_statistic', mutates_args=(), device_types='cuda')
def ks_statistic(a_sorted: torch.Tensor, b_sorted: torch.Tensor) -> torch.Tensor:
    """Two-sample KS statistic D of two sorted float32 samples (at2_ks_statistic_f32)."""
    for name, t in (('a_sorted', a_sorted), ('b_sorted', b_sorted)):
        _require_cuda(t, name)
        if t.dtype != torch.float32 or t.dim() != 1 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 1-D float32 tensor')
    with torch.cuda.device(a_sorted.device):
        d = torch.zeros(1, device=a_sorted.device, dtype=torch.float32)
        nat.check(nat.lib.at2_ks_statistic_f32(a_sorted.data_ptr(), a_sorted.numel(), b_sorted.data_ptr(),
                                               b_sorted.numel(), d.data_ptr(), _stream_ptr(a_sorted.device)))
    return d


@torch.library.custom_op('at2::segmented_topk', mutates_args=(), device_types='cuda')
def segmented_topk(losses: torch.Tensor, seg_offsets: torch.Tensor, keep: torch.Tensor, max_seg_len: int,
                   descending: bool) -> torch.Tensor:
    """The successive-halving selection alone (at2_segmented_topk_f32 / _f64): per segment of `losses` (float32 or
    float64, 1-D; segment s = [seg_offsets[s], seg_offsets[s+1]), int64) the positions of its keep[s] (int32) best losses,
    best first, in python's stable sorted() order (hyperband_optimiser.py:46-57); -1 past keep.  int32, laid out like `losses`."""
    _require_cuda(losses, 'losses')
    for name, t, dt in (('seg_offsets', seg_offsets, torch.int64), ('keep', keep, torch.int32)):
        _require_cuda(t, name)
        if t.dtype != dt or t.dim() != 1 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous 1-D {dt} tensor')
    if losses.dtype not in (torch.float32, torch.float64) or losses.dim() != 1 or not losses.is_contiguous():
        raise RuntimeError('losses must be a contiguous 1-D float32 or float64 tensor')
    if seg_offsets.numel() != keep.numel() + 1:
        raise RuntimeError('seg_offsets must have one more entry than keep')
    fn = nat.lib.at2_segmented_topk_f32 if losses.dtype == torch.float32 else nat.lib.at2_segmented_topk_f64
    with torch.cuda.device(losses.device):
        idx = torch.full((losses.numel(),), -1, device=losses.device, dtype=torch.int32)
        nat.check(fn(losses.data_ptr(), seg_offsets.data_ptr(), keep.data_ptr(), keep.numel(), int(max_seg_len),
                     int(bool(descending)), idx.data_ptr(), _stream_ptr(losses.device)))
    return idx
