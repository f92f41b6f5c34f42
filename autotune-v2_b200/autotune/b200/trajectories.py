"""Full loss trajectories on the device: the per-evaluator API (`SimulationEvaluator.simulate` / `.fs`) and the input
of the loss-curve family profiles.  Not the hot path - the fused kernels never store trajectories."""
import ctypes as C
import functools

import numpy as np
import torch

from . import _native as nat
from .engine import HyperbandSimSpec, _cuda_device

EVALUATOR_SEED = 0x5EED0A72


def savgol_matrix(R):
    """The R x R operator W with savgol_filter(x, window(R), 3) == W @ x (host helper at2_savgol_rows)."""
    rows = np.arange(R, dtype=np.int32)
    out = np.zeros((R, R))
    rc = nat.lib.at2_savgol_rows(int(R), int(R), rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise ValueError(nat.lib.at2_last_error().decode())
    return out


@functools.lru_cache(maxsize=64)
def _savgol_device(R, device):
    return torch.from_numpy(savgol_matrix(R)).to(device)


def smooth(raw):
    """Savitzky-Golay read-out of raw trajectories [n, R] (core/simulation_evaluator.py:76-82), float64 on the device."""
    W = _savgol_device(raw.shape[-1], str(raw.device))
    return raw.double() @ W.T


def simulate_arms(spec, f1, fn, family, stream_id, seed=0, device=None):
    """Raw trajectories [n, R] of arms with first value f1, target fn, family index and Philox stream id."""
    device = _cuda_device(device)

    def dev(a, dt):
        return torch.as_tensor(np.asarray(a)).to(device=device, dtype=dt).contiguous()
    return torch.ops.at2.sim_trajectories(spec.tables(device, torch.float32), spec.plan_bytes, spec.cfg_bytes,
                                          dev(f1, torch.float32), dev(fn, torch.float32), dev(family, torch.int32),
                                          dev(stream_id, torch.int64), int(seed))


@functools.lru_cache(maxsize=64)
def _single_family_spec(n, ml, nec, up):
    # the test function and shifts are irrelevant here: f_1 and f_n are given explicitly
    return HyperbandSimSpec('branin', [(ml, nec, up, False, 0, 0)], max_iter=n, eta=max(2, n), init_noise=0)


def simulate_one(f_1, f_n, n, ml_aggressiveness, necessary_aggressiveness, up_spikiness, stream_id, draw):
    """One evaluator's raw trajectory as a python list of n floats (SimulationEvaluator.simulate(n, n, f_n))."""
    if necessary_aggressiveness == np.inf:
        raise NotImplementedError('necessary_aggressiveness = inf is not supported on the device path')
    spec = _single_family_spec(int(n), float(ml_aggressiveness), float(necessary_aggressiveness), float(up_spikiness))
    raw = simulate_arms(spec, [f_1], [f_n], [0], [(int(stream_id) << 20) + int(draw)], seed=EVALUATOR_SEED)
    return [float(v) for v in raw[0].cpu()]


def smooth_one(non_smooth_fs, max_resources):
    window = nat.lib.at2_savgol_window(int(max_resources))
    if window > len(non_smooth_fs):
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    raw = torch.tensor([non_smooth_fs], dtype=torch.float64, device=_cuda_device(None))
    # the window is set by max_resources, the operator by the actual trajectory length (they coincide in practice)
    return [float(v) for v in smooth(raw)[0].cpu()] if len(non_smooth_fs) == max_resources else \
        _smooth_general(non_smooth_fs, window)


def _smooth_general(values, window):
    raise NotImplementedError('smoothing a trajectory whose length differs from max_resources is not supported')
