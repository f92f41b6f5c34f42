"""EPDF-OFE jobs: Monte-Carlo repetitions sharded over the GPUs of one box, one all-gather of the OFE vector, then the
kernel density estimate (reference experiments/run_simulation.py:144-182 + simulation_evaluation/plot_results.py:11-17).

One process per GPU (torch.distributed, NCCL over NVLink); repetitions are independent, so the only exchange is the
all-gather of each rank's OFE slice.  RNG streams are keyed by the *global* repetition id, so the gathered vector is
bit-identical for every world size.
"""
import numpy as np
import torch
import torch.distributed as dist

KERNEL_IDS = {'epanechnikov': 0, 'gaussian': 1}


def shard_range(n_total, rank, world):
    """[lo, hi) of the repetitions rank `rank` owns; the first n_total % world ranks take one extra."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_slices(local, n_total, group=None):
    """All-gather the per-rank slices (in shard_range order) into the full [n_total] vector on every rank.
    Works on CUDA tensors (NCCL) and CPU tensors (gloo); slices are padded to a common length for the collective."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.numel() != n_total:
            raise ValueError(f'single process holds {local.numel()} of {n_total} values')
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_total, rank, world)
    if local.numel() != hi - lo:
        raise ValueError(f'rank {rank} holds {local.numel()} values, expected {hi - lo}')
    width = -(-n_total // world)
    send = local if local.numel() == width else torch.cat([local, local.new_zeros(width - local.numel())])
    recv = local.new_empty(world * width)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if n_total % world == 0:
        return recv
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        parts.append(recv[r * width:r * width + (b - a)])
    return torch.cat(parts)


def density(ofes, start, end, bandwidth, grid_points=1000, kernel='epanechnikov'):
    """EPDF of a float32 CUDA vector on linspace(start, end, grid_points)."""
    return torch.ops.at2.kde(ofes.contiguous(), float(start), float(end), int(grid_points), float(bandwidth),
                             KERNEL_IDS[kernel])


def summary_stats(ofes):
    """mean, sample std (ddof=1), best quartile and its mean - definitions of run_simulation.py:163-181, on device."""
    n = ofes.numel()
    srt = torch.sort(ofes.double()).values
    q = srt[:int(n / 4)]
    return {'avg_optimum': ofes.double().mean().item(),
            'sample_std_dev': ofes.double().std(unbiased=True).item() if n > 1 else float('nan'),
            'avg_top_25%': q.mean().item() if q.numel() else float('nan'), 'top_quartile': q}


def run_epdf_ofe(spec, n_reps, seed=0, *, start, end, bandwidth, grid_points=1000, kernel='epanechnikov',
                 group=None, device=None, runner=None):
    """The whole job: this rank's share of `n_reps` Hyperband repetitions -> all-gather -> KDE.

    Returns {'ofes': [n_reps] float32 (all ranks' OFEs, in repetition order), 'density': [grid_points], 'grid': ...}.
    `runner(spec, n, seed, rep_offset, device)` defaults to the fused Hyperband kernel; other optimisers plug in here.
    """
    from .engine import _cuda_device, run_hyperband_repetitions
    device = _cuda_device(device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(n_reps, rank, world)
    if runner is None:
        def runner(spec_, n, seed_, off, dev):
            return run_hyperband_repetitions(spec_, n, seed=seed_, rep_offset=off, device=dev)['ofe']
    local = runner(spec, hi - lo, seed, lo, device) if hi > lo else torch.empty(0, device=device)
    ofes = gather_slices(local, n_reps, group)
    dens = density(ofes, start, end, bandwidth, grid_points, kernel)
    return {'ofes': ofes, 'density': dens, 'grid': np.linspace(start, end, grid_points)}
