"""EPDF-OFE jobs: Monte-Carlo repetitions sharded over the GPUs of one box, one all-gather of the OFE vector, then the
kernel density estimate (reference experiments/run_simulation.py:144-182 + simulation_evaluation/plot_results.py:11-17).

One process per GPU (torch.distributed, NCCL over NVLink); repetitions are independent, so the only exchange is the
all-gather of each rank's OFE slice.  RNG streams are keyed by the *global* repetition id, so the gathered vector is
bit-identical for every world size.
"""
import functools

import numpy as np
import torch
import torch.distributed as dist

from . import ops

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


@functools.lru_cache(maxsize=64)
def _grid(start, end, grid_points):
    g = np.linspace(start, end, grid_points)
    g.setflags(write=False)
    return g


class FusedGather:
    """The all-gather of the OFE vector fused into the simulation kernel (at2_hb_sim_gather_f32): every rank's kernel stores
    its OFEs straight into all ranks' buffers over NVLink / NVSwitch while it is still simulating - one store per repetition
    to the NVSwitch multicast address when the fabric offers one, else one store per peer - and a barrier across the ranks
    replaces the NCCL collective.  Buffers are torch symmetric memory; two of them alternate so that one barrier per job is
    enough (a rank can run at most one job ahead of the slowest reader).

    Built by the collective `FusedGather.create`, which returns None (on every rank alike) when symmetric memory cannot be
    set up; callers then fall back to `gather_slices` (NCCL)."""

    MAX_GRID = 4096      # grid points a sharded KDE can exchange (the reference uses 1000)

    def __init__(self, n_total, device, group, bufs, handles, use_multicast, kbufs=None, khandles=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_total, self.device = int(n_total), torch.device(device)
        self.bufs, self.handles, self.targets = bufs, handles, []

        def targets_of(hdl):
            mc = int(getattr(hdl, 'multicast_ptr', 0) or 0) if use_multicast else 0
            return [mc] if mc else [int(p) for p in hdl.buffer_ptrs]
        self.targets = [targets_of(h) for h in handles]
        self.multicast = len(self.targets[0]) == 1 and self.world > 1
        # float64 slots [world][MAX_GRID] for the ranks' partial kernel sums of the sharded KDE (two alternate, like bufs)
        self.kbufs, self.khandles = kbufs, khandles
        self.ktargets = [targets_of(h) for h in khandles] if khandles else None
        self.turn = 0

    @classmethod
    def create(cls, n_total, device, group=None, use_multicast=True):
        """A FusedGather, or None - on every rank alike - when symmetric memory cannot be set up (callers then use NCCL).
        Collective: every rank of the group must call it.  The local allocation and the collective rendezvous are agreed on
        separately, so that a rank whose allocation failed never leaves the others waiting inside the rendezvous."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        group = group if group is not None else dist.group.WORLD

        def all_ok(ok):
            flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1
        bufs = handles = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            bufs = [symm_mem.empty(int(n_total), dtype=torch.float32, device=torch.device(device)) for _ in range(2)]
        except Exception:  # noqa: BLE001
            bufs = None
        if not all_ok(bufs is not None):
            return None
        kbufs = khandles = None
        try:
            handles = [symm_mem.rendezvous(b, group) for b in bufs]
            ok = all(len(h.buffer_ptrs) == dist.get_world_size(group) for h in handles)
        except Exception:  # noqa: BLE001
            ok = False
        if not all_ok(ok):
            return None
        try:   # optional: the slots of the sharded KDE; without them the KDE runs over the gathered vector on every rank
            kbufs = [symm_mem.empty(dist.get_world_size(group) * cls.MAX_GRID, dtype=torch.float64, device=torch.device(device))
                     for _ in range(2)]
            khandles = [symm_mem.rendezvous(b, group) for b in kbufs]
            kok = all(len(h.buffer_ptrs) == dist.get_world_size(group) for h in khandles)
        except Exception:  # noqa: BLE001
            kok = False
        if not all_ok(kok):
            kbufs = khandles = None
        return cls(n_total, device, group, bufs, handles, use_multicast, kbufs, khandles)

    def run(self, spec, seed, early_stop=True):
        """This rank's share of the job's repetitions; returns the gathered [n_total] OFE vector (a view of a symmetric
        buffer that stays valid until the job after the next one).  early_stop: the early-stopping kernel (the product's
        default; same OFEs bit for bit) or the full simulation."""
        lo, hi = shard_range(self.n_total, self.rank, self.world)
        buf, hdl, targets = self.bufs[self.turn], self.handles[self.turn], self.targets[self.turn]
        self.turn ^= 1
        ops.hb_sim_gather_direct(spec.tables(self.device, torch.float32), spec.plan, spec.cfg, hi - lo, lo, int(seed), buf,
                                 targets, lo, self.multicast, bool(early_stop))
        hdl.barrier(channel=0)          # on the current stream: every rank's stores are in every buffer after it
        return buf

    def run_epdf(self, spec, seed, start, end, bandwidth, grid_points=1000, kernel='epanechnikov', early_stop=True,
                 sim_events=None):
        """The whole EPDF-OFE step with the KDE sharded as well: this rank's repetitions (OFEs scattered into every rank's
        buffer by the kernel) -> kernel sums of its OWN slice on the grid (float64, scattered into every rank's slot array the
        same way) -> ONE barrier across the ranks -> the ranks' sums added in rank order and normalised.  Every rank ends up
        with the same gathered vector and the same density bits, and evaluates only 1/world of the N x G pair terms.
        Returns (ofes [n_total], density [grid_points]).  sim_events: an optional pair of CUDA events recorded around the
        simulation launch alone (measurements)."""
        stream = torch.cuda.current_stream(self.device)
        if self.ktargets is None or grid_points > self.MAX_GRID:
            if sim_events is not None:
                sim_events[0].record(stream)
            ofes = self.run(spec, seed, early_stop)
            if sim_events is not None:
                sim_events[1].record(stream)
            return ofes, density(ofes, start, end, bandwidth, grid_points, kernel)
        lo, hi = shard_range(self.n_total, self.rank, self.world)
        turn = self.turn
        buf, hdl, targets = self.bufs[turn], self.handles[turn], self.targets[turn]
        kbuf, ktargets = self.kbufs[turn], self.ktargets[turn]
        self.turn ^= 1
        if sim_events is not None:
            sim_events[0].record(stream)
        local = ops.hb_sim_gather_direct(spec.tables(self.device, torch.float32), spec.plan, spec.cfg, hi - lo, lo, int(seed),
                                         buf, targets, lo, self.multicast, bool(early_stop))[0]
        if sim_events is not None:
            sim_events[1].record(stream)
        # the kernel sums are taken over the launch's own (locally written) OFE output, not over the symmetric buffer
        ops.kde_partial_direct(local, float(start), float(end), int(grid_points), float(bandwidth), KERNEL_IDS[kernel], kbuf,
                               ktargets, self.multicast, self.rank * int(grid_points))
        hdl.barrier(channel=0)          # covers both scatters: they precede it in stream order on every rank
        dens = ops.kde_combine_direct(kbuf, self.world, int(grid_points), self.n_total, float(bandwidth), KERNEL_IDS[kernel])
        return buf, dens


def density(ofes, start, end, bandwidth, grid_points=1000, kernel='epanechnikov'):
    """EPDF of a float32 CUDA vector on linspace(start, end, grid_points)."""
    return ops.kde_direct(ofes.contiguous(), float(start), float(end), int(grid_points), float(bandwidth), KERNEL_IDS[kernel])


def summary_stats(ofes):
    """mean, sample std (ddof=1), best quartile and its mean - definitions of run_simulation.py:163-181, on device."""
    n = ofes.numel()
    srt = torch.sort(ofes.double()).values
    q = srt[:int(n / 4)]
    return {'avg_optimum': ofes.double().mean().item(),
            'sample_std_dev': ofes.double().std(unbiased=True).item() if n > 1 else float('nan'),
            'avg_top_25%': q.mean().item() if q.numel() else float('nan'), 'top_quartile': q}


def run_epdf_ofe(spec, n_reps, seed=0, *, start, end, bandwidth, grid_points=1000, kernel='epanechnikov',
                 group=None, device=None, runner=None, fused=None, early_stop=True):
    """The whole job: this rank's share of `n_reps` Hyperband repetitions -> all-gather -> KDE.

    Returns {'ofes': [n_reps] float32 (all ranks' OFEs, in repetition order), 'density': [grid_points], 'grid': ...}.
    `runner(spec, n, seed, rep_offset, device)` defaults to the fused Hyperband kernel; other optimisers plug in here.
    `fused`: a FusedGather for (n_reps, device, group) - the kernel then scatters its OFEs into every rank's buffer itself
    and no NCCL collective runs (plain Hyperband only); the gathered vector is bit-identical either way.
    `early_stop` (default True): the early-stopping kernel; False simulates every arm to R like the reference - same OFEs.
    """
    from .engine import _cuda_device, run_hyperband_repetitions
    device = _cuda_device(device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if fused is not None and runner is None:
        if fused.n_total != int(n_reps):
            raise ValueError(f'FusedGather was built for {fused.n_total} repetitions, not {n_reps}')
        ofes, dens = fused.run_epdf(spec, seed, start, end, bandwidth, grid_points, kernel, early_stop)
        return {'ofes': ofes, 'density': dens, 'grid': _grid(start, end, grid_points)}
    lo, hi = shard_range(n_reps, rank, world)
    if runner is None:
        def runner(spec_, n, seed_, off, dev):
            return run_hyperband_repetitions(spec_, n, seed=seed_, rep_offset=off, device=dev, early_stop=early_stop)['ofe']
    local = runner(spec, hi - lo, seed, lo, device) if hi > lo else torch.empty(0, device=device)
    ofes = gather_slices(local, n_reps, group)
    dens = density(ofes, start, end, bandwidth, grid_points, kernel)
    return {'ofes': ofes, 'density': dens, 'grid': _grid(start, end, grid_points)}
