#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 hot path (contract: see the task statement / DESIGN.md "Measurement").

Workload (north_star Target; BASELINE.json configs[0]/[1] shape): Hyperband(R=81, eta=3) EPDF-OFE on the six-family Rastrigin
Gamma-simulation problem (the reference's own driver configuration, experiments/run_simulation.py:22-47: uniform scheduler,
init_noise 10), 10^5 repetitions per GPU per step (weak scaling: repetitions are independent and shard with no data-path
collective; the one exchange is the all-gather of the OFE vector for the KDE).  A step = fused simulation + successive
halving kernel over the rank's 10^5 repetitions -> all-gather (N > 1, fused into the kernel) -> Epanechnikov KDE on the
1000-point grid (plot_run_simulations_results.py:27-54, rastrigin-families-2).

`value` times the FULL simulation (every arm simulated to R like the reference does - the roofline headline);
`value_product_default` times the same job through the product's default path (early stopping: same OFEs bit for bit).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line on rank 0)
  python bench.py --impl reference ...                           the reference's own classes on the box's host cores
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'autotune-v2_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]   # run_simulation.py:41-47
EGG3 = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]      # run_simulation.py:36-40
FLAT = [(0, 1, 0, False, 0, 0)]
FUNC, FAMILIES, R, ETA, NOISE = 'rastrigin', FAM6, 81, 3, 10.0
REPS_PER_GPU = 100000
KDE_START, KDE_END, KDE_H, KDE_G = -400.0, -370.0, 3.0, 1000   # plot_run_simulations_results.py:27-54 (rastrigin-families-2)
ARMS, ARM_EPOCHS = 117, 117 * 80                              # per repetition
FLOP_PER_ARM_EPOCH = 32                                       # algorithmic FP32 flop per arm-epoch
WORKLOAD = ('Hyperband(R=81,eta=3) EPDF-OFE, six-family Rastrigin Gamma-simulation (run_simulation.py:41-47, uniform '
            'scheduler, init_noise 10), 1e5 repetitions per GPU per step (north_star Target; BASELINE.json configs[0]/[1])')
METRIC, UNIT = 'simulated optimiser runs/sec', 'runs/s'
_REAL_STDOUT = None


def emit(line):
    """The one JSON line of this process, on the real stdout."""
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--reps-per-gpu', type=int, default=REPS_PER_GPU)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the other configurations / sweeps (profiling runs)')
    ap.add_argument('--nccl-gather', action='store_true', help='N > 1: NCCL all-gather instead of the fused gather')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
def cpu_pool():
    """(pool, kind, cores): the UNMODIFIED reference from baseline/_ref (kind 'reference') when it travelled with the
    snapshot, else the oracle's reference-shaped port (kind 'port').  Fork pools: create before CUDA is initialised."""
    from oracle import ref_runner
    if ref_runner.available():
        pool = ref_runner.ReferencePool()
        return pool, 'reference', pool.procs
    from oracle.cpu_bench import HyperbandCpuPool
    pool = HyperbandCpuPool()
    return pool, 'port', pool.procs


def run_reference_arm(args, rank):
    """The reference's own CPU implementation of the path - HyperbandOptimiser + OptFunctionSimulationProblem driven as in
    experiments/run_simulation.py:147-157, one process per host core - each step a bounded sample of the same workload."""
    if rank != 0:
        return
    pool, kind, procs = cpu_pool()
    _, w, _ = pool.run(FUNC, FAMILIES, R, ETA, NOISE, 1, seed0=77)          # untimed: imports, first-call costs
    _, w, _ = pool.run(FUNC, FAMILIES, R, ETA, NOISE, 1, seed0=78)
    # size a step so that the whole run takes about a minute
    reps_per_proc = max(1, min(40, int(60.0 / max(args.steps + min(args.warmup, 2), 1) / max(w, 1e-3))))
    for i in range(min(args.warmup, 2)):
        pool.run(FUNC, FAMILIES, R, ETA, NOISE, reps_per_proc, seed0=500 + 131 * i)
    t0 = time.perf_counter()
    n = 0
    for s in range(args.steps):
        _, _, ofes = pool.run(FUNC, FAMILIES, R, ETA, NOISE, reps_per_proc, seed0=2000 + 131 * s)
        n += len(ofes)
    wall = time.perf_counter() - t0
    pool.close()
    value = n / wall
    sample = f'{procs} processes x {reps_per_proc} repetitions per step ({procs * reps_per_proc} runs/step)'
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': wall / max(args.steps, 1) * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'arm_epochs_per_s': value * ARM_EPOCHS,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': kind, 'sample': sample,
                         'what': ("the unmodified reference (baseline/_ref/autotune): HyperbandOptimiser.run_optimisation on "
                                  "OptFunctionSimulationProblem in the loop of experiments/run_simulation.py:147-157, "
                                  "multiprocessing fork pool") if kind == 'reference' else
                                 'oracle port of the same loop (baseline/_ref absent on this box)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def cpu_baseline_prepare():
    """Fork the CPU-baseline worker pools (they must exist before CUDA is initialised) and let the workers import their
    modules; the work itself runs AFTER the GPU measurements (cpu_baseline_measure): 15 s of all-core load right before the
    timed GPU steps was visible in the wall-clock e2e figure (-5 %)."""
    pool, kind, procs = cpu_pool()
    pool.run(FUNC, FAMILIES, R, ETA, NOISE, 1, seed0=11)
    port = None
    try:
        from oracle.cpu_bench import HyperbandCpuPool
        port = HyperbandCpuPool(procs)
        port.run(FUNC, FAMILIES, R, ETA, NOISE, 1)
    except Exception:  # noqa: BLE001
        port = None
    return pool, kind, procs, port


def cpu_baseline_measure(ctx):
    """cpu_baseline object of our arm's line (rank 0, N = 1): ~15 s of the reference on all host cores, plus - as extras,
    so that the GPU/CPU ratio is not overstated - the oracle port, batched NumPy on the same cores and the reference's
    sklearn KDE call."""
    pool, kind, procs, port = ctx
    _, w, _ = pool.run(FUNC, FAMILIES, R, ETA, NOISE, 1, seed0=12)
    reps_per_proc = max(2, min(400, int(15.0 / max(w, 1e-3))))
    v, wall, _ = pool.run(FUNC, FAMILIES, R, ETA, NOISE, reps_per_proc)
    pool.close()
    out = {'value': v, 'unit': UNIT, 'cores': procs, 'kind': kind,
           'sample': f'{procs} processes x {reps_per_proc} repetitions of the same workload, {wall:.1f} s wall'}
    try:
        if port is None:
            raise RuntimeError('oracle port pool unavailable')
        pv, pw, _ = port.run(FUNC, FAMILIES, R, ETA, NOISE, 40)
        out['oracle_port'] = {'value': pv, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                              'sample': f'{procs} processes x 40 repetitions, {pw:.1f} s wall'}
        from oracle import vectorised as ov
        vec_reps = 2000
        jobs = [(FUNC, FAMILIES, R, ETA, NOISE, vec_reps, 500, 7000 + w_) for w_ in range(procs)]
        t0 = time.perf_counter()
        port.pool.map(ov._worker, jobs) if port.pool else [ov._worker(jobs[0])]
        vwall = time.perf_counter() - t0
        out['vectorised_numpy'] = {
            'value': procs * vec_reps / vwall, 'unit': UNIT, 'cores': procs, 'kind': "port (batched NumPy, not the "
            "reference's code shape)", 'sample': f'{procs} processes x {vec_reps} repetitions in blocks of 500, {vwall:.1f} s wall'}
        port.close()
    except Exception as e:  # noqa: BLE001
        out['oracle_port'] = {'unavailable': repr(e)}
    try:   # the reference's own KDE call (plot_results.py:13-15: sklearn KernelDensity, epanechnikov, 1000-point grid)
        import numpy as np
        from sklearn.neighbors import KernelDensity
        xs = np.random.default_rng(0).normal(-385.0, 4.0, REPS_PER_GPU)[:, np.newaxis]
        t0 = time.perf_counter()
        kde = KernelDensity(kernel='epanechnikov', bandwidth=KDE_H).fit(xs)
        np.exp(kde.score_samples(np.linspace(KDE_START, KDE_END, KDE_G)[:, np.newaxis]))
        out['kde_sklearn'] = {'seconds': time.perf_counter() - t0, 'samples': REPS_PER_GPU, 'grid': KDE_G, 'cores': 1,
                              'kind': 'reference call (sklearn.neighbors.KernelDensity)'}
    except Exception as e:  # noqa: BLE001
        out['kde_sklearn'] = {'unavailable': repr(e)}
    return out


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self._stop_evt = index, period, [], threading.Event()
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        if not self.ok:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable: ' + getattr(self, 'err', '?')]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        names = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
                 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
                 0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}
        reasons = [n for b, n in names.items() if bits & b and n != 'gpu_idle']
        return {'sm_mhz': mhz[len(mhz) // 2] if mhz else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(inside)}


def measured_traffic(key='dram_bytes_per_launch_avg'):
    """DRAM bytes (or SASS-counted FP32 flop) per launch of the headline simulation kernel from the committed ncu captures
    (profiles/), newest round first, or None."""
    for tag in ('r02', 'r01'):
        try:
            with open(os.path.join(ROOT, 'profiles', f'{tag}_dram_traffic.json')) as f:
                d = json.load(f)
            for k, v in d.items():
                if k.startswith('hb_sim_kernel<float, 0, 1, 5') or k.startswith('hb_sim_kernel<float, 0, 1, 6') or (tag == 'r01' and k.startswith('hb_sim_kernel<float, 0, 0, 1')):
                    if v.get(key) is not None:
                        return v[key], f'profiles/{tag}_dram_traffic.json [{k}]'
        except Exception:  # noqa: BLE001
            pass
    return None, None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured (MEASURED_PEAKS.json)'
    except Exception:  # noqa: BLE001
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback (B200_PROFILING.md)'


def main():
    # stdout carries exactly ONE JSON line.  Libraries write banners to file descriptor 1 behind python's back (NCCL prints
    # its version there when the first communicator comes up), so fd 1 is pointed at stderr for the whole run and the JSON
    # line is written to the saved descriptor at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    # CPU baseline: the fork pools must exist before CUDA is initialised, the work runs after the GPU measurements; rank 0, N=1 only
    cpu_ctx = cpu_baseline_prepare() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    import torch
    import torch.distributed as dist
    from autotune.b200 import _native as nat
    from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
    from autotune.b200.epdf import FusedGather, density, gather_slices, run_epdf_ofe

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_local = args.reps_per_gpu
    n_total = n_local * world
    spec = HyperbandSimSpec(FUNC, FAMILIES, R, ETA, NOISE, 'uniform', 'min')
    stream = torch.cuda.current_stream(dev)
    K = max(args.steps, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- FP32 FMA peak, measured in this run (roofline denominator of the simulation kernel) ----
    sink = torch.zeros(1, device=dev)
    blocks, threads, iters = 148 * 8, 256, 20000
    best_ms, by_variant = 1e30, {}
    for variant in (0, 1, 2):
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            torch.ops.at2.fma_peak_probe(sink, blocks, threads, iters, variant)
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b)
            by_variant[variant] = min(by_variant.get(variant, 1e30), ms)
            best_ms = min(best_ms, ms)
    fma_peak_tflops = 2.0 * 16 * iters * blocks * threads / (best_ms * 1e-3) / 1e12
    fma_variants = {str(k): 2.0 * 16 * iters * blocks * threads / (v * 1e-3) / 1e12 for k, v in by_variant.items()}

    # N > 1: the all-gather of the OFEs is fused into the simulation kernel (peer / NVSwitch-multicast stores + one barrier);
    # NCCL all-gather if symmetric memory cannot be set up (decided collectively) or with --nccl-gather
    fused = None if (world == 1 or args.nccl_gather) else FusedGather.create(n_total, dev)
    gather_kind = 'none (one GPU)' if world == 1 else (
        'NCCL all-gather' if fused is None else
        'fused into the simulation kernel (' + ('multimem.st to the NVSwitch multicast address' if fused.multicast
                                               else 'peer stores over NVLink')
        + ' + one barrier, torch symmetric memory)')

    def make_step(spec_, n_loc, n_tot, fg, early_stop, kde=(KDE_START, KDE_END, KDE_H, KDE_G)):
        def step(seed, sim_events=None):
            if fg is not None:
                # this rank's repetitions with the OFEs scattered to every rank by the kernel, the KDE of its own slice,
                # one barrier, the ranks' kernel sums added in rank order
                return fg.run_epdf(spec_, seed, kde[0], kde[1], kde[2], kde[3], 'epanechnikov', early_stop, sim_events)
            if sim_events is not None:
                sim_events[0].record(stream)
            ofe = run_hyperband_repetitions(spec_, n_loc, seed=seed, rep_offset=rank * n_loc, device=dev,
                                            early_stop=early_stop)['ofe']
            if sim_events is not None:
                sim_events[1].record(stream)
            full = gather_slices(ofe, n_tot) if world > 1 else ofe
            return full, density(full, *kde[:3], kde[3])
        return step

    def time_steps(step, n_tot, steps, warm=3, with_sim_events=False):
        """(runs/s over all ranks, ms per step, kernel ms): W warm-up steps, then `steps` steps between barriers, CUDA events on
        the launching stream, max over ranks."""
        for i in range(warm):
            step(1000 + i)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] \
            if with_sim_events else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(steps):
            step(i, ev[i] if ev else None)
        e1.record(stream)
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        sim_ms = sum(a.elapsed_time(b) for a, b in ev) / steps if ev else None
        return n_tot * steps / (ms_total * 1e-3), ms_total / steps, sim_ms

    step_full = make_step(spec, n_local, n_total, fused, False)
    step_default = make_step(spec, n_local, n_total, fused, True)

    # bring the SM clock up before the W warm-up steps: a cold B200 needs a few hundred ms of load to leave its idle
    # clock, far more than W steps of 3 ms (untimed, not counted as steps)
    for i in range(100):
        step_full(2000 + i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    value, ms_per_step, sim_ms = time_steps(step_full, n_total, K, warm=max(args.warmup, 3), with_sim_events=True)
    t1 = time.perf_counter()
    sampler.stop()
    # per step: fused simulation + halving, KDE (one launch: the last CTA reduces); N > 1: + the combine of the ranks' sums
    launches = (3 if fused is not None else 2) * K
    value_default, ms_default, sim_ms_default = time_steps(step_default, n_total, K, warm=3, with_sim_events=True)

    # ---- end to end through the public API with host buffers (H2D of the step's constant tables from pinned memory, D2H of
    # ---- the OFE vector - every rank its own slice - and of the density on rank 0) ----
    import ctypes as C
    nbytes = nat.lib.at2_hb_tables_bytes(C.byref(spec.plan), C.byref(spec.cfg), 0)
    host_tab = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    host_ofe = torch.empty(n_local, dtype=torch.float32).pin_memory()
    host_den = torch.empty(KDE_G, dtype=torch.float32).pin_memory()
    tab_key = (str(dev), torch.float32, bytes(spec.plan))

    def make_e2e(spec_, n_tot, n_loc, fg, early_stop, kde):
        # the step's host inputs: the configuration's constant tables, built once into pinned memory, uploaded every step
        nat.check(nat.lib.at2_hb_tables_build(C.byref(spec_.plan), C.byref(spec_.cfg), 0, host_tab.data_ptr()))

        def e2e_step(seed):
            spec_._tables[tab_key] = host_tab.to(dev, non_blocking=True)
            res = run_epdf_ofe(spec_, n_tot, seed, start=kde[0], end=kde[1], bandwidth=kde[2], grid_points=kde[3],
                               device=dev, fused=fg, early_stop=early_stop)
            host_ofe[:n_loc].copy_(res['ofes'][rank * n_loc:(rank + 1) * n_loc], non_blocking=True)
            if rank == 0:
                host_den.copy_(res['density'], non_blocking=True)
            torch.cuda.synchronize(dev)
            return host_den[0].item()
        return e2e_step

    def time_e2e(fn, n_tot, steps):
        for i in range(3):
            fn(5000 + i)
        barrier()
        w0 = time.perf_counter()
        for i in range(steps):
            fn(6000 + i)
        barrier()
        return n_tot * steps / max_over_ranks(time.perf_counter() - w0)

    kde_hl = (KDE_START, KDE_END, KDE_H, KDE_G)
    e2e_value = time_e2e(make_e2e(spec, n_total, n_local, fused, False, kde_hl), n_total, K)
    e2e_default = time_e2e(make_e2e(spec, n_total, n_local, fused, True, kde_hl), n_total, K)
    # a reference-sized job (7 000 repetitions, run_simulation.py:31) end to end: what the op-call overhead looks like
    e2e_small = None
    if world == 1:
        e2e_small_runs = time_e2e(make_e2e(spec, 7000, 7000, None, True, kde_hl), 7000, max(K, 20))
        e2e_small = {'repetitions': 7000, 'runs_per_s': e2e_small_runs, 'ms_per_job': 7000 / e2e_small_runs * 1e3,
                     'path': 'product default (early stopping) through run_epdf_ofe, host buffers, synchronised per job'}

    # ---- N > 1: the fused gather is checked in this very run against NCCL and against a single-rank recomputation ----
    gather_check = None
    if world > 1:
        ok = True
        for es in (False, True):
            seed_chk = 424242 + int(es)
            got = (fused.run(spec, seed_chk, es) if fused is not None else None)
            mine = run_hyperband_repetitions(spec, n_local, seed=seed_chk, rep_offset=rank * n_local, device=dev,
                                             early_stop=es)['ofe']
            nccl = gather_slices(mine, n_total)
            peer = (rank + 1) % world
            foreign = run_hyperband_repetitions(spec, n_local, seed=seed_chk, rep_offset=peer * n_local, device=dev,
                                                early_stop=not es)['ofe']       # the other kernel: same bits
            ok = ok and torch.equal(nccl[peer * n_local:(peer + 1) * n_local], foreign)
            if got is not None:
                ok = ok and torch.equal(got, nccl)
        flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = ('bit-identical' if int(flag.item()) == 1 else 'MISMATCH') + \
            (' (fused gather == NCCL all-gather == single-rank recomputation of a foreign slice; full and early-stopping kernels)'
             if fused is not None else ' (NCCL all-gather == single-rank recomputation of a foreign slice)')

    extras = {}
    # ---- N > 1 (collective: every rank takes part): strong scaling of config 2 as written and config 5's sweep ----
    if world > 1 and not args.no_extras:
        def sharded_job(spec_, n_tot, kde, label):
            if n_tot % world:
                return
            fg = None if args.nccl_gather else FusedGather.create(n_tot, dev)
            res = {}
            for es, key in ((False, 'full_simulation'), (True, 'product_default_early_stop')):
                runs, ms, _ = time_steps(make_step(spec_, n_tot // world, n_tot, fg, es, kde), n_tot, 8, warm=3)
                res[key] = {'ms_per_job': ms, 'runs_per_s': runs}
            res['repetitions_total'], res['repetitions_per_gpu'] = n_tot, n_tot // world
            extras[label] = res
            del fg
        sharded_job(spec, 100000, kde_hl, f'strong_scaling_1e5_total_repetitions_rastrigin6_R81_x{world}gpus')
        s_e = HyperbandSimSpec('egg', EGG3, 243, 3, 10)
        for n_tot in (1000, 10000, 100000, 1000000):
            sharded_job(s_e, n_tot - n_tot % world, (-2400.0, -1300.0, 30.0, 1000),
                        f'config5_egg3_R243_{n_tot}_total_repetitions_x{world}gpus')

    # ---- other configurations of BASELINE.json, device-timed on rank 0's GPU ----
    if rank == 0 and not args.no_extras:
        from autotune.b200.engine import run_hybrid_repetitions
        peaks_hbm = measured_peaks()[0].get('hbm_gbs')

        def timed(fn, n_units, steps=8):
            for i in range(3):
                fn(900 + i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(steps):
                fn(i)
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / steps
            return {'ms_per_step': ms, 'units_per_step': n_units, 'runs_per_s': n_units / (ms * 1e-3)}
        s_b = HyperbandSimSpec('branin', FLAT, 81, 3, 0)
        s_e = HyperbandSimSpec('egg', EGG3, 243, 3, 10)
        for label, sp_, n in (('hb_flat_branin_R81_1e5reps (BASELINE.json configs[1])', s_b, 100000),
                              ('hb_egg_3families_R243_2e4reps', s_e, 20000)):
            extras[label] = k = timed(lambda sd: run_hyperband_repetitions(sp_, n, seed=sd, device=dev, early_stop=False), n)
            k['arm_epochs_per_s'] = k['runs_per_s'] * sp_.arm_epochs_per_repetition
            k['alg_tflops'] = k['arm_epochs_per_s'] * FLOP_PER_ARM_EPOCH / 1e12
            k['frac_of_measured_ffma_peak'] = k['alg_tflops'] / fma_peak_tflops
            extras[label + ' early_stop'] = timed(
                lambda sd: run_hyperband_repetitions(sp_, n, seed=sd, device=dev, early_stop=True), n)
        if world == 1:   # config 5's sweep on one GPU (N > 1: the sharded sweep above)
            for n in (1000, 10000, 100000, 1000000):
                extras[f'config5_egg3_R243_{n}_total_repetitions_x1gpus'] = {
                    'full_simulation': timed(lambda sd: run_hyperband_repetitions(s_e, n, seed=sd, device=dev, early_stop=False), n, 4),
                    'product_default_early_stop': timed(
                        lambda sd: run_hyperband_repetitions(s_e, n, seed=sd, device=dev, early_stop=True), n, 4)}
        extras['hybrid_tpe_transfer_all_rastrigin_R81_2e4reps'] = timed(
            lambda sd: run_hybrid_repetitions(spec, 20000, 'all', seed=sd, device=dev), 20000, steps=4)
        try:
            from autotune.b200.engine import KnownFnSpec, run_knownfn_hyperband
            from autotune.benchmarks.mnist_like import synthetic_mnist_database
            kspec = KnownFnSpec(*synthetic_mnist_database(425, 300, seed=0), device=dev)
            extras['hb_closest_known_fn_mnist_shape_A425_D4_R81_1e4reps'] = timed(
                lambda sd: run_knownfn_hyperband(kspec, 81, 3, 10000, seed=sd), 10000, steps=4)
        except Exception as e:  # noqa: BLE001
            extras['hb_closest_known_fn_mnist_shape_A425_D4_R81_1e4reps'] = {'error': repr(e)}
        big = torch.randn(1000000, device=dev)
        extras['kde_epanechnikov_1e6_samples_1000_grid'] = k = timed(lambda sd: density(big, -4.0, 4.0, 0.3, 1000), 1000000)
        k['fp32_tflops_6_flop_per_pair'] = 6.0 * 1e6 * 1000 / (k['ms_per_step'] * 1e-3) / 1e12
        k['frac_of_measured_ffma_peak'] = k['fp32_tflops_6_flop_per_pair'] / fma_peak_tflops
        k['input_gb_per_s'] = (4e6 + 4e3) / (k['ms_per_step'] * 1e-3) / 1e9
        try:   # scaled-up nearest-recorded-arm scan: A = 2^20 recorded arms x D = 8, 2^13 queries
            from autotune.b200 import ops as at2ops
            from autotune.b200.engine import domain_desc
            from autotune.core import Domain, Param
            dom8 = Domain(**{f'h{i}': Param(f'h{i}', 0.0, 1.0, distrib='uniform', scale='linear') for i in range(8)})
            desc8, names8 = domain_desc(dom8, tuple(dom8.hyperparams_names()))
            dbytes = at2ops.struct_tensor(desc8)
            A8, Q8 = 1 << 20, 1 << 13
            db8 = torch.rand(A8, 8, device=dev, dtype=torch.float64)
            q8 = torch.rand(Q8, 8, device=dev, dtype=torch.float64)
            none_c, none_t = db8.new_empty((0, 0)), torch.empty(0, device=dev, dtype=torch.int32)
            extras['nn_lookup_A2^20_D8_2^13_queries'] = k = timed(
                lambda sd: torch.ops.at2.nn_lookup(dbytes, db8, q8, none_c, none_t), Q8, steps=3)
            dp_instr = float(Q8) * A8 * (3 * 8 - 1)
            k['lookups_per_s'] = k.pop('runs_per_s')
            k['fp64_instr_per_s'] = dp_instr / (k['ms_per_step'] * 1e-3)
            k['frac_of_fp64_issue_rate'] = k['fp64_instr_per_s'] / (148 * 64 * 1.965e9)
            k['db_stream_gb_per_s'] = (Q8 / 128) * A8 * 8 * 8 / (k['ms_per_step'] * 1e-3) / 1e9
            k['note'] = ('float64, un-fused, first-minimum; every 128-query CTA streams the 64 MB table (L2-resident) once; '
                         'FP64-pipe bound (3D-1 DP instructions per recorded arm per query), not HBM bound: '
                         f'hbm copy peak {peaks_hbm} GB/s')
            del db8, q8
        except Exception as e:  # noqa: BLE001
            extras['nn_lookup_A2^20_D8_2^13_queries'] = {'error': repr(e)}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        achieved_tflops = n_local * ARM_EPOCHS * FLOP_PER_ARM_EPOCH / (sim_ms * 1e-3) / 1e12
        traffic, traffic_src = measured_traffic()
        sass_flop, _ = measured_traffic('sass_fp32_flop_per_launch')
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'reps_per_gpu_per_step': n_local, 'arms_per_rep': ARMS,
                       'arm_epochs_per_rep': ARM_EPOCHS, 'kde': 'epanechnikov h=3 on linspace(-400,-370,1000)',
                       'parallelism': f'repetition-sharded x{world}; OFE all-gather: {gather_kind}',
                       'path': 'full simulation (every arm to R, as the reference does); value_product_default: early stopping',
                       'l2': 'no HBM inputs (Philox RNG in-kernel); per-step outputs 1.6 MB; seed changes every step'},
            'arm_epochs_per_s': value * ARM_EPOCHS,
            'value_product_default': value_default,
            'product_default': {'value': value_default, 'unit': UNIT, 'ms_per_step': ms_default, 'kernel_ms': sim_ms_default,
                                'e2e': e2e_default,
                                'what': 'the same job through the default path of run_repetitions / run_epdf_ofe / the CLI: '
                                        'early stopping (arms simulated only as far as the round that reads them needs; a raw '
                                        'arm read at R is its target f_n) - OFEs, winners, round-bests, survivors bit-identical '
                                        'to the full simulation (tests/test_hb_gpu.py)'},
            'clocks': sampler.summary(t0, t1),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(nbytes) * world,
                    'd2h_bytes_per_step': int(4 * n_total + 4 * KDE_G),
                    'note': 'full-simulation path; per step every rank uploads the constant tables from pinned memory and reads '
                            'its own OFE slice back, rank 0 also the density'},
            'gpu_launches': launches,
            'roofline': {'kernel': 'hb_sim_kernel<float,PHILOX,HAS_SMOOTH,CK=5> (fused simulation + successive halving)',
                         'bound': 'fp32', 'achieved': achieved_tflops, 'peak': fma_peak_tflops, 'unit': 'TFLOP/s',
                         'frac': achieved_tflops / fma_peak_tflops, 'traffic': traffic,
                         'traffic_note': f'DRAM bytes per launch from the committed ncu capture ({traffic_src}); the kernel has '
                                         'no HBM inputs, its 1.6 MB of outputs stay in L2 during the launch',
                         'peak_source': 'best of 3 FFMA micro-benchmark variants measured in this run '
                                        '(at2_fma_peak_probe); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s',
                         'fma_probe_variants_tflops': fma_variants,
                         'sass_counted_fp32_flop_per_launch': sass_flop,
                         'kernel_ms': sim_ms, 'algorithmic_flop_per_launch': n_local * ARM_EPOCHS * FLOP_PER_ARM_EPOCH,
                         'hbm_peak_gbs': peaks.get('hbm_gbs'), 'hbm_peak_source': peak_src},
        }
        if e2e_small is not None:
            line['e2e_reference_sized_job'] = e2e_small
        if gather_check is not None:
            line['gather_check'] = gather_check
        line['extras'] = extras
        if cpu_ctx is not None:       # after every GPU measurement: see cpu_baseline_prepare
            line['cpu_baseline'] = cpu_baseline_measure(cpu_ctx)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
