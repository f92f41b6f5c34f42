"""BASELINE.json configs[4]: scaling sweep, 10^3..10^6 Hyperband(R=243, eta=3) repetitions on the Egg-holder simulation
(families of experiments/run_simulation.py:36-40, init_noise 10), STRONG scaling: the repetitions of one job are
sharded over the ranks, OFEs all-gathered, KDE on every rank.  One JSON line per job size on rank 0.

  python scripts/scaling_sweep.py [--func egg|egg4] [--sizes 1000,10000,...]            one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
      scripts/scaling_sweep.py ...                                                       N GPUs of one box

'egg' is the reference's 2-D function (the parity-pinned one); 'egg4' is the 4-D extension (no reference equivalent).
Device-timed with CUDA events (max over ranks), 3 warm-up + 5 timed jobs per size, a fresh seed per job.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from autotune.b200.engine import HyperbandSimSpec  # noqa: E402
from autotune.b200.epdf import FusedGather, run_epdf_ofe  # noqa: E402

EGG_FAMILIES = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--func', default='egg', choices=['egg', 'egg4'])
    ap.add_argument('--sizes', default='1000,10000,100000,1000000')
    ap.add_argument('--max-iter', type=int, default=243)
    ap.add_argument('--early-stop', action='store_true')
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    spec = HyperbandSimSpec(args.func, EGG_FAMILIES, args.max_iter, 3, 10)
    stream = torch.cuda.current_stream(dev)
    runner = None
    if args.early_stop:
        from autotune.b200.engine import run_hyperband_repetitions

        def runner(spec_, n, seed_, off, d):
            return run_hyperband_repetitions(spec_, n, seed=seed_, rep_offset=off, device=d, early_stop=True)['ofe']
    # KDE window of the egg families: the curves end near f - 1000
    lo, hi, h = (-2100.0, -1500.0, 15.0) if args.func == 'egg' else (-4200.0, -2500.0, 30.0)
    for n in (int(x) for x in args.sizes.split(',')):
        # N > 1, plain Hyperband: the OFE all-gather is fused into the simulation kernel (None -> NCCL all-gather)
        fused = FusedGather.create(n, dev) if (world > 1 and runner is None) else None

        def job(seed):
            return run_epdf_ofe(spec, n, seed, start=lo, end=hi, bandwidth=h, device=dev, runner=runner, fused=fused)
        for i in range(3):
            job(100 + i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a.record(stream)
        for i in range(5):
            res = job(i)
        b.record(stream)
        torch.cuda.synchronize(dev)
        ms = torch.tensor([a.elapsed_time(b) / 5], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = ms.item()
            print(json.dumps({'workload': f'Hyperband(R={args.max_iter},eta=3) EPDF-OFE on the {args.func} simulation, '
                                          f'3 families, {n} repetitions' + (', early stop' if args.early_stop else ''),
                              'n_gpus': world, 'scaling': 'strong', 'repetitions': n, 'ms_per_job': ms,
                              'gather': 'fused' if fused is not None else ('nccl' if world > 1 else 'none'),
                              'runs_per_s': n / ms * 1e3,
                              'arm_epochs_per_s': n * spec.arm_epochs_per_repetition / ms * 1e3,
                              'ofe_mean': res['ofes'].double().mean().item(),
                              'density_integral': float(res['density'].double().sum().item() * (hi - lo) / 999)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
