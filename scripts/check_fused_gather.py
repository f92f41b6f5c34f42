"""Multi-GPU check of the fused simulate + all-gather path (FusedGather) against the NCCL all-gather path.  Launch with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_fused_gather.py
Every rank verifies: gathered OFE vectors bit-identical (peer stores, multicast stores, NCCL) and equal to the single-process
run, densities of the sharded KDE equal to 2e-7 and bit-identical across the ranks, full and early-stopping kernels; then both paths are timed (CUDA events, max over ranks)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions  # noqa: E402
from autotune.b200.epdf import FusedGather, run_epdf_ofe  # noqa: E402

rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
dist.init_process_group('nccl', device_id=dev)
FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
report = {'world': world}
for name, spec, n, kde in (('flat-branin', HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0), 100000 * world + 3, (0.37, 1.5, 0.1)),
                           ('rastrigin-6fam', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 7001, (-400.0, -370.0, 3.0))):
    kw = dict(start=kde[0], end=kde[1], bandwidth=kde[2], device=dev)
    ref = run_epdf_ofe(spec, n, 5, **kw)                                   # NCCL all-gather
    single = run_hyperband_repetitions(spec, n, seed=5, device=dev)['ofe']  # everything on this GPU
    assert torch.equal(ref['ofes'], single), 'NCCL path differs from the single-GPU run'
    modes = {}
    for mode, use_mc in (('peer-stores', False), ('multicast', True)):
        fg = FusedGather.create(n, dev, use_multicast=use_mc)
        if fg is None:
            modes[mode] = 'unavailable'
            continue
        if use_mc and not fg.multicast:
            modes[mode] = 'no multicast support'
            continue
        for seed in (5, 6, 5):                                            # alternating buffers, re-used buffers
            got = run_epdf_ofe(spec, n, seed, fused=fg, early_stop=(seed == 6), **kw)
            want = ref if seed == 5 else run_epdf_ofe(spec, n, seed, **kw)
            assert torch.equal(got['ofes'], want['ofes']), f'{mode}: gathered OFEs differ from the NCCL path'
            # the KDE is sharded on the fused path (every rank sums its own slice, the float64 sums are added in rank order):
            # same density up to the last float bit of the differently grouped float64 sum, and the SAME bits on every rank
            err = (got['density'] - want['density']).abs().max().item()
            assert err <= 2e-7 * want['density'].abs().max().item(), f'{mode}: densities differ by {err}'
            every = [torch.empty_like(got['density']) for _ in range(world)]
            dist.all_gather(every, got['density'].contiguous())
            assert all(torch.equal(e, every[0]) for e in every), f'{mode}: the ranks hold different density bits'

        def timed(fn, steps=30):
            for i in range(5):
                fn(100 + i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize(dev)
            a.record()
            for i in range(steps):
                fn(i)
            b.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        modes[mode] = {'ms_per_job_fused': timed(lambda s: run_epdf_ofe(spec, n, s, fused=fg, **kw)),
                       'ms_per_job_nccl': timed(lambda s: run_epdf_ofe(spec, n, s, **kw))}
        del fg
    report[name] = {'repetitions': n, **modes}
if rank == 0:
    print(json.dumps(report), flush=True)
dist.destroy_process_group()
