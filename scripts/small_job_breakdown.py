"""Development aid: where the host time of a reference-sized EPDF-OFE job (7 000 repetitions) goes.
usage: python scripts/small_job_breakdown.py"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch  # noqa: E402

from autotune.b200 import ops  # noqa: E402
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions  # noqa: E402
from autotune.b200.epdf import density, run_epdf_ofe  # noqa: E402

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
dev = torch.device('cuda', 0)
n = 7000
host_ofe = torch.empty(n, dtype=torch.float32).pin_memory()
host_den = torch.empty(1000, dtype=torch.float32).pin_memory()
for _ in range(200):
    run_epdf_ofe(spec, n, 1, start=-400.0, end=-370.0, bandwidth=3.0, device=dev)
torch.cuda.synchronize()


def timeit(name, fn, reps=300, sync=True):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        if sync:
            torch.cuda.synchronize()
    if not sync:
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f'{name:60s} host-only {t_host / reps * 1e6:8.1f} us   with drain {(time.perf_counter() - t0) / reps * 1e6:8.1f} us')
    else:
        print(f'{name:60s} {(time.perf_counter() - t0) / reps * 1e6:8.1f} us per call (synchronised)')


tab = spec.tables(dev, torch.float32)
ofe = run_hyperband_repetitions(spec, n, seed=1, device=dev)['ofe']
timeit('empty synchronize', lambda: None)
timeit('spec.tables lookup', lambda: spec.tables(dev, torch.float32), sync=False)
timeit('hb_sim_direct early stop (enqueue only)', lambda: ops.hb_sim_direct(tab, spec.plan, spec.cfg, n, 0, 1, False, True), sync=False)
timeit('hb_sim_direct early stop', lambda: ops.hb_sim_direct(tab, spec.plan, spec.cfg, n, 0, 1, False, True))
timeit('run_hyperband_repetitions', lambda: run_hyperband_repetitions(spec, n, seed=1, device=dev))
timeit('torch.ops.at2.hb_sim_early_stop (dispatcher)', lambda: torch.ops.at2.hb_sim_early_stop(tab, spec.plan_bytes, spec.cfg_bytes, n, 0, 1, False))
timeit('density (enqueue only)', lambda: density(ofe, -400.0, -370.0, 3.0, 1000), sync=False)
timeit('density', lambda: density(ofe, -400.0, -370.0, 3.0, 1000))
timeit('run_epdf_ofe', lambda: run_epdf_ofe(spec, n, 1, start=-400.0, end=-370.0, bandwidth=3.0, device=dev))
res = run_epdf_ofe(spec, n, 1, start=-400.0, end=-370.0, bandwidth=3.0, device=dev)
timeit('2 x D2H copy_ non_blocking + sync', lambda: (host_ofe.copy_(res['ofes'], non_blocking=True), host_den.copy_(res['density'], non_blocking=True)))


def whole():
    r = run_epdf_ofe(spec, n, 1, start=-400.0, end=-370.0, bandwidth=3.0, device=dev)
    host_ofe.copy_(r['ofes'], non_blocking=True)
    host_den.copy_(r['density'], non_blocking=True)


timeit('whole job incl. D2H', whole)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100):
    ops.hb_sim_direct(tab, spec.plan, spec.cfg, n, 0, i, False, True)
e1.record()
torch.cuda.synchronize()
print(f'device time of the early-stopping kernel, back to back: {e0.elapsed_time(e1) * 10:.1f} us per launch')
