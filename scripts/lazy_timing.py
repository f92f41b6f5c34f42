"""Ad-hoc device timing of the early-stopping Hyperband kernel (development aid, not the bench contract).
usage: python scripts/lazy_timing.py [n_reps[,n_reps...]] [config[,config...]: flat|fam6|egg] [G[,G...] (0 = chosen by the library)]
Every launch is timed with its own event pair; the median is reported (a fresh box stalls the host now and then)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200._native import check, lib
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
EGG = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]
ns = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '100000').split(',')]
which = (sys.argv[2] if len(sys.argv) > 2 else 'flat').split(',')
gs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '0').split(',')]
specs = {'flat': HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0),
         'fam6': HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10),
         'egg': HyperbandSimSpec('egg', EGG, 243, 3, 10)}
for _ in range(50):        # clocks up
    run_hyperband_repetitions(specs['flat'], 100000, seed=0)
for w in which:
    spec = specs[w]
    for n in ns:
        for G in gs:
            if G > 0:
                check(lib.at2_debug_option(b'lazy_reps_per_tile', int(G)))
            else:
                check(lib.at2_debug_option(b'reset', 0))
            for _ in range(5):
                run_hyperband_repetitions(spec, n, seed=0, early_stop=True)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
            for i, (e0, e1) in enumerate(ev):
                e0.record()
                out = run_hyperband_repetitions(spec, n, seed=i, early_stop=True)
                e1.record()
            torch.cuda.synchronize()
            ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
            ms = ts[len(ts) // 2]
            full = run_hyperband_repetitions(spec, n, seed=29)
            same = bool(torch.equal(full['ofe'], out['ofe']) and torch.equal(full['ofe_arm'], out['ofe_arm']))
            print(f'{w} n={n} G={G}: median {ms:.3f} ms (min {ts[0]:.3f}, max {ts[-1]:.3f}) -> {n / ms * 1e3:.3e} runs/s; '
                  f'identical to the full simulation: {same}', flush=True)
