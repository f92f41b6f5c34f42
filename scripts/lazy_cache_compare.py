"""Development aid: the early-stopping kernel with and without its set-up cache (debug option lazy_no_cache): device time and
bit-identity of the outputs.  usage: python scripts/lazy_cache_compare.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch  # noqa: E402

from autotune.b200._native import debug_options  # noqa: E402
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions  # noqa: E402

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
cases = [('branin-flat', HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0), 100000),
         ('rastrigin-6fam', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 100000),
         ('rastrigin-6fam-7000', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 7000),
         ('egg-3fam-R243', HyperbandSimSpec('egg', [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000),
                                                    (0.2, 4, 7, True, 0, 1000)], 243, 3, 10), 20000),
         ('rastrigin-raw2-R27', HyperbandSimSpec('rastrigin', FAM6[:2], 27, 3, 10), 200000)]
for _ in range(60):
    run_hyperband_repetitions(cases[0][1], 100000, seed=0, early_stop=False)
for name, spec, n in cases:
    line, ref = f'{name:22s}', None
    for off in (1, 0, 1, 0):
        with debug_options(lazy_no_cache=off):
            for _ in range(3):
                run_hyperband_repetitions(spec, n, seed=0, early_stop=True)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
            for i, (a, b) in enumerate(ev):
                a.record()
                run_hyperband_repetitions(spec, n, seed=i, early_stop=True)
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            out = run_hyperband_repetitions(spec, min(n, 5000), seed=5, early_stop=True, details=True)
            key = (out['ofe'], out['survivors'], out['round_best'])
            same = ref is None or all(torch.equal(x, y) for x, y in zip(ref, key))
            ref = ref or key
            line += f' | {"no cache" if off else "cache   "} {ts[len(ts) // 2]:.3f} ms same={same}'
    print(line, flush=True)
