"""Ad-hoc device timing of the fused Hyperband simulation kernel with a debug option (at2_debug_option) toggled in-process
(development aid, not the bench contract).  usage: python scripts/hb_timing.py [option] - times every configuration with the
option unset and set to 1, alternating, each launch with its own event pair (median reported), and checks that the outputs
agree."""
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
var = sys.argv[1] if len(sys.argv) > 1 else None
cases = [('branin-flat R81', HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0), 100000),
         ('rastrigin-6fam R81', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 100000),
         ('egg-3fam R243', HyperbandSimSpec('egg', EGG, 243, 3, 10), 20000)]
for _ in range(40):        # clocks up
    run_hyperband_repetitions(cases[0][1], 100000, seed=0, early_stop=False)


def timed(spec, n):
    for _ in range(3):
        run_hyperband_repetitions(spec, n, seed=0, early_stop=False)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(25)]
    for i, (e0, e1) in enumerate(ev):
        e0.record()
        out = run_hyperband_repetitions(spec, n, seed=i, details=False, early_stop=False)
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], run_hyperband_repetitions(spec, min(n, 20000), seed=77, details=True, early_stop=False)


for name, spec, n in cases:
    res = {}
    for rnd in range(2):
        for setting in ([None, '1'] if var else [None]):
            if var:
                check(lib.at2_debug_option(b'reset', 0))
                if setting:
                    check(lib.at2_debug_option(var.encode(), int(setting)))
            ms, out = timed(spec, n)
            res.setdefault(setting, []).append((ms, out))
    base = res[None][0][1]
    line = f'{name}: default {min(m for m, _ in res[None]):.3f} ms'
    if var:
        other = res['1'][0][1]
        same = all(torch.equal(base[k], other[k]) for k in base)
        line += f'; {var}=1 {min(m for m, _ in res["1"]):.3f} ms; outputs identical: {same}'
    ae = spec.arm_epochs_per_repetition * n
    print(line + f'  ({ae * 32 / min(m for m, _ in res[None]) * 1e3 / 1e12:.2f} alg-TFLOP/s)', flush=True)
