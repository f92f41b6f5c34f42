"""Development aid: device time of the fused Hyperband kernel under different tile / CTA shapes (at2_debug_option).
usage: python scripts/occupancy_sweep.py 'hb_warps_per_cta=20,hb_singles=2' 'hb_static_schedule=1' ...   ('' = defaults)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200._native import debug_options
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
cases = [('branin-flat', HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0), 100000),
         ('rastrigin-6fam', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 100000),
         ('egg-3fam-R243', HyperbandSimSpec('egg', [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000),
                                                    (0.2, 4, 7, True, 0, 1000)], 243, 3, 10), 20000),
         ('wave-6fam-R27', HyperbandSimSpec('wave', FAM6, 27, 3, 10), 200000)]
for _ in range(60):
    run_hyperband_repetitions(cases[0][1], 100000, seed=0, early_stop=False)
ref = {}
for setting in (sys.argv[1:] or ['']):
    opts = dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in filter(None, setting.split(',')))
    line = f'{setting or "default":48s}'
    for name, spec, n in cases:
      with debug_options(**opts):
        try:
            for _ in range(3):
                run_hyperband_repetitions(spec, n, seed=0, early_stop=False)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
            for i, (a, b) in enumerate(ev):
                a.record()
                out = run_hyperband_repetitions(spec, n, seed=i, early_stop=False)
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            chk = run_hyperband_repetitions(spec, 5000, seed=5, early_stop=False)['ofe']
            same = torch.equal(ref.setdefault(name, chk), chk)
            line += f' | {name} {ts[len(ts) // 2]:.3f} ms (min {ts[0]:.3f}) same={same}'
        except Exception as e:  # noqa: BLE001
            line += f' | {name} ERROR {str(e)[:60]}'
    print(line, flush=True)
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active',
                      '--format=csv,noheader'], capture_output=True, text=True).stdout.strip(), flush=True)
