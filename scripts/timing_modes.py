"""Development aid: per-launch vs back-to-back device timing of the fused Hyperband kernel (flat Branin), with SM clocks."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200._native import debug_options
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions
spec = HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0)
n = 100000
def smi():
    return subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active', '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
for opts in ({}, {'hb_static_schedule': 1}, {}):
  with debug_options(**opts):
    for _ in range(60):
        run_hyperband_repetitions(spec, n, seed=0, early_stop=False)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for i, (a, b) in enumerate(ev):
        a.record(); run_hyperband_repetitions(spec, n, seed=i, early_stop=False); b.record()
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in ev]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run_hyperband_repetitions(spec, n, seed=i, early_stop=False)
    e1.record()
    torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 20
    e0.record()
    for i in range(200):
        run_hyperband_repetitions(spec, n, seed=i, early_stop=False)
    e1.record()
    torch.cuda.synchronize()
    b2b200 = e0.elapsed_time(e1) / 200
    print(opts, 'per-launch', ' '.join(f'{t:.3f}' for t in per), '| back-to-back x20', f'{b2b:.3f}', 'x200', f'{b2b200:.3f}', '|', smi(), flush=True)
