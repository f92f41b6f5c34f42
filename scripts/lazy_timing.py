"""Ad-hoc device timing of the early-stopping Hyperband kernel (development aid, not the bench contract).
usage: python scripts/lazy_timing.py [n_reps] [config: flat|fam6|egg]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
EGG = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
which = sys.argv[2] if len(sys.argv) > 2 else 'flat'
spec = {'flat': HyperbandSimSpec('branin', [(0, 1, 0, False, 0, 0)], 81, 3, 0),
        'fam6': HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10),
        'egg': HyperbandSimSpec('egg', EGG, 243, 3, 10)}[which]
for _ in range(20):
    run_hyperband_repetitions(spec, n, seed=0, early_stop=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    out = run_hyperband_repetitions(spec, n, seed=i, early_stop=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
full = run_hyperband_repetitions(spec, n, seed=19)
same = bool(torch.equal(full['ofe'], out['ofe']) and torch.equal(full['ofe_arm'], out['ofe_arm']))
print(f'{which}: early-stop {n} reps in {ms:.3f} ms -> {n / ms * 1e3:.3e} runs/s; ofe mean {out["ofe"].mean().item():.4f}; '
      f'identical to the full simulation: {same}', flush=True)
