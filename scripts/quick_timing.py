"""Ad-hoc device timing of the fused Hyperband simulation kernel (development aid, not the bench contract)."""
import sys
import os
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200.engine import HyperbandSimSpec, run_hyperband_repetitions

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]
EGG = [(1.5, 10, 15, False, 0, 1000), (0.5, 7, 10, False, 0, 1000), (0.2, 4, 7, True, 0, 1000)]

for name, spec, n in [('branin-flat R81', HyperbandSimSpec('branin', FLAT, 81, 3, 0), 100000),
                      ('rastrigin-6fam R81', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10), 100000),
                      ('egg-3fam R243', HyperbandSimSpec('egg', EGG, 243, 3, 10), 20000)]:
    for _ in range(30):          # ~100 ms of work: lets the clocks ramp up before timing
        run_hyperband_repetitions(spec, n, seed=0, early_stop=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for i in range(reps):
        out = run_hyperband_repetitions(spec, n, seed=i, early_stop=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ae = spec.arm_epochs_per_repetition * n
    print(f'{name}: {n} reps in {ms:.3f} ms -> {n / ms * 1e3:.3e} runs/s, {ae / ms * 1e3:.3e} arm-epochs/s, '
          f'{ae * 32 / ms * 1e3 / 1e12:.2f} alg-TFLOP/s; ofe mean {out["ofe"].mean().item():.4f}', flush=True)
    if '--lazy' in sys.argv:
        for _ in range(10):
            run_hyperband_repetitions(spec, n, seed=0, early_stop=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            out = run_hyperband_repetitions(spec, n, seed=i, early_stop=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f'   early-stop: {ms:.3f} ms -> {n / ms * 1e3:.3e} runs/s; ofe mean {out["ofe"].mean().item():.4f}', flush=True)
