"""Ad-hoc device timing of the hybrid (Hyperband + TPE) kernel (development aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch
from autotune.b200.engine import HyperbandSimSpec, run_hybrid_repetitions, run_optimiser_repetitions, tpe_config

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for name, spec in [('rastrigin-6fam', HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)),
                   ('branin-flat', HyperbandSimSpec('branin', FLAT, 81, 3, 0))]:
    for mode, cfg in [('tpe-off', tpe_config(False)), ('none', tpe_config(True, 'none')), ('all', tpe_config(True, 'all'))]:
        for _ in range(2):
            run_optimiser_repetitions(spec, n, cfg, seed=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3):
            out = run_optimiser_repetitions(spec, n, cfg, seed=i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f'{name} {mode}: {n} reps in {ms:.2f} ms -> {n / ms * 1e3:.3e} runs/s; ofe mean {out["ofe"].mean().item():.3f}', flush=True)
