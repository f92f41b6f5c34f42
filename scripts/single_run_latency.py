"""Development aid: latency of ONE optimiser.run_optimisation(problem) call through the drop-in class API (what the
reference's own repetition loop, experiments/run_simulation.py:147-157, would pay per iteration).
usage: python scripts/single_run_latency.py"""
import io
import os
import sys
import time
from contextlib import redirect_stdout
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import torch  # noqa: E402

from autotune.benchmarks import OptFunctionSimulationProblem  # noqa: E402
from autotune.core import ShapeFamily, UniformShapeFamilyScheduler  # noqa: E402
from autotune.optimisers import HyperbandOptimiser  # noqa: E402

families = (ShapeFamily(None, 1.5, 10, 15, False), ShapeFamily(None, 0.5, 7, 10, False), ShapeFamily(None, 0.2, 4, 7, True),
            ShapeFamily(None, 1.5, 10, 15, False, 200, 400), ShapeFamily(None, 0.5, 7, 10, False, 200, 400),
            ShapeFamily(None, 0.2, 4, 7, True, 200, 400))
with redirect_stdout(io.StringIO()):
    problem = OptFunctionSimulationProblem('rastrigin')


def one():
    scheduler = UniformShapeFamilyScheduler(families, max_resources=81, init_noise=10)
    opt = HyperbandOptimiser(eta=3, max_iter=81, min_or_max=min, optimisation_func=lambda g: g.fval, is_simulation=True,
                             scheduler=scheduler)
    with redirect_stdout(io.StringIO()):
        return opt.run_optimisation(problem, verbosity=False)


for _ in range(20):
    best = one()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 200
for _ in range(n):
    best = one()
dt = (time.perf_counter() - t0) / n
print(f'one run_optimisation call: {dt * 1e3:.3f} ms -> {1 / dt:.0f} runs/s through the unbatched class API; '
      f'best fval {best.optimisation_goals.fval:.3f}')
import cProfile  # noqa: E402
import pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    one()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22)
print(s.getvalue()[:3500])
