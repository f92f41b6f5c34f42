"""Ad-hoc device timing of the closest-known-function kernels (development aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import io  # noqa: E402
from contextlib import redirect_stdout  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from autotune.b200.engine import KnownFnSpec, run_knownfn_hyperband  # noqa: E402
from autotune.benchmarks.mnist_like import cifar_domain_problem, mnist_domain_problem, synthetic_database  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for name, mk, n_arms in (('mnist A=425 D=4', mnist_domain_problem, 425), ('cifar A=300 D=9', cifar_domain_problem, 300)):
    with redirect_stdout(io.StringIO()):
        prob = mk()
    ks = KnownFnSpec(synthetic_database(prob, n_arms, 300, seed=0), prob.domain, prob.hyperparams_to_opt)
    for _ in range(5):
        run_knownfn_hyperband(ks, 81, 3, n, seed=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        out = run_knownfn_hyperband(ks, 81, 3, n, seed=i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name}: HB(R=81) x {n} reps in {ms:.3f} ms -> {n / ms * 1e3:.3e} runs/s, {n * 117 / ms * 1e3:.3e} lookups/s; '
          f'ofe mean {out["ofe"].mean().item():.5f}', flush=True)
    q = torch.rand(1 << 20, ks.db.shape[1], device='cuda', dtype=torch.float64)
    lo = torch.tensor([prob.domain[k].min_val for k in ks.names], device='cuda', dtype=torch.float64)
    q = ks.db[torch.randint(0, n_arms, (1 << 20,), device='cuda')] * (0.9 + 0.2 * q)
    for _ in range(3):
        ks.lookup(q)
    torch.cuda.synchronize()
    e0.record()
    for i in range(5):
        ks.lookup(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'   standalone lookup of 2^20 queries: {ms:.3f} ms -> {(1 << 20) / ms * 1e3:.3e} lookups/s', flush=True)
