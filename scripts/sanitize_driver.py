"""Small invocations of every kernel path, meant to run under compute-sanitizer (scripts/sanitize.sh)."""
import io
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from autotune.b200._native import check, lib
from autotune.b200.engine import (HyperbandSimSpec, KnownFnSpec, run_hybrid_repetitions, run_hyperband_repetitions,  # noqa: E402
                                  run_knownfn_hyperband, run_random_repetitions, run_tpe_repetitions)
from autotune.b200.epdf import density  # noqa: E402
from autotune.benchmarks.mnist_like import cifar_domain_problem, mnist_domain_problem, synthetic_database  # noqa: E402

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for func, fams, R, eta in (('rastrigin', FAM6, 81, 3), ('branin', FLAT, 81, 3), ('egg', FAM6[:3], 243, 3), ('egg4', FLAT, 27, 3),
                           ('wave', FAM6, 16, 2), ('rastrigin', FAM6, 729, 3)):
    spec = HyperbandSimSpec(func, fams, R, eta, 10)
    m = n if R <= 243 else 4
    a = run_hyperband_repetitions(spec, m, seed=1, details=True)
    b = run_hyperband_repetitions(spec, m, seed=1, details=True, early_stop=True)
    assert torch.equal(a['ofe'], b['ofe']) and torch.equal(a['survivors'], b['survivors'])
    if func != 'egg4' and R <= 243:
        for transfer in ('none', 'all'):
            run_hybrid_repetitions(spec, m, transfer, seed=1, details=True)
    print('ok', func, R, flush=True)
spec = HyperbandSimSpec('rastrigin', FAM6, 81, 3, 10)
run_tpe_repetitions(spec, n, 24, 81, seed=2)
run_random_repetitions(spec, n, 24, 81, seed=2)
check(lib.at2_debug_option(b'force_gtab', 1))
run_hyperband_repetitions(spec, n, seed=1, details=True)
run_hyperband_repetitions(spec, n, seed=1, details=True, early_stop=True)
check(lib.at2_debug_option(b'reset', 0))
for mk, n_arms in ((mnist_domain_problem, 425), (cifar_domain_problem, 300)):
    with redirect_stdout(io.StringIO()):
        prob = mk()
    ks = KnownFnSpec(synthetic_database(prob, n_arms, 300, seed=0), prob.domain, prob.hyperparams_to_opt)
    run_knownfn_hyperband(ks, 81, 3, n, seed=0, details=True)
    ks.lookup(ks.db[:100] * 1.01)
print('ok knownfn', flush=True)
density(torch.randn(5000, device='cuda'), -3.0, 3.0, 0.3, 1000)
density(torch.randn(5000, device='cuda'), -3.0, 3.0, 0.3, 1000, kernel='gaussian')
from autotune.b200 import ops  # noqa: E402
for dt in (torch.float32, torch.float64):      # stand-alone halving step: every width class
    lens = [1, 9, 32, 33, 128, 129, 256, 257, 700]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device='cuda')
    ops.segmented_topk(torch.randn(int(offs[-1]), device='cuda', dtype=dt), offs,
                       torch.tensor([max(1, k // 3) for k in lens], dtype=torch.int32, device='cuda'), 700, False)
torch.cuda.synchronize()
print('sanitize driver done', flush=True)
