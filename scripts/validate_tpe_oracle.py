"""Distributional check of the (unpinned) TPE restatement in oracle/tpe.py + oracle/hybrid.py against the reference's
stored hybrid OFE vectors (tests/golden/epdf_ofes.npz).  CPU only; ~N*0.36 s of work per variant.

    python scripts/validate_tpe_oracle.py [N] [variant ...]
"""
import multiprocessing as mp
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from scipy.stats import ks_2samp  # noqa: E402

from oracle import hybrid, hyperband  # noqa: E402

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]


def work(args):
    variant, seed, n = args
    np.random.seed(seed)
    random.seed(seed)
    rs = np.random.RandomState(seed + 10 ** 6)
    if variant == 'hb':
        return [hyperband.run_global_rng('rastrigin', FAM6, 81, 3, 10) for _ in range(n)]
    return [hybrid.run_hybrid('rastrigin', FAM6, 81, 3, 10, variant, rstate=rs) for _ in range(n)]


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    variants = sys.argv[2:] or ['none', 'all']
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'epdf_ofes.npz'))
    procs = os.cpu_count()
    with mp.Pool(procs) as pool:
        for v in variants:
            per = -(-n // procs)
            got = np.asarray([o for part in pool.map(work, [(v, 100 + w, per) for w in range(procs)]) for o in part])
            line = f'{v}: n={len(got)} mean={got.mean():.3f} sd={got.std(ddof=1):.3f}'
            for ref in ['hb', 'hb+tpe+transfer+none', 'hb+tpe+transfer+all', 'hb+tpe+transfer+same',
                        'hb+tpe+transfer+longest']:
                want = g[f'simulation-rastrigin-families-2/hist-sim-rastrigin-sim({ref})-7000-1']
                r = ks_2samp(got, want)
                line += f' | vs {ref}: D={r.statistic:.3f} p={r.pvalue:.3g}'
            print(line, flush=True)
