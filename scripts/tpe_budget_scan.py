"""Which evaluation count reproduces the reference's stored stand-alone TPE vectors (sim(tpe), sim(tpe2xbudget))?  The
pickles do not record max_iter; this scans it on the device.  usage: python scripts/tpe_budget_scan.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import numpy as np  # noqa: E402
from scipy.stats import ks_2samp  # noqa: E402

from autotune.b200.engine import HyperbandSimSpec, run_tpe_repetitions  # noqa: E402

FAM6 = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200),
        (1.5, 10, 15, False, 200, 400), (0.5, 7, 10, False, 200, 400), (0.2, 4, 7, True, 200, 400)]
FLAT = [(0, 1, 0, False, 0, 0)]
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'epdf_ofes.npz'))
for pre, func, fams, noise in [('simulation-flat-branin/hist-', 'branin', FLAT, 0),
                               ('simulation-flat-drop-wave/hist-sim-wave-', 'wave', FLAT, 0),
                               ('simulation-flat-rastrigin/hist-sim-rastrigin-', 'rastrigin', FLAT, 0),
                               ('simulation-rastrigin-families-2/hist-sim-rastrigin-', 'rastrigin', FAM6, 10)]:
    spec = HyperbandSimSpec(func, fams, 81, 3, noise)
    for m, ns in (('tpe', (11, 12, 13, 14, 15)), ('tpe2xbudget', (24, 25, 26, 27, 28, 30))):
        key = f'{pre}sim({m})-7000-1'
        if key not in g.files:
            continue
        want = g[key]
        line = f'{key}: mean {want.mean():.3f} median {np.median(want):.3f} |'
        for n in ns:
            for res in ((24,) if len(fams) == 1 else (24, 81)):
                got = run_tpe_repetitions(spec, 20000, res, n, seed=23)['ofe'].cpu().numpy().astype(np.float64)
                r = ks_2samp(got, want)
                line += f' n={n}' + (f',r={res}' if len(fams) > 1 else '') + f': D={r.statistic:.3f} p={r.pvalue:.2g} (mean {got.mean():.3f}) |'
        print(line, flush=True)
