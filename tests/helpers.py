"""Shared helpers for the tests: golden-fixture loading."""
import ast
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def hb_case_names():
    return sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, 'hb_*.npz')))


def load_hb_case(name):
    d = np.load(os.path.join(GOLDEN, f'hb_{name}.npz'))
    cfg = dict(func=str(d['func']), families=[tuple(r) for r in d['families'].tolist()], R=int(d['R']),
               eta=int(d['eta']), noise=float(d['noise']), min_or_max=str(d['min_or_max']), n_reps=int(d['n_reps']),
               n_arms=int(d['n_arms']))
    reps = []
    for r in range(cfg['n_reps']):
        pre = f'rep{r}_'
        reps.append({k[len(pre):]: d[k] for k in d.files if k.startswith(pre)})
    return cfg, reps


def load_seeded():
    d = np.load(os.path.join(GOLDEN, 'seeded_ofes.npz'))
    out = {}
    for k in d.files:
        if k.endswith('_cfg'):
            continue
        out[k] = (ast.literal_eval(str(d[k + '_cfg'])), d[k])
    return out
