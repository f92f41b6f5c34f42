"""Error statistics of the device Parzen scores (tpe_score: full and the lean form used inside tpe_sim) against the float64
oracle over many random suggestion problems: max |score error|, arg-max agreement, and the oracle's score gap where the
arg-max differs.  usage: python scripts/tpe_score_stats.py [n_problems]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'autotune-v2_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from autotune.b200 import ops  # noqa: E402
from autotune.b200.engine import tpe_config  # noqa: E402
from oracle import tpe as otpe  # noqa: E402

n_prob = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lo, hi, n_cand = -5.12, 5.12, 24
for n_obs in (12, 30, 80, 117):
    rng = np.random.default_rng(n_obs)
    obs = rng.uniform(lo, hi, (n_prob, n_obs)).astype(np.float32)
    loss = rng.normal(size=(n_prob, n_obs)).astype(np.float32)
    cand = rng.uniform(lo, hi, (n_prob, n_cand)).astype(np.float32)
    scores, _, lean, _ = torch.ops.at2.tpe_score(torch.from_numpy(obs).cuda(), torch.from_numpy(loss).cuda(), lo, hi,
                                                 ops.struct_tensor(tpe_config()), torch.from_numpy(cand).cuda())
    scores, lean = scores.cpu().numpy().astype(np.float64), lean.cpu().numpy().astype(np.float64)
    err, err_lean, flips, flips_lean, gaps = 0.0, 0.0, 0, 0, []
    for k in range(n_prob):
        below, above = otpe.split_trials(obs[k].astype(np.float64), loss[k].astype(np.float64))
        wb, mb, sb = otpe.adaptive_parzen_normal(below, 1.0, 0.5 * (hi + lo), hi - lo)
        wa, ma, sa = otpe.adaptive_parzen_normal(above, 1.0, 0.5 * (hi + lo), hi - lo)
        want = otpe.gmm1_lpdf(cand[k].astype(np.float64), wb, mb, sb, lo, hi) - \
            otpe.gmm1_lpdf(cand[k].astype(np.float64), wa, ma, sa, lo, hi)
        err = max(err, np.abs(scores[k] - want).max())
        d = lean[k] - want
        err_lean = max(err_lean, np.abs(d - d.mean()).max())
        a = int(np.argmax(want))
        if int(np.argmax(scores[k])) != a:
            flips += 1
        b = int(np.argmax(lean[k]))
        if b != a:
            flips_lean += 1
            gaps.append(want[a] - want[b])
    print(f'n_obs={n_obs}: {n_prob} problems, max|score - oracle|={err:.2e}, lean (shift removed)={err_lean:.2e}, '
          f'arg-max flips full={flips} lean={flips_lean}, oracle gaps at lean flips={["%.1e" % g for g in gaps[:8]]}', flush=True)
